// Microbenchmark: throughput of shared-memory atomics that target a RANDOM CTA of the same thread-block cluster
// (distributed shared memory). Question it answers: can a cluster of 16 SMs hold one private 480x640x2 u32 count
// frame (2.4 MB) in its combined shared memory and take events at a higher rate than the 135 G/s L2 RED ceiling?
#include <cooperative_groups.h>
#include <cuda_runtime.h>
#include <cstdint>
#include <cstdio>
namespace cg = cooperative_groups;

__device__ __forceinline__ uint32_t hash32(uint32_t x) { x ^= x >> 16; x *= 0x7feb352dU; x ^= x >> 15; x *= 0x846ca68bU; x ^= x >> 16; return x; }

__global__ void __launch_bounds__(1024, 1) k_dsmem_atomics(unsigned* out, int iters, int cells, int local_only) {
    extern __shared__ unsigned s[];
    cg::cluster_group cluster = cg::this_cluster();
    const unsigned cs = cluster.num_blocks();
    for (int i = threadIdx.x; i < cells; i += blockDim.x) s[i] = 0;
    cluster.sync();
    uint32_t h = hash32(blockIdx.x * 1024 + threadIdx.x + 1);
    const unsigned me = cluster.block_rank();
    for (int i = 0; i < iters; ++i) {
        h = hash32(h + i);
        const unsigned rank = local_only ? me : (h >> 20) % cs;
        unsigned* remote = cluster.map_shared_rank(s, rank);
        atomicAdd(remote + (h % cells), 1u);
    }
    cluster.sync();
    unsigned acc = 0;
    for (int i = threadIdx.x; i < cells; i += blockDim.x) acc += s[i];
    if (acc == 0xdeadbeefu) out[0] = acc;
}

int main() {
    unsigned* d_out;
    cudaMalloc(&d_out, 4);
    const int cells = 150 * 1024 / 4, iters = 256;
    cudaFuncSetAttribute(k_dsmem_atomics, cudaFuncAttributeMaxDynamicSharedMemorySize, cells * 4);
    cudaFuncSetAttribute(k_dsmem_atomics, cudaFuncAttributeNonPortableClusterSizeAllowed, 1);
    for (int cs : {1, 2, 4, 8, 16}) {
        for (int local_only : {1, 0}) {
            cudaLaunchConfig_t cfg = {};
            cfg.blockDim = dim3(1024);
            cfg.dynamicSmemBytes = cells * 4;
            cudaLaunchAttribute at[1];
            at[0].id = cudaLaunchAttributeClusterDimension;
            at[0].val.clusterDim.x = cs; at[0].val.clusterDim.y = 1; at[0].val.clusterDim.z = 1;
            cfg.attrs = at; cfg.numAttrs = 1;
            int max_clusters = 0;
            cfg.gridDim = dim3(cs);
            cudaOccupancyMaxActiveClusters(&max_clusters, k_dsmem_atomics, &cfg);
            if (max_clusters < 1) { printf("{\"cluster\": %d, \"error\": \"cannot co-schedule\"}\n", cs); continue; }
            cfg.gridDim = dim3(max_clusters * cs);
            cudaEvent_t e0, e1;
            cudaEventCreate(&e0); cudaEventCreate(&e1);
            for (int rep = 0; rep < 2; ++rep) {
                cudaEventRecord(e0);
                cudaLaunchKernelEx(&cfg, k_dsmem_atomics, d_out, iters, cells, local_only);
                cudaEventRecord(e1);
                cudaEventSynchronize(e1);
            }
            float ms = 0;
            cudaEventElapsedTime(&ms, e0, e1);
            cudaError_t err = cudaGetLastError();
            const double ops = (double)max_clusters * cs * 1024 * iters;
            printf("{\"cluster\": %d, \"clusters_resident\": %d, \"ctas\": %d, \"local_only\": %d, \"G_atomics_s\": %.1f, \"ms\": %.3f, \"err\": \"%s\"}\n",
                   cs, max_clusters, max_clusters * cs, local_only, ops / ms / 1e6, ms, cudaGetErrorString(err));
        }
    }
    return 0;
}
