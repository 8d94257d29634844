// tc_probe.cu (TEST-ONLY library tests/native/libevfly_tc_probe.so, not part of the product ABI) -- hardware probe for ONE assumption the halo-reuse convolution relies on: a UMMA
// shared-memory descriptor may start at a row that is NOT aligned to the swizzle atom (8 rows) when
// its base-offset field is set to (start_address >> 7) & 7, so that a tile TMA wrote once can be
// read by tcgen05.mma at a shift of 1 or 2 rows (the kw taps of a 3x3 conv). The probe loads
// 128+8 rows, issues the MMAs from row `shift`, and returns D; the test compares with the shifted GEMM.
#include "common.cuh"
#include <cuda.h>
#include <cuda_bf16.h>

namespace evfly {
namespace probe {

__device__ __forceinline__ uint32_t smem_u32(const void* p) { return (uint32_t)__cvta_generic_to_shared(p); }
__device__ __forceinline__ void mbar_init(uint64_t* bar, uint32_t count) { asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(smem_u32(bar)), "r"(count)); }
__device__ __forceinline__ void mbar_expect_tx(uint64_t* bar, uint32_t bytes) { asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(smem_u32(bar)), "r"(bytes) : "memory"); }
__device__ __forceinline__ void mbar_wait(uint64_t* bar, uint32_t parity) {
    asm volatile("{\n.reg .pred p;\nW_%=:\nmbarrier.try_wait.parity.shared::cta.b64 p, [%0], %1;\n@p bra D_%=;\nbra W_%=;\nD_%=:\n}\n" ::"r"(smem_u32(bar)), "r"(parity) : "memory");
}
__device__ __forceinline__ void tma_load_2d(void* dst, const CUtensorMap* map, uint64_t* bar, int c0, int c1) {
    asm volatile("cp.async.bulk.tensor.2d.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1, {%3, %4}], [%2];" ::"r"(smem_u32(dst)), "l"(map), "r"(smem_u32(bar)), "r"(c0), "r"(c1) : "memory");
}

template <int KC>
__global__ void __launch_bounds__(128, 1)
k_tc_shift_probe(const __grid_constant__ CUtensorMap map_a, const __grid_constant__ CUtensorMap map_b, float* __restrict__ out, int shift,
                 int use_base_offset) {
    constexpr int TN = 32, ROWS = 136;
    constexpr uint32_t LAYOUT = (KC == 64) ? 2u : 4u;
    constexpr uint32_t SBO = 8 * KC * 2;
    extern __shared__ uint8_t raw[];
    uint8_t* smem = reinterpret_cast<uint8_t*>((reinterpret_cast<uintptr_t>(raw) + 1023) & ~(uintptr_t)1023);
    uint8_t* sa = smem;                       // [136][KC] bf16, swizzled by TMA
    uint8_t* sb = smem + 32 * 1024;           // [32][KC]
    uint64_t* bar = reinterpret_cast<uint64_t*>(smem + 48 * 1024);
    uint64_t* done = bar + 1;
    uint32_t* tptr = reinterpret_cast<uint32_t*>(bar + 2);
    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    if (threadIdx.x == 0) {
        mbar_init(bar, 1);
        mbar_init(done, 1);
        asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
    }
    if (warp == 1) {
        asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(smem_u32(tptr)), "r"(32) : "memory");
        asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;" ::: "memory");
    }
    asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
    __syncthreads();
    asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
    const uint32_t tmem = *tptr;
    if (threadIdx.x == 0) {
        mbar_expect_tx(bar, ROWS * KC * 2 + TN * KC * 2);
        tma_load_2d(sa, &map_a, bar, 0, 0);
        tma_load_2d(sb, &map_b, bar, 0, 0);
        mbar_wait(bar, 0);
        asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
        const uint32_t idesc = (1u << 4) | (1u << 7) | (1u << 10) | ((uint32_t)(TN >> 3) << 17) | ((uint32_t)(128 >> 4) << 24);
        for (int k = 0; k < KC / 16; ++k) {
            const uint32_t a_addr = smem_u32(sa) + shift * KC * 2 + k * 32, b_addr = smem_u32(sb) + k * 32;
            uint64_t da = (uint64_t)((a_addr >> 4) & 0x3FFFu) | ((uint64_t)((SBO >> 4) & 0x3FFFu) << 32) | ((uint64_t)1 << 46) | ((uint64_t)LAYOUT << 61);
            if (use_base_offset) da |= (uint64_t)((a_addr >> 7) & 0x7u) << 49;
            const uint64_t db = (uint64_t)((b_addr >> 4) & 0x3FFFu) | ((uint64_t)((SBO >> 4) & 0x3FFFu) << 32) | ((uint64_t)1 << 46) | ((uint64_t)LAYOUT << 61);
            const uint32_t accum = k != 0;
            asm volatile("{\n.reg .pred p;\nsetp.ne.b32 p, %4, 0;\ntcgen05.mma.cta_group::1.kind::f16 [%0], %1, %2, %3, p;\n}\n" ::"r"(tmem), "l"(da), "l"(db), "r"(idesc), "r"(accum) : "memory");
        }
        asm volatile("tcgen05.commit.cta_group::1.mbarrier::arrive::one.shared::cluster.b64 [%0];" ::"r"(smem_u32(done)) : "memory");
    }
    mbar_wait(done, 0);
    asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
    uint32_t r[32];
    asm volatile(
        "tcgen05.ld.sync.aligned.32x32b.x32.b32 "
        "{%0, %1, %2, %3, %4, %5, %6, %7, %8, %9, %10, %11, %12, %13, %14, %15, "
        "%16, %17, %18, %19, %20, %21, %22, %23, %24, %25, %26, %27, %28, %29, %30, %31}, [%32];"
        : "=r"(r[0]), "=r"(r[1]), "=r"(r[2]), "=r"(r[3]), "=r"(r[4]), "=r"(r[5]), "=r"(r[6]), "=r"(r[7]), "=r"(r[8]), "=r"(r[9]),
          "=r"(r[10]), "=r"(r[11]), "=r"(r[12]), "=r"(r[13]), "=r"(r[14]), "=r"(r[15]), "=r"(r[16]), "=r"(r[17]), "=r"(r[18]),
          "=r"(r[19]), "=r"(r[20]), "=r"(r[21]), "=r"(r[22]), "=r"(r[23]), "=r"(r[24]), "=r"(r[25]), "=r"(r[26]), "=r"(r[27]),
          "=r"(r[28]), "=r"(r[29]), "=r"(r[30]), "=r"(r[31])
        : "r"(tmem + ((uint32_t)(warp * 32) << 16))
        : "memory");
    asm volatile("tcgen05.wait::ld.sync.aligned;" ::: "memory");
    for (int j = 0; j < 32; ++j) out[(warp * 32 + lane) * 32 + j] = __uint_as_float(r[j]);
    asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
    __syncthreads();
    if (warp == 1) asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" ::"r"(tmem), "r"(32) : "memory");
}

typedef CUresult (*PFN_enc)(CUtensorMap*, CUtensorMapDataType, cuuint32_t, void*, const cuuint64_t*, const cuuint64_t*, const cuuint32_t*,
                            const cuuint32_t*, CUtensorMapInterleave, CUtensorMapSwizzle, CUtensorMapL2promotion, CUtensorMapFloatOOBfill);

static int make_map(CUtensorMap* map, const void* base, uint64_t rows, uint64_t cols, uint32_t box_rows, uint32_t box_cols) {
    void* fp = nullptr;
    cudaDriverEntryPointQueryResult q;
    if (cudaGetDriverEntryPoint("cuTensorMapEncodeTiled", &fp, cudaEnableDefault, &q) != cudaSuccess || q != cudaDriverEntryPointSuccess) {
        set_error("cuTensorMapEncodeTiled unavailable");
        return EVFLY_ERR_CUDA;
    }
    cuuint64_t dims[2] = {cols, rows};
    cuuint64_t strides[1] = {cols * 2};
    cuuint32_t box[2] = {box_cols, box_rows};
    cuuint32_t es[2] = {1, 1};
    const CUtensorMapSwizzle sw = (box_cols * 2 == 128) ? CU_TENSOR_MAP_SWIZZLE_128B : CU_TENSOR_MAP_SWIZZLE_64B;
    CUresult r = ((PFN_enc)fp)(map, CU_TENSOR_MAP_DATA_TYPE_BFLOAT16, 2, const_cast<void*>(base), dims, strides, box, es, CU_TENSOR_MAP_INTERLEAVE_NONE, sw,
                               CU_TENSOR_MAP_L2_PROMOTION_NONE, CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
    if (r != CUDA_SUCCESS) {
        set_error("probe: cuTensorMapEncodeTiled failed (%d)", (int)r);
        return EVFLY_ERR_CUDA;
    }
    return EVFLY_OK;
}

}  // namespace probe
}  // namespace evfly

using namespace evfly;

// x bf16 [136, KC], w bf16 [32, KC] -> out fp32 [128, 32] = x[shift : shift+128] @ w^T if the assumption holds
extern "C" int evfly_tc_shift_probe(const void* d_x, const void* d_w, float* d_out, int KC, int shift, int use_base_offset, void* stream) {
    EVFLY_REQUIRE(d_x && d_w && d_out && (KC == 32 || KC == 64) && shift >= 0 && shift <= 8, "tc_shift_probe: bad argument");
    CUtensorMap ma, mb;
    int rc = probe::make_map(&ma, d_x, 136, KC, 136, KC);
    if (rc) return rc;
    rc = probe::make_map(&mb, d_w, 32, KC, 32, KC);
    if (rc) return rc;
    const int smem = 50 * 1024 + 1024;
    if (KC == 64) {
        EVFLY_CUDA(cudaFuncSetAttribute(probe::k_tc_shift_probe<64>, cudaFuncAttributeMaxDynamicSharedMemorySize, smem));
        probe::k_tc_shift_probe<64><<<1, 128, smem, (cudaStream_t)stream>>>(ma, mb, d_out, shift, use_base_offset);
    } else {
        EVFLY_CUDA(cudaFuncSetAttribute(probe::k_tc_shift_probe<32>, cudaFuncAttributeMaxDynamicSharedMemorySize, smem));
        probe::k_tc_shift_probe<32><<<1, 128, smem, (cudaStream_t)stream>>>(ma, mb, d_out, shift, use_base_offset);
    }
    EVFLY_LAUNCHED();
    return EVFLY_OK;
}
