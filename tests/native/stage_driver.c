/* stage_driver.c (TEST-ONLY) -- drives the stage-level C ABI of libevfly_b200.so from plain C, without Python or torch:
 * what the C++ side of evfly_ros would do around evfly_ros/run.py:259-262. Reads a blob written by
 * tests/test_stage_abi_gpu.py (shapes, the packed OrigUNet weights of include/evfly_b200.h evfly_unet_weights in field
 * order, then the frames), runs evfly_unet_forward on cuda:0 and writes depth, y_upconv, hT, cT to the output file.
 *
 *   stage_driver <blob.bin> <out.bin>                                                                              */
#include <cuda_runtime.h>
#include <stdint.h>
#include <stdio.h>
#include <stdlib.h>
#include "../../include/evfly_b200.h"

#define CK(call)                                                                           \
    do {                                                                                   \
        cudaError_t e_ = (call);                                                           \
        if (e_ != cudaSuccess) {                                                           \
            fprintf(stderr, "%s: %s\n", #call, cudaGetErrorString(e_));                    \
            return 2;                                                                      \
        }                                                                                  \
    } while (0)

static void* next_tensor(FILE* f) { /* int64 nbytes + payload -> device pointer */
    int64_t nbytes = 0;
    if (fread(&nbytes, 8, 1, f) != 1 || nbytes <= 0) return NULL;
    void* h = malloc((size_t)nbytes);
    void* d = NULL;
    if (!h || fread(h, 1, (size_t)nbytes, f) != (size_t)nbytes) return NULL;
    if (cudaMalloc(&d, (size_t)nbytes) != cudaSuccess) return NULL;
    if (cudaMemcpy(d, h, (size_t)nbytes, cudaMemcpyHostToDevice) != cudaSuccess) return NULL;
    free(h);
    return d;
}

static int dump(FILE* f, const void* d, size_t nbytes) {
    void* h = malloc(nbytes);
    if (!h || cudaMemcpy(h, d, nbytes, cudaMemcpyDeviceToHost) != cudaSuccess) return 1;
    fwrite(h, 1, nbytes, f);
    free(h);
    return 0;
}

int main(int argc, char** argv) {
    if (argc != 3) return 64;
    FILE* f = fopen(argv[1], "rb");
    if (!f) return 66;
    int32_t hdr[4];
    float cutoff;
    if (fread(hdr, 4, 4, f) != 4 || fread(&cutoff, 4, 1, f) != 1) return 65;
    const int N = hdr[0], n_traj = hdr[1], H = hdr[2], W = hdr[3];
    evfly_unet_weights w;
    w.e11_w = (const float*)next_tensor(f);
    w.e11_b = (const float*)next_tensor(f);
    for (int i = 0; i < 17; ++i) {
        w.conv_w[i] = next_tensor(f);
        w.conv_b[i] = (const float*)next_tensor(f);
    }
    for (int i = 0; i < 4; ++i) {
        w.up_w[i] = next_tensor(f);
        w.up_b[i] = (const float*)next_tensor(f);
    }
    w.out_w = next_tensor(f);
    w.out_b = (const float*)next_tensor(f);
    w.lstm_wx = next_tensor(f);
    w.lstm_wh = next_tensor(f);
    float* frames = (float*)next_tensor(f);
    fclose(f);
    if (!frames || !w.lstm_wh) {
        fprintf(stderr, "short blob\n");
        return 65;
    }
    int h = H, ww = W;
    for (int l = 0; l < 4; ++l) { h = (h - 4) / 2; ww = (ww - 4) / 2; }
    const int vh5 = h - 4, vw5 = ww - 4;
    int vh = vh5, vw = vw5;
    for (int l = 0; l < 4; ++l) { vh = 2 * vh - 4; vw = 2 * vw - 4; }
    const int64_t ws_bytes = evfly_unet_workspace_bytes(N, n_traj, H, W);
    if (ws_bytes <= 0) return 3;
    void* ws;
    float *depth, *yu, *hT, *cT;
    CK(cudaMalloc(&ws, (size_t)ws_bytes));
    CK(cudaMalloc((void**)&depth, (size_t)N * H * W * 4));
    CK(cudaMalloc((void**)&yu, (size_t)N * vh * vw * 4));
    CK(cudaMalloc((void**)&hT, (size_t)n_traj * 512 * vh5 * vw5 * 4));
    CK(cudaMalloc((void**)&cT, (size_t)n_traj * 512 * vh5 * vw5 * 4));
    cudaStream_t st;
    CK(cudaStreamCreate(&st));
    const int64_t l0 = evfly_launch_count();
    int rc = evfly_unet_forward(&w, frames, N, n_traj, H, W, cutoff, NULL, NULL, hT, cT, depth, yu, ws, ws_bytes, st);
    if (rc) {
        fprintf(stderr, "evfly_unet_forward rc=%d: %s\n", rc, evfly_last_error());
        return 4;
    }
    /* second call: the state of the first carried in (run.py keeps origunet_hidden_state between ticks) */
    float *h2, *c2, *depth2;
    CK(cudaMalloc((void**)&h2, (size_t)n_traj * 512 * vh5 * vw5 * 4));
    CK(cudaMalloc((void**)&c2, (size_t)n_traj * 512 * vh5 * vw5 * 4));
    CK(cudaMalloc((void**)&depth2, (size_t)N * H * W * 4));
    rc = evfly_unet_forward(&w, frames, N, n_traj, H, W, cutoff, hT, cT, h2, c2, depth2, yu, ws, ws_bytes, st);
    if (rc) {
        fprintf(stderr, "evfly_unet_forward (carried state) rc=%d: %s\n", rc, evfly_last_error());
        return 4;
    }
    CK(cudaStreamSynchronize(st));
    FILE* o = fopen(argv[2], "wb");
    if (!o) return 73;
    if (dump(o, depth, (size_t)N * H * W * 4) || dump(o, hT, (size_t)n_traj * 512 * vh5 * vw5 * 4) || dump(o, cT, (size_t)n_traj * 512 * vh5 * vw5 * 4) ||
        dump(o, depth2, (size_t)N * H * W * 4))
        return 5;
    fclose(o);
    printf("ok: %lld kernel launches for two forwards of %d frames, workspace %lld bytes\n", (long long)(evfly_launch_count() - l0), N, (long long)ws_bytes);
    return 0;
}
