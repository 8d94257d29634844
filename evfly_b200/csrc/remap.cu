// remap.cu -- "next" row N3: rectification of event frames / depth images between accumulation and crop
// (utils/calibration_tools/rectify_bag.py:91-138 remap_img / remap_events, evfly_ros/run.py:339-340).
// The reference calls cv2.remap(img, mapx, mapy, INTER_CUBIC) with the default constant-0 border; the kernel
// reproduces OpenCV's arithmetic bit for bit (imgproc/src/imgwarp.cpp: map coordinates rounded half-even to 1/32
// pixel, 2-D weights = float products of two rows of the A = -0.75 cubic table, interior pixels summed row by row,
// border pixels tap by tap from zero, no fused multiply-add). One thread per output pixel, the maps are shared by
// the N images of a batch, and the uint8 -> (v - 128) * 0.2 decode of run.py:334-336 can be applied on the fly so
// that the accumulator's byte image is the only thing read.
#include "common.cuh"

namespace evfly {

// OpenCV's bicubic coefficient table travels as a kernel argument (constant bank): a __constant__ symbol would exist
// once per device and need an upload per device, which is illegal inside a CUDA-graph capture
struct CubicTab { float v[32 * 4]; };

static void cubic_table_host(float* tab) {
    const float A = -0.75f;
    for (int i = 0; i < 32; ++i) {
        const float x = (float)i * (1.0f / 32), x1 = x + 1.f, y = 1.f - x;
        tab[i * 4 + 0] = ((A * x1 - 5 * A) * x1 + 8 * A) * x1 - 4 * A;       // every intermediate is exact in fp32
        tab[i * 4 + 1] = ((A + 2) * x - (A + 3)) * x * x + 1;
        tab[i * 4 + 2] = ((A + 2) * y - (A + 3)) * y * y + 1;
        tab[i * 4 + 3] = 1.f - tab[i * 4 + 0] - tab[i * 4 + 1] - tab[i * 4 + 2];
    }
}

template <bool U8>
__device__ __forceinline__ float src_at(const void* img, int W, int y, int x, int flip) {
    const int xs = flip ? W - 1 - x : x;                                     // remap_img: img[:, ::-1] before the remap
    if (U8) {
        const float v = (float)reinterpret_cast<const unsigned char*>(img)[(size_t)y * W + xs];
        return __fmul_rn(__fsub_rn(v, 128.f), 0.2f);                        // run.py:334-336 in float32
    }
    return reinterpret_cast<const float*>(img)[(size_t)y * W + xs];
}

template <bool U8>
__global__ void __launch_bounds__(256)
k_remap_bicubic(const void* __restrict__ src, int N, int H, int W, const float* __restrict__ mapx, const float* __restrict__ mapy,
                long long map_ld, int OH, int OW, int flip, int rotate, float* __restrict__ dst, const __grid_constant__ CubicTab tab) {
    const long long total = (long long)N * OH * OW;
    const long long stride = (long long)gridDim.x * blockDim.x;
    for (long long idx = (long long)blockIdx.x * blockDim.x + threadIdx.x; idx < total; idx += stride) {
        const int j = (int)(idx % OW), i = (int)((idx / OW) % OH);
        const long long n = idx / ((long long)OW * OH);
        const void* img = U8 ? (const void*)(reinterpret_cast<const unsigned char*>(src) + n * (long long)H * W)
                             : (const void*)(reinterpret_cast<const float*>(src) + n * (long long)H * W);
        const int sx = __float2int_rn(__fmul_rn(mapx[i * map_ld + j], 32.f));   // cvRound(x * INTER_TAB_SIZE)
        const int sy = __float2int_rn(__fmul_rn(mapy[i * map_ld + j], 32.f));
        const float* wx = tab.v + (sx & 31) * 4;
        const float* wy = tab.v + (sy & 31) * 4;
        const int ix = min(max(sx >> 5, -32768), 32767) - 1;                 // saturate_cast<short>, then the 4x4 window starts one left/up
        const int iy = min(max(sy >> 5, -32768), 32767) - 1;
        float sum;
        if ((unsigned)ix < (unsigned)max(W - 3, 0) && (unsigned)iy < (unsigned)max(H - 3, 0)) {
            sum = 0.f;
#pragma unroll
            for (int k1 = 0; k1 < 4; ++k1) {
                float row = __fmul_rn(src_at<U8>(img, W, iy + k1, ix, flip), __fmul_rn(wy[k1], wx[0]));
#pragma unroll
                for (int k2 = 1; k2 < 4; ++k2)
                    row = __fadd_rn(row, __fmul_rn(src_at<U8>(img, W, iy + k1, ix + k2, flip), __fmul_rn(wy[k1], wx[k2])));
                sum = k1 == 0 ? row : __fadd_rn(sum, row);
            }
        } else {
            sum = 0.f;                                                        // also the value when the window misses the image
#pragma unroll
            for (int k1 = 0; k1 < 4; ++k1) {
                const int yy = iy + k1;
                if (yy < 0 || yy >= H) continue;
#pragma unroll
                for (int k2 = 0; k2 < 4; ++k2) {
                    const int xx = ix + k2;
                    if (xx >= 0 && xx < W) sum = __fadd_rn(sum, __fmul_rn(src_at<U8>(img, W, yy, xx, flip), __fmul_rn(wy[k1], wx[k2])));
                }
            }
        }
        const int oi = rotate ? OH - 1 - i : i, oj = rotate ? OW - 1 - j : j;   // cv2.rotate(ROTATE_180)
        dst[(n * OH + oi) * OW + oj] = sum;
    }
}

// remap_events (rectify_bag.py:101-116): x' = mapx[y, x], y' = mapy[y, x], optional 180-degree rotation, in-frame mask
__global__ void __launch_bounds__(256)
k_remap_events(const int* __restrict__ x, const int* __restrict__ y, long long n, const float* __restrict__ mapx,
               const float* __restrict__ mapy, int H, int W, int rotate, int tw, int th, float* __restrict__ ox,
               float* __restrict__ oy, unsigned char* __restrict__ keep) {
    const long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= n) return;
    const int xi = x[i], yi = y[i];
    float fx = 0.f, fy = 0.f;
    bool ok = (unsigned)xi < (unsigned)W && (unsigned)yi < (unsigned)H;
    if (ok) {
        fx = mapx[(size_t)yi * W + xi];
        fy = mapy[(size_t)yi * W + xi];
        if (rotate) {
            fx = __fsub_rn((float)(tw - 1), fx);
            fy = __fsub_rn((float)(th - 1), fy);
        }
        ok = fx >= 0.f && fx <= (float)(tw - 1) && fy >= 0.f && fy <= (float)(th - 1);
    }
    ox[i] = fx;
    oy[i] = fy;
    keep[i] = ok ? 1 : 0;
}

}  // namespace evfly

using namespace evfly;

extern "C" int evfly_remap_bicubic_f32(const void* d_src, int src_is_u8, int N, int H, int W, const float* d_mapx, const float* d_mapy,
                                       int64_t map_ld, int OH, int OW, int flip, int rotate, float* d_dst, void* stream) {
    EVFLY_REQUIRE(d_src && d_mapx && d_mapy && d_dst && N >= 0 && H > 0 && W > 0 && OH > 0 && OW > 0 && map_ld >= OW, "remap_bicubic_f32: bad argument");
    if (N == 0) return EVFLY_OK;
    static const CubicTab tab = [] { CubicTab t; cubic_table_host(t.v); return t; }();     // host-side, thread-safe (C++11 static init)
    const long long total = (long long)N * OH * OW;
    const int grid = stream_grid(total, 256, 16);
    cudaStream_t st = (cudaStream_t)stream;
    if (src_is_u8)
        k_remap_bicubic<true><<<grid, 256, 0, st>>>(d_src, N, H, W, d_mapx, d_mapy, map_ld, OH, OW, flip, rotate, d_dst, tab);
    else
        k_remap_bicubic<false><<<grid, 256, 0, st>>>(d_src, N, H, W, d_mapx, d_mapy, map_ld, OH, OW, flip, rotate, d_dst, tab);
    EVFLY_LAUNCHED();
    return EVFLY_OK;
}

extern "C" int evfly_remap_events_f32(const int32_t* d_x, const int32_t* d_y, int64_t n, const float* d_mapx, const float* d_mapy, int H,
                                      int W, int rotate, int target_w, int target_h, float* d_out_x, float* d_out_y, uint8_t* d_keep,
                                      void* stream) {
    EVFLY_REQUIRE(n >= 0 && H > 0 && W > 0 && d_mapx && d_mapy && (n == 0 || (d_x && d_y && d_out_x && d_out_y && d_keep)), "remap_events_f32: bad argument");
    if (n == 0) return EVFLY_OK;
    k_remap_events<<<(unsigned)ceil_div(n, 256), 256, 0, (cudaStream_t)stream>>>(d_x, d_y, n, d_mapx, d_mapy, H, W, rotate, target_w, target_h,
                                                                                d_out_x, d_out_y, d_keep);
    EVFLY_LAUNCHED();
    return EVFLY_OK;
}
