// Shared-memory tile engine of the 2-D SSIM kernels (the metric smoe_ssim and the loss smoe_ssim_loss).
//
// One CTA of 256 threads owns a 32x32 tile of window positions.  The (32+10)^2 input tiles of all channels are staged
// in shared memory once (planar, rows padded to 43 floats), then per channel a horizontal 11-tap pass (each thread
// slides a register window over 8 consecutive outputs of one row) writes 42x32 intermediate rows (padded to 33 floats:
// the stores of the 8 rows a warp works on fall into different banks) and a vertical pass (4 consecutive outputs of one
// column per thread) finishes the window sums.  Every window sum is accumulated in ascending tap order with one fmaf
// per tap, so the values are bit-identical to the separable global-memory passes they replace.
#pragma once
#include "smoe_common.cuh"

namespace smoe {
namespace tile2d {

constexpr int T = 32;             // tile edge (window positions)
constexpr int H = T + 10;         // input rows / columns per tile
constexpr int HP = H + 1;         // padded input row
constexpr int OP = T + 1;         // padded intermediate row
constexpr int SEGH = 8, SEGV = 4;
constexpr int NT = 256;

// exp(-(k-5)^2 / (2 * 1.5^2)) / sum, rounded to float32 (ops/image_ops_impl.py:131-151): a compile-time initialiser
// (no host-to-device copy; capturable)
static __constant__ float c_gauss11[11] = {1.028380124e-03f, 7.598758209e-03f, 3.600077331e-02f, 1.093606874e-01f,
                                           2.130055428e-01f, 2.660117149e-01f, 2.130055428e-01f, 1.093606874e-01f,
                                           3.600077331e-02f, 7.598758209e-03f, 1.028380124e-03f};

__device__ __forceinline__ void cp_async4(float* smem_dst, const float* gsrc) {
    asm volatile("cp.async.ca.shared.global [%0], [%1], 4;" ::"r"((uint32_t)__cvta_generic_to_shared(smem_dst)), "l"(gsrc)
                 : "memory");
}
// wait for this thread's asynchronous copies; the CTA barrier that follows makes everybody's visible
__device__ __forceinline__ void cp_async_wait_all() { asm volatile("cp.async.wait_all;" ::: "memory"); }

__device__ __forceinline__ int reflect_sym(int i, int n) {
    // numpy / tf "SYMMETRIC": -1 -> 0, -2 -> 1, n -> n-1, n+1 -> n-2
    while (i < 0 || i >= n) i = i < 0 ? -i - 1 : 2 * n - 1 - i;
    return i;
}

// dst[c][yy][xx] (planar, row pitch HP) <- src[((y0 - 5 + yy) * pitch + (x0 - 5 + xx)) * C + c] for yy, xx in [0, H);
// positions outside [0, n0) x [0, n1) are SYMMETRIC-reflected (REFLECT) or read as 0.
template <int C, bool REFLECT>
__device__ __forceinline__ void load_tile(float* __restrict__ dst, const float* __restrict__ src, int pitch, int y0,
                                          int x0, int n0, int n1, int tid) {
    constexpr int ROW = H * C;
    const bool interior = y0 >= 5 && x0 >= 5 && y0 + T + 5 <= n0 && x0 + T + 5 <= n1;
    const int lane = tid & 31;
    // one warp per tile row: the row is ROW contiguous floats in global memory (interior tiles), no division per element
    for (int yy = tid >> 5; yy < H; yy += NT / 32) {
        float* drow = dst + yy * HP;
        if (interior) {
            // asynchronous 4-byte copies (LDGSTS): every copy of the tile is in flight before the first one is awaited,
            // and the planar transposition is just the destination address
            const float* srow = src + ((size_t)(y0 - 5 + yy) * pitch + (x0 - 5)) * C;
#pragma unroll
            for (int e0 = 0; e0 < ROW; e0 += 32) {
                const int e = e0 + lane;
                if (e < ROW) {
                    const int xx = e / C, c = e - xx * C;
                    cp_async4(drow + c * H * HP + xx, srow + e);
                }
            }
        } else {
            int gy = y0 - 5 + yy;
            const bool yin = gy >= 0 && gy < n0;
            gy = REFLECT ? reflect_sym(gy, n0) : (yin ? gy : 0);
            const float* srow = src + (size_t)gy * pitch * C;
#pragma unroll
            for (int e0 = 0; e0 < ROW; e0 += 32) {
                const int e = e0 + lane;
                if (e < ROW) {
                    const int xx = e / C, c = e - xx * C;
                    int gx = x0 - 5 + xx;
                    float v = 0.f;
                    if (REFLECT) {
                        gx = reflect_sym(gx, n1);
                        v = srow[gx * C + c];
                    } else if (yin && gx >= 0 && gx < n1) {
                        v = srow[gx * C + c];
                    }
                    drow[c * H * HP + xx] = v;
                }
            }
        }
    }
}

// fixed-order sum of one float per thread over the CTA (shuffle tree per warp, then warp 0 lane order): double result
// valid in thread 0
__device__ __forceinline__ double cta_sum_fixed(float v, double* s_warp /*[NT / 32]*/, int tid) {
    double d = (double)v;
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) d += __shfl_down_sync(0xffffffffu, d, o);
    if ((tid & 31) == 0) s_warp[tid >> 5] = d;
    __syncthreads();
    double t = 0.0;
    if (tid == 0) {
#pragma unroll
        for (int q = 0; q < NT / 32; ++q) t += s_warp[q];
    }
    __syncthreads();
    return t;
}

struct MomentArgs {
    const float* x;          // element (gy, gx, c) of the compute rectangle at x[((size_t)gy * pitch + gx) * C + c]
    const float* y;
    int pitch;               // pixels per buffer row
    int n0, n1;              // compute rectangle: windows are centred on its positions, SYMMETRIC padding at its borders
    int clo0, clo1, cn0, cn1;   // count rectangle (relative to the compute rectangle): positions that enter the sum
    float c1, c2;
    float* maps;             // MAPS: [3][n0 * n1 * C]  d ssim / d (W*x), d / d (W*(x^2+y^2)), d / d (W*(xy)) per position
    double* partial;         // [blocks][4] per-channel SSIM sums of the counted positions
};

// Window moments W*x, W*y, W*(x^2+y^2), W*(xy) of one tile, the SSIM map (ops/image_ops_impl.py:185-233) and its sum;
// with MAPS also the partial derivatives the adjoint pass of the loss needs (ssim_loss.cu).
template <int C, bool MAPS>
__global__ void __launch_bounds__(NT) ssim2d_tile_kernel(MomentArgs g) {
    extern __shared__ float sm[];
    float* ta = sm;                          // [C][H][HP]
    float* tb = ta + C * H * HP;             // [C][H][HP]
    float* hb = tb + C * H * HP;             // [4][H][OP]
    __shared__ double s_warp[NT / 32];
    const int tid = threadIdx.x;
    const int y0 = blockIdx.y * T, x0 = blockIdx.x * T;
    load_tile<C, true>(ta, g.x, g.pitch, y0, x0, g.n0, g.n1, tid);
    load_tile<C, true>(tb, g.y, g.pitch, y0, x0, g.n0, g.n1, tid);
    cp_async_wait_all();
    __syncthreads();
    float w[11];
#pragma unroll
    for (int k = 0; k < 11; ++k) w[k] = c_gauss11[k];
    const size_t total = (size_t)g.n0 * g.n1 * C;
    double* part = g.partial + ((size_t)blockIdx.y * gridDim.x + blockIdx.x) * 4;
    if (tid < 4 && tid >= C) part[tid] = 0.0;
    for (int c = 0; c < C; ++c) {
        // horizontal pass: one item = SEGH consecutive outputs of one row
        if (tid < H * (T / SEGH)) {
            const int yy = tid / (T / SEGH), xs = (tid % (T / SEGH)) * SEGH;
            float h[4][SEGH];
#pragma unroll
            for (int q = 0; q < 4; ++q)
#pragma unroll
                for (int u = 0; u < SEGH; ++u) h[q][u] = 0.f;
            const float* ra = ta + (c * H + yy) * HP + xs;
            const float* rb = tb + (c * H + yy) * HP + xs;
#pragma unroll
            for (int j = 0; j < SEGH + 10; ++j) {
                const float x = ra[j], y = rb[j];
                const float s2 = fmaf(x, x, y * y), xy = x * y;
#pragma unroll
                for (int u = 0; u < SEGH; ++u) {
                    const int k = j - u;
                    if (k >= 0 && k < 11) {
                        h[0][u] = fmaf(w[k], x, h[0][u]);
                        h[1][u] = fmaf(w[k], y, h[1][u]);
                        h[2][u] = fmaf(w[k], s2, h[2][u]);
                        h[3][u] = fmaf(w[k], xy, h[3][u]);
                    }
                }
            }
#pragma unroll
            for (int q = 0; q < 4; ++q)
#pragma unroll
                for (int u = 0; u < SEGH; ++u) hb[(q * H + yy) * OP + xs + u] = h[q][u];
        }
        __syncthreads();
        // vertical pass + SSIM: one item = SEGV consecutive outputs of one column
        float acc = 0.f;
        {
            const int xx = tid & (T - 1), ys = (tid / T) * SEGV;
            float v[4][SEGV];
#pragma unroll
            for (int q = 0; q < 4; ++q)
#pragma unroll
                for (int u = 0; u < SEGV; ++u) v[q][u] = 0.f;
#pragma unroll
            for (int j = 0; j < SEGV + 10; ++j) {
                float in[4];
#pragma unroll
                for (int q = 0; q < 4; ++q) in[q] = hb[(q * H + ys + j) * OP + xx];
#pragma unroll
                for (int u = 0; u < SEGV; ++u) {
                    const int k = j - u;
                    if (k >= 0 && k < 11) {
#pragma unroll
                        for (int q = 0; q < 4; ++q) v[q][u] = fmaf(w[k], in[q], v[q][u]);
                    }
                }
            }
            const int gx = x0 + xx;
#pragma unroll
            for (int u = 0; u < SEGV; ++u) {
                const int gy = y0 + ys + u;
                if (gy < g.n0 && gx < g.n1) {
                    const float mx = v[0][u], my = v[1][u];
                    const float num0 = mx * my * 2.0f;
                    const float den0 = mx * mx + my * my;
                    const float Dl = den0 + g.c1, Dc = v[2][u] - den0 + g.c2;
                    const float lum = (num0 + g.c1) / Dl;
                    const float cs = (v[3][u] * 2.0f - num0 + g.c2) / Dc;
                    const float ssim = lum * cs;
                    if (MAPS) {
                        const size_t i = ((size_t)gy * g.n1 + gx) * C + c;
                        g.maps[i] = 2.f * cs * (my - lum * mx) / Dl + 2.f * lum * (cs * mx - my) / Dc;
                        g.maps[total + i] = -ssim / Dc;
                        g.maps[2 * total + i] = 2.f * lum / Dc;
                    }
                    const bool counted = gy >= g.clo0 && gy < g.clo0 + g.cn0 && gx >= g.clo1 && gx < g.clo1 + g.cn1;
                    if (counted) acc += ssim;
                }
            }
        }
        const double s = cta_sum_fixed(acc, s_warp, tid);      // (its barriers also protect hb for the next channel)
        if (tid == 0) part[c] = s;
    }
}

template <int C, bool MAPS>
inline cudaError_t launch_moments(const MomentArgs& g, cudaStream_t st, int* nblocks) {
    dim3 grid((g.n1 + T - 1) / T, (g.n0 + T - 1) / T);
    const size_t smb = (size_t)(2 * C * H * HP + 4 * H * OP) * sizeof(float);
    cudaError_t e = cudaFuncSetAttribute(ssim2d_tile_kernel<C, MAPS>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smb);
    if (e != cudaSuccess) return e;
    ssim2d_tile_kernel<C, MAPS><<<grid, NT, smb, st>>>(g);
    *nblocks = (int)(grid.x * grid.y);
    return cudaSuccess;
}

}  // namespace tile2d
}  // namespace smoe
