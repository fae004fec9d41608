// SSIM as the training loss (smoe_ssim_loss).  Replaces the `ssim_opt` branch of the loss graph, smoe.py:981-1010:
//     res, target -> crop the overlap halo -> SYMMETRIC pad 5 -> custom_ssim (ops/image_ops_impl.py:235-293)
//     -> ssim = sum(ssim_c * [6,1,1]) / 8  or  mean_c ;  loss_pixel = 1 - ssim
// and the part of tf.gradients (smoe.py:1148) that runs through it: d loss / d res, through the output
// fake-quant and the clip (straight-through inside [0,1]), written as the per-pixel backward state g_c, gr
// that smoe_backward consumes (the planes smoe_forward fills for the squared-error loss).
//
// With x = res, y = target, the Gaussian window W (11 taps per axis, sigma 1.5) and per output position
//     mx = W*x, my = W*y, e2 = W*(x^2+y^2), exy = W*(xy),
//     lum = (2 mx my + c1) / (mx^2 + my^2 + c1),  cs = (2 exy - 2 mx my + c2) / (e2 - mx^2 - my^2 + c2),
// the derivative of mean(lum * cs) with respect to a pixel is  (P^T a + 2 x P^T b + y P^T c) / Np  with
//     a = d/d mx = 2 cs (my - lum mx) / Dl + 2 lum (cs mx - my) / Dc,   b = d/d e2 = -lum cs / Dc,
//     c = d/d exy = 2 lum / Dc,   Dl = mx^2+my^2+c1,  Dc = e2-mx^2-my^2+c2,
// where P is "symmetric pad, then VALID correlation" and P^T its adjoint.  P is separable, so both P and
// P^T run as one 11-tap pass per axis over planes local to the batch's loss rectangle.  HBM-bound
// elementwise / stencil work; no atomics, fixed-order reductions.
//
// Images (d = 2) take two shared-memory tile kernels instead of the six global-memory passes: ssim2d_tile_kernel
// (ssim_tile.cuh) computes the window moments, the SSIM sum and the maps a, b, c of a 32x32 tile of positions;
// sl_adj_apply_tile applies P^T to a 32x32 tile of pixels out of shared memory and writes the backward state.  Both
// accumulate every sum in the order of the separable passes (kept for video and for rectangles below 16 pixels),
// so the two routes give the same bits (tested).
#include <limits.h>
#include <math.h>
#include <stdlib.h>
#include "smoe_common.cuh"
#include "ssim_tile.cuh"

namespace smoe {

// exp(-(k-5)^2 / (2 * 1.5^2)) / sum, rounded to float32 (ops/image_ops_impl.py:131-151); a compile-time
// initialiser, so that the launches below can be captured into a CUDA graph
__constant__ float c_lwin[11] = {1.028380124e-03f, 7.598758209e-03f, 3.600077331e-02f, 1.093606874e-01f, 2.130055428e-01f, 2.660117149e-01f, 2.130055428e-01f, 1.093606874e-01f, 3.600077331e-02f, 7.598758209e-03f, 1.028380124e-03f};

__device__ __forceinline__ int refl(int i, int n) {
    while (i < 0 || i >= n) i = i < 0 ? -i - 1 : 2 * n - 1 - i;
    return i;
}

struct LRect {
    int dims[3];      // resident image extents
    int lo[3];        // compute rectangle inside the image buffer: SSIM windows are centred on its positions and the
    int n[3];         // symmetric padding reflects at its borders
    int clo[3];       // count rectangle (relative to lo): positions whose SSIM values enter the sum and whose pixels
    int cn[3];        // receive a gradient.  Equal to the compute rectangle unless the caller set a region.
    float inv_np;     // 1 / number of positions of the mean
    int C;
};
__device__ __forceinline__ bool in_count(const LRect& r, const int (&id)[3]) {
    return id[0] >= r.clo[0] && id[0] < r.clo[0] + r.cn[0] && id[1] >= r.clo[1] && id[1] < r.clo[1] + r.cn[1] &&
           id[2] >= r.clo[2] && id[2] < r.clo[2] + r.cn[2];
}

__device__ __forceinline__ size_t rect_global(const LRect& r, int i0, int i1, int i2, int c) {
    return (((size_t)(r.lo[0] + i0) * r.dims[1] + (r.lo[1] + i1)) * r.dims[2] + (r.lo[2] + i2)) * r.C + c;
}

// forward passes of P over planes 0: W*x, 1: W*y, 2: W*(x^2+y^2), 3: W*(xy); the last one turns them into
// the maps a, b, c and reduces the SSIM values per channel
template <bool FIRST, bool LAST>
__global__ void __launch_bounds__(256) sl_fwd_pass(LRect r, const float* __restrict__ x, const float* __restrict__ y,
                                                   const float* __restrict__ src, float* __restrict__ dst, int axis,
                                                   float c1, float c2, double* __restrict__ partial) {
    const size_t total = (size_t)r.n[0] * r.n[1] * r.n[2] * r.C;
    const size_t i = (size_t)blockIdx.x * 256 + threadIdx.x;
    float ssim = 0.f;
    int ch = 0;
    if (i < total) {
        ch = (int)(i % r.C);
        const size_t pix = i / r.C;
        int id[3] = {(int)(pix / ((size_t)r.n[2] * r.n[1])), (int)((pix / r.n[2]) % r.n[1]), (int)(pix % r.n[2])};
        const int n = r.n[axis], pos = id[axis];
        const bool counted = in_count(r, id);
        const size_t stride = axis == 0 ? (size_t)r.n[1] * r.n[2] * r.C : (axis == 1 ? (size_t)r.n[2] * r.C : (size_t)r.C);
        const size_t base = i - (size_t)pos * stride;
        float acc[4] = {0.f, 0.f, 0.f, 0.f};
#pragma unroll
        for (int k = 0; k < 11; ++k) {
            const int q = refl(pos + k - 5, n);
            const float wk = c_lwin[k];
            if (FIRST) {
                id[axis] = q;
                const size_t g = rect_global(r, id[0], id[1], id[2], ch);
                const float xv = x[g], yv = y[g];
                acc[0] = fmaf(wk, xv, acc[0]);
                acc[1] = fmaf(wk, yv, acc[1]);
                acc[2] = fmaf(wk, fmaf(xv, xv, yv * yv), acc[2]);
                acc[3] = fmaf(wk, xv * yv, acc[3]);
            } else {
                const size_t j = base + (size_t)q * stride;
#pragma unroll
                for (int p = 0; p < 4; ++p) acc[p] = fmaf(wk, src[(size_t)p * total + j], acc[p]);
            }
        }
        if (LAST) {
            const float mx = acc[0], my = acc[1];
            const float num0 = mx * my * 2.0f;
            const float den0 = mx * mx + my * my;
            const float Dl = den0 + c1, Dc = acc[2] - den0 + c2;
            const float lum = (num0 + c1) / Dl;
            const float cs = (acc[3] * 2.0f - num0 + c2) / Dc;
            ssim = lum * cs;
            dst[i] = 2.f * cs * (my - lum * mx) / Dl + 2.f * lum * (cs * mx - my) / Dc;
            dst[total + i] = -ssim / Dc;
            dst[2 * total + i] = 2.f * lum / Dc;
            if (!counted) ssim = 0.f;             // a position of the halo ring: another rank's (the sum is per owner)
        } else {
#pragma unroll
            for (int p = 0; p < 4; ++p) dst[(size_t)p * total + i] = acc[p];
        }
    }
    if (LAST) {
        __shared__ float s_v[256];
        __shared__ int s_c[256];
        s_v[threadIdx.x] = (i < total) ? ssim : 0.f;
        s_c[threadIdx.x] = ch;
        __syncthreads();
        if (threadIdx.x < 4) {
            double s = 0.0;
            for (int t = 0; t < 256; ++t)
                if (s_c[t] == (int)threadIdx.x) s += (double)s_v[t];
            partial[(size_t)blockIdx.x * 4 + threadIdx.x] = s;
        }
    }
}

// one axis of P^T over the 3 planes a, b, c:  out[i] = sum over padded positions j' that reflect onto i of
// sum_t w[t] in[j' + t - 5]  (in = 0 outside the rectangle)
__global__ void __launch_bounds__(256) sl_adj_pass(LRect r, const float* __restrict__ src, float* __restrict__ dst,
                                                   int axis) {
    const size_t total = (size_t)r.n[0] * r.n[1] * r.n[2] * r.C;
    const size_t i = (size_t)blockIdx.x * 256 + threadIdx.x;
    if (i >= total) return;
    const size_t pix = i / r.C;
    const int id[3] = {(int)(pix / ((size_t)r.n[2] * r.n[1])), (int)((pix / r.n[2]) % r.n[1]), (int)(pix % r.n[2])};
    const int n = r.n[axis], pos = id[axis];
    const size_t stride = axis == 0 ? (size_t)r.n[1] * r.n[2] * r.C : (axis == 1 ? (size_t)r.n[2] * r.C : (size_t)r.C);
    const size_t base = i - (size_t)pos * stride;
    float acc[3] = {0.f, 0.f, 0.f};
    auto tap = [&](int jp) {
#pragma unroll
        for (int k = 0; k < 11; ++k) {
            const int q = jp + k - 5;
            if (q >= 0 && q < n) {
                const float wk = c_lwin[k];
                const size_t j = base + (size_t)q * stride;
#pragma unroll
                for (int p = 0; p < 3; ++p) acc[p] = fmaf(wk, src[(size_t)p * total + j], acc[p]);
            }
        }
    };
    tap(pos);
    if (pos < 5 || pos >= n - 5) {       // mirror images of this pixel in the symmetric padding
        for (int jp = -5; jp < 0; ++jp)
            if (refl(jp, n) == pos) tap(jp);
        for (int jp = n; jp < n + 5; ++jp)
            if (refl(jp, n) == pos) tap(jp);
    }
#pragma unroll
    for (int p = 0; p < 3; ++p) dst[(size_t)p * total + i] = acc[p];
}

__global__ void sl_final(const double* __restrict__ partial, int nblocks, int C, float* __restrict__ scalars) {
    __shared__ double s[256];
    for (int c = 0; c < C; ++c) {
        double acc = 0.0;
        for (int bI = threadIdx.x; bI < nblocks; bI += 256) acc += partial[(size_t)bI * 4 + c];
        s[threadIdx.x] = acc;
        __syncthreads();
        if (threadIdx.x == 0) {
            double t = 0.0;
            for (int q = 0; q < 256; ++q) t += s[q];
            scalars[8 + c] += (float)t;
        }
        __syncthreads();
    }
}

// d loss / d res -> the g_c and gr planes of the backward state (layout of smoe_forward's epilogue)
template <int D, int C>
__global__ void __launch_bounds__(256) sl_apply(smoe_cfg cfg, smoe_batch b, LRect r, const float* __restrict__ adj,
                                                const float* __restrict__ res, const float* __restrict__ image,
                                                const float* __restrict__ res_pre, float* __restrict__ pix,
                                                int nt1, int nt2, float tau) {
    const size_t total = (size_t)r.n[0] * r.n[1] * r.n[2];            // positions of the compute rectangle (plane size)
    const size_t ncount = (size_t)r.cn[0] * r.cn[1] * r.cn[2];
    const size_t iq = (size_t)blockIdx.x * 256 + threadIdx.x;
    if (iq >= ncount) return;
    // position inside the count rectangle -> inside the compute rectangle (the planes' index space)
    const int i2 = r.clo[2] + (int)(iq % r.cn[2]), i1 = r.clo[1] + (int)((iq / r.cn[2]) % r.cn[1]),
              i0 = r.clo[0] + (int)(iq / ((size_t)r.cn[2] * r.cn[1]));
    const size_t ip = ((size_t)i0 * r.n[1] + i1) * r.n[2] + i2;
    // position inside the batch (forward) rectangle -> tile and slot
    const int f[3] = {r.lo[0] + i0 - b.origin[0], r.lo[1] + i1 - b.origin[1], r.lo[2] + i2 - b.origin[2]};
    const int t0 = f[0] / b.tile[0], t1 = f[1] / b.tile[1], t2 = f[2] / b.tile[2];
    const int tile = (t0 * nt1 + t1) * nt2 + t2;
    const int j = ((f[0] % b.tile[0]) * b.tile[1] + (f[1] % b.tile[1])) * b.tile[2] + (f[2] % b.tile[2]);
    float* tp = pix + (size_t)tile * pix_stride(D, C, b.tile[D - 1]);
    const size_t nC = total * C;
    const float inv_np = r.inv_np;
    float gr = 0.f;
#pragma unroll
    for (int c = 0; c < C; ++c) {
        const size_t g = rect_global(r, i0, i1, i2, c);
        const size_t l = ip * C + c;
        const float x = res[g], y = image[g], rv = res_pre[g];
        const float cw = cfg.use_yuv ? (C == 3 ? (c == 0 ? 0.75f : 0.125f) : 1.0f) : (1.0f / C);   // smoe.py:1006-1009
        const float dS = fmaf(y, adj[2 * nC + l], fmaf(2.f * x, adj[nC + l], adj[l])) * inv_np;
        const bool ste = (rv >= 0.f) && (rv <= 1.f);
        const float gc = ste ? -cw * dS : 0.f;                    // loss_pixel = 1 - ssim
        gr = fmaf(gc, rv, gr);
        tp[(PL_G + c) * SMOE_TPIX + j] = gc;
    }
    // S > 1e-11 (smoe.py:821): the forward stored log2f(max(S, floor)), same device log2f here
    const bool live = tp[PL_QTHR * SMOE_TPIX + j] > log2f(kSFloor);
    tp[PL_GR * SMOE_TPIX + j] = live ? gr : 0.f;
}

// P^T of the maps a, b, c and d loss / d res -> the g_c / gr planes, for one 32x32 tile of pixels of the count
// rectangle.  The maps of the tile plus a ring of 5 are staged in shared memory (0 outside the compute rectangle); a
// pixel within 5 of a border of the compute rectangle also collects the windows centred on its mirror image in the
// symmetric padding (the one position jp outside [0, n) with reflect(jp) = pos; n >= 16 makes it unique).
struct AdjArgs {
    smoe_cfg cfg;
    smoe_batch b;
    LRect r;
    const float* maps;        // [3][n0 * n1 * C]
    const float* res;
    const float* image;
    const float* res_pre;
    float* pix;
    int nt1, nt2;
};

template <int D, int C>
__global__ void __launch_bounds__(tile2d::NT) sl_adj_apply_tile(AdjArgs g) {
    using namespace tile2d;
    extern __shared__ float sm[];
    float* tm = sm;                           // [3][C][H][HP]   maps of the tile + ring
    float* hb = tm + 3 * C * H * HP;          // [3][T][HP]      after the pass along y (same axis order as sl_adj_pass)
    const int tid = threadIdx.x;
    const LRect& r = g.r;
    const int n0 = r.n[0], n1 = r.n[1];
    const int y0 = r.clo[0] + blockIdx.y * T, x0 = r.clo[1] + blockIdx.x * T;     // in compute-rectangle coordinates
    const size_t total = (size_t)n0 * n1 * C;
#pragma unroll
    for (int p = 0; p < 3; ++p) load_tile<C, false>(tm + p * C * H * HP, g.maps + p * total, n1, y0, x0, n0, n1, tid);
    cp_async_wait_all();
    __syncthreads();
    float w[11];
#pragma unroll
    for (int k = 0; k < 11; ++k) w[k] = c_gauss11[k];
    // the mirror image of position pos (if it has one): centre of the extra windows
    auto mirror = [](int pos, int n) { return pos < 5 ? -pos - 1 : (pos >= n - 5 ? 2 * n - 1 - pos : INT_MIN); };
    const int prow = tid / (T / SEGV), pxs = (tid % (T / SEGV)) * SEGV;     // this thread's pixels: SEGV along x
    // this thread's pixels: slot in the backward state and index in the resident buffers, -1 outside the count rectangle
    float gr[SEGV];
    int slot[SEGV], gidx[SEGV];
    {
        const int gy = y0 + prow;
        const int f0 = r.lo[0] + gy - g.b.origin[0];
        const int pstr = pix_stride(D, C, g.b.tile[D - 1]);
#pragma unroll
        for (int u = 0; u < SEGV; ++u) {
            gr[u] = 0.f;
            const int gx = x0 + pxs + u;
            const bool in = gy < r.clo[0] + r.cn[0] && gx < r.clo[1] + r.cn[1];
            const int f1 = r.lo[1] + gx - g.b.origin[1];
            const int tile = ((f0 / g.b.tile[0]) * g.nt1 + f1 / g.b.tile[1]) * g.nt2;
            const int j = ((f0 % g.b.tile[0]) * g.b.tile[1] + (f1 % g.b.tile[1])) * g.b.tile[2];
            slot[u] = in ? tile * pstr + j : -1;
            gidx[u] = ((r.lo[0] + gy) * r.dims[1] + (r.lo[1] + gx)) * C;
        }
    }
    for (int c = 0; c < C; ++c) {
        // adjoint along y: SEGH consecutive rows of one column (of the tile + ring) per thread
        if (tid < H * (T / SEGH)) {
            const int xx = tid % H, ys = (tid / H) * SEGH;
            float h[3][SEGH];
#pragma unroll
            for (int p = 0; p < 3; ++p)
#pragma unroll
                for (int u = 0; u < SEGH; ++u) h[p][u] = 0.f;
            const float* col = tm + c * H * HP + xx;
#pragma unroll
            for (int j = 0; j < SEGH + 10; ++j) {
                float in[3];
#pragma unroll
                for (int p = 0; p < 3; ++p) in[p] = col[p * C * H * HP + (ys + j) * HP];
#pragma unroll
                for (int u = 0; u < SEGH; ++u) {
                    const int k = j - u;
                    if (k >= 0 && k < 11) {
#pragma unroll
                        for (int p = 0; p < 3; ++p) h[p][u] = fmaf(w[k], in[p], h[p][u]);
                    }
                }
            }
            if (y0 + ys < 5 || y0 + ys + SEGH > n0 - 5) {          // some row of the segment has a mirror image
#pragma unroll
                for (int u = 0; u < SEGH; ++u) {
                    const int gy = y0 + ys + u;
                    const int jp = gy < n0 ? mirror(gy, n0) : INT_MIN;
                    if (jp == INT_MIN) continue;
                    for (int k = 0; k < 11; ++k) {
                        const int q = jp + k - 5;
                        if (q >= 0 && q < n0) {
#pragma unroll
                            for (int p = 0; p < 3; ++p)
                                h[p][u] = fmaf(c_gauss11[k], col[p * C * H * HP + (q - (y0 - 5)) * HP], h[p][u]);
                        }
                    }
                }
            }
#pragma unroll
            for (int p = 0; p < 3; ++p)
#pragma unroll
                for (int u = 0; u < SEGH; ++u) hb[(p * T + ys + u) * HP + xx] = h[p][u];
        }
        __syncthreads();
        // adjoint along x: SEGV consecutive pixels of one row per thread, then the pixels' gradient
        {
            float v[3][SEGV];
#pragma unroll
            for (int p = 0; p < 3; ++p)
#pragma unroll
                for (int u = 0; u < SEGV; ++u) v[p][u] = 0.f;
            const float* row = hb + prow * HP;
#pragma unroll
            for (int j = 0; j < SEGV + 10; ++j) {
                float in[3];
#pragma unroll
                for (int p = 0; p < 3; ++p) in[p] = row[p * T * HP + pxs + j];
#pragma unroll
                for (int u = 0; u < SEGV; ++u) {
                    const int k = j - u;
                    if (k >= 0 && k < 11) {
#pragma unroll
                        for (int p = 0; p < 3; ++p) v[p][u] = fmaf(w[k], in[p], v[p][u]);
                    }
                }
            }
            if (x0 + pxs < 5 || x0 + pxs + SEGV > n1 - 5) {
#pragma unroll
                for (int u = 0; u < SEGV; ++u) {
                    const int gx = x0 + pxs + u;
                    const int jp = gx < n1 ? mirror(gx, n1) : INT_MIN;
                    if (jp == INT_MIN) continue;
                    for (int k = 0; k < 11; ++k) {
                        const int q = jp + k - 5;
                        if (q >= 0 && q < n1) {
#pragma unroll
                            for (int p = 0; p < 3; ++p)
                                v[p][u] = fmaf(c_gauss11[k], row[p * T * HP + q - (x0 - 5)], v[p][u]);
                        }
                    }
                }
            }
#pragma unroll
            for (int u = 0; u < SEGV; ++u) {
                if (slot[u] < 0) continue;
                const int gi = gidx[u] + c;
                const float x = g.res[gi], y = g.image[gi], rv = g.res_pre[gi];
                const float cw = g.cfg.use_yuv ? (C == 3 ? (c == 0 ? 0.75f : 0.125f) : 1.0f) : (1.0f / C);   // smoe.py:1006-1009
                const float dS = fmaf(y, v[2][u], fmaf(2.f * x, v[1][u], v[0][u])) * r.inv_np;
                const bool ste = (rv >= 0.f) && (rv <= 1.f);
                const float gc = ste ? -cw * dS : 0.f;                    // loss_pixel = 1 - ssim
                gr[u] = fmaf(gc, rv, gr[u]);
                g.pix[(size_t)slot[u] + (PL_G + c) * SMOE_TPIX] = gc;
            }
        }
        __syncthreads();
    }
#pragma unroll
    for (int u = 0; u < SEGV; ++u) {
        if (slot[u] < 0) continue;
        float* tp = g.pix + (size_t)slot[u];
        // S > 1e-11 (smoe.py:821): the forward stored log2f(max(S, floor)), same device log2f here
        const bool live = tp[PL_QTHR * SMOE_TPIX] > log2f(kSFloor);
        tp[PL_GR * SMOE_TPIX] = live ? gr[u] : 0.f;
    }
}

static LRect loss_rect(const smoe_cfg* cfg, const smoe_batch* b, const smoe_ssim_region* reg) {
    LRect r;
    for (int a = 0; a < 3; ++a) {
        r.dims[a] = b->dims[a];
        int lo = b->origin[a], hi = b->origin[a] + b->extent[a];
        if (a < cfg->d && b->halo > 0) {          // the halo is cropped on every side that is not the image border
            if (lo > 0) lo += b->halo;
            if (hi < b->dims[a]) hi -= b->halo;
        }
        r.lo[a] = lo;
        r.n[a] = hi - lo;
        r.clo[a] = 0;
        r.cn[a] = r.n[a];
        if (reg) {                                 // compute rectangle given by the caller; the batch is the count rectangle
            r.clo[a] = lo - reg->lo[a];
            r.lo[a] = reg->lo[a];
            r.n[a] = reg->n[a];
        }
    }
    r.inv_np = reg ? reg->inv_count : 1.0f / (float)((size_t)r.n[0] * r.n[1] * r.n[2]);
    r.C = cfg->C;
    return r;
}

}  // namespace smoe

using namespace smoe;

extern "C" size_t smoe_ssim_loss_workspace_bytes(const smoe_cfg* cfg, const smoe_batch* batch) {
    if (!cfg || !batch) return 0;
    // sized for the whole resident buffer, so that a caller-given compute region (batch + halo ring) fits too
    const size_t total = (size_t)batch->dims[0] * batch->dims[1] * batch->dims[2] * cfg->C;
    const size_t nblocks = (total + 255) / 256;
    return 2 * 4 * total * sizeof(float) + 256 + nblocks * 4 * sizeof(double) + 256;
}

extern "C" int smoe_ssim_loss(const smoe_cfg* cfg, const smoe_batch* batch, const smoe_ssim_region* region,
                              const float* res, const float* image, const float* res_pre, float* pix, float* scalars,
                              void* workspace, void* stream) {
    SMOE_REQUIRE(cfg && batch && res && image && res_pre && scalars && workspace, "null argument");
    SMOE_REQUIRE(cfg->d == 2 || cfg->d == 3, "unsupported d");
    SMOE_REQUIRE(!region || batch->halo == 0, "a compute region and an overlap halo cannot be combined");
    const LRect r = loss_rect(cfg, batch, region);
    for (int a = 0; a < 3; ++a) {
        SMOE_REQUIRE(r.n[a] > 0 && r.cn[a] > 0, "overlap halo swallows the batch");
        SMOE_REQUIRE(r.lo[a] >= 0 && r.lo[a] + r.n[a] <= r.dims[a] && r.clo[a] >= 0 && r.clo[a] + r.cn[a] <= r.n[a],
                     "compute region must lie inside the buffer and contain the batch");
    }
    cudaStream_t st = (cudaStream_t)stream;
    const size_t npos = (size_t)r.n[0] * r.n[1] * r.n[2];
    const size_t total = npos * r.C;
    const size_t cap = (size_t)batch->dims[0] * batch->dims[1] * batch->dims[2] * cfg->C;
    const int nblocks = (int)((total + 255) / 256);
    float* p0 = (float*)workspace;
    float* p1 = p0 + 4 * cap;
    double* partial = (double*)((char*)workspace + (2 * 4 * cap * sizeof(float) + 255) / 256 * 256);
    const float c1 = 0.01f * 0.01f, c2 = 0.03f * 0.03f;
    const int nt1 = (batch->extent[1] + batch->tile[1] - 1) / batch->tile[1];
    const int nt2 = (batch->extent[2] + batch->tile[2] - 1) / batch->tile[2];
    // images: two shared-memory tile kernels; video, tiny rectangles and SMOE_SSIM_GENERIC=1: separable global passes
    const bool tiled = cfg->d == 2 && r.n[2] == 1 && r.n[0] >= 16 && r.n[1] >= 16 && !getenv("SMOE_SSIM_GENERIC");
    if (tiled) {
        tile2d::MomentArgs g;
        const size_t org = ((size_t)r.lo[0] * r.dims[1] + r.lo[1]) * r.C;
        g.x = res + org; g.y = image + org; g.pitch = r.dims[1]; g.n0 = r.n[0]; g.n1 = r.n[1];
        g.clo0 = r.clo[0]; g.clo1 = r.clo[1]; g.cn0 = r.cn[0]; g.cn1 = r.cn[1];
        g.c1 = c1; g.c2 = c2; g.maps = p0; g.partial = partial;
        int nb = 0;
        cudaError_t e = cfg->C == 1 ? (pix ? tile2d::launch_moments<1, true>(g, st, &nb) : tile2d::launch_moments<1, false>(g, st, &nb))
                                    : (pix ? tile2d::launch_moments<3, true>(g, st, &nb) : tile2d::launch_moments<3, false>(g, st, &nb));
        if (e != cudaSuccess) { set_error("smoe_ssim_loss: %s", cudaGetErrorString(e)); return (int)e; }
        sl_final<<<1, 256, 0, st>>>(partial, nb, r.C, scalars);
        if (pix) {
            AdjArgs a;
            a.cfg = *cfg; a.b = *batch; a.r = r; a.maps = p0; a.res = res; a.image = image; a.res_pre = res_pre;
            a.pix = pix; a.nt1 = nt1; a.nt2 = nt2;
            dim3 grid((r.cn[1] + tile2d::T - 1) / tile2d::T, (r.cn[0] + tile2d::T - 1) / tile2d::T);
            const size_t smb = (size_t)(3 * r.C * tile2d::H * tile2d::HP + 3 * tile2d::T * tile2d::HP) * sizeof(float);
            if (cfg->C == 1) {
                e = cudaFuncSetAttribute(sl_adj_apply_tile<2, 1>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smb);
                if (e == cudaSuccess) sl_adj_apply_tile<2, 1><<<grid, tile2d::NT, smb, st>>>(a);
            } else {
                e = cudaFuncSetAttribute(sl_adj_apply_tile<2, 3>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smb);
                if (e == cudaSuccess) sl_adj_apply_tile<2, 3><<<grid, tile2d::NT, smb, st>>>(a);
            }
            if (e != cudaSuccess) { set_error("smoe_ssim_loss: %s", cudaGetErrorString(e)); return (int)e; }
        }
        return check_launch("smoe_ssim_loss");
    }
    float* src = nullptr;
    float* dst = p0;
    for (int axis = cfg->d - 1; axis >= 0; --axis) {
        const bool first = axis == cfg->d - 1, last = axis == 0;
        if (first)
            sl_fwd_pass<true, false><<<nblocks, 256, 0, st>>>(r, res, image, nullptr, dst, axis, c1, c2, nullptr);
        else if (!last)
            sl_fwd_pass<false, false><<<nblocks, 256, 0, st>>>(r, res, image, src, dst, axis, c1, c2, nullptr);
        else
            sl_fwd_pass<false, true><<<nblocks, 256, 0, st>>>(r, res, image, src, dst, axis, c1, c2, partial);
        src = dst;
        dst = (dst == p0) ? p1 : p0;
    }
    sl_final<<<1, 256, 0, st>>>(partial, nblocks, r.C, scalars);
    if (pix) {
        for (int axis = 0; axis < cfg->d; ++axis) {
            sl_adj_pass<<<nblocks, 256, 0, st>>>(r, src, dst, axis);
            src = dst;
            dst = (dst == p0) ? p1 : p0;
        }
        const float tau = 0.5f / (float)(1 << cfg->precision);
        const int nb = (int)(((size_t)r.cn[0] * r.cn[1] * r.cn[2] + 255) / 256);
#define CALL(D, C) sl_apply<D, C><<<nb, 256, 0, st>>>(*cfg, *batch, r, src, res, image, res_pre, pix, nt1, nt2, tau);
        SMOE_DISPATCH_DC(cfg->d, cfg->C, CALL)
#undef CALL
    }
    return check_launch("smoe_ssim_loss");
}
