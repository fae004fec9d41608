// GPU metrics and quantiser (smoe_ssim, smoe_sqerr, smoe_quantize, smoe_rescale, smoe_colminmax,
// smoe_suggest_splits).  HBM-bound pieces; reported in GB/s.
//
// SSIM follows ops/image_ops_impl.py:77-293 as the loss graph calls it (smoe.py:993-1010):
// SYMMETRIC pad by 5 on every domain axis, 11-tap sigma-1.5 Gaussian window (the reference's
// softmax-normalised 11^d window is the outer product of the normalised 1-D windows), VALID
// correlation of x, y, x*x+y*y, x*y, K1=.01, K2=.03, mean over positions per channel.
// 2-D: one shared-memory tile kernel (ssim_tile.cuh).  3-D: separable passes (innermost axis first) over a 4-plane
// workspace; the last pass fuses the SSIM formula and a fixed-order block reduction (no atomics).
#include <math.h>
#include "smoe_common.cuh"
#include "ssim_tile.cuh"

namespace smoe {

// exp(-(k-5)^2 / (2 * 1.5^2)) / sum, rounded to float32 (ops/image_ops_impl.py:131-151): a compile-time initialiser,
// so smoe_ssim needs no host-to-device copy (and can be captured into a CUDA graph)
__constant__ float c_win[11] = {1.028380124e-03f, 7.598758209e-03f, 3.600077331e-02f, 1.093606874e-01f, 2.130055428e-01f,
                                2.660117149e-01f, 2.130055428e-01f, 1.093606874e-01f, 3.600077331e-02f, 7.598758209e-03f,
                                1.028380124e-03f};

__device__ __forceinline__ int reflect_sym(int i, int n) {
    // numpy / tf "SYMMETRIC": -1 -> 0, -2 -> 1, n -> n-1, n+1 -> n-2
    while (i < 0 || i >= n) i = i < 0 ? -i - 1 : 2 * n - 1 - i;
    return i;
}

// planes: 0: E[x], 1: E[y], 2: E[x^2+y^2], 3: E[xy]
template <bool FIRST, bool LAST>
__global__ void __launch_bounds__(256) ssim_pass_kernel(const float* __restrict__ a, const float* __restrict__ b,
                                                        const float* __restrict__ src, float* __restrict__ dst,
                                                        int n0, int n1, int n2, int C, int axis, float c1, float c2,
                                                        double* __restrict__ partial) {
    const size_t total = (size_t)n0 * n1 * n2 * C;
    const size_t i = (size_t)blockIdx.x * 256 + threadIdx.x;
    float ssim = 0.f;
    int ch = 0;
    if (i < total) {
        ch = (int)(i % C);
        size_t pix = i / C;
        int i2 = (int)(pix % n2), i1 = (int)((pix / n2) % n1), i0 = (int)(pix / ((size_t)n2 * n1));
        const int n = axis == 0 ? n0 : (axis == 1 ? n1 : n2);
        const int pos = axis == 0 ? i0 : (axis == 1 ? i1 : i2);
        const size_t stride = axis == 0 ? (size_t)n1 * n2 * C : (axis == 1 ? (size_t)n2 * C : (size_t)C);
        const size_t base = i - (size_t)pos * stride;
        float acc[4] = {0.f, 0.f, 0.f, 0.f};
#pragma unroll
        for (int k = 0; k < 11; ++k) {
            const size_t j = base + (size_t)reflect_sym(pos + k - 5, n) * stride;
            const float wk = c_win[k];
            if (FIRST) {
                const float x = a[j], y = b[j];
                acc[0] = fmaf(wk, x, acc[0]);
                acc[1] = fmaf(wk, y, acc[1]);
                acc[2] = fmaf(wk, fmaf(x, x, y * y), acc[2]);
                acc[3] = fmaf(wk, x * y, acc[3]);
            } else {
#pragma unroll
                for (int q = 0; q < 4; ++q) acc[q] = fmaf(wk, src[(size_t)q * total + j], acc[q]);
            }
        }
        if (LAST) {
            const float num0 = acc[0] * acc[1] * 2.0f;
            const float den0 = acc[0] * acc[0] + acc[1] * acc[1];
            const float lum = (num0 + c1) / (den0 + c1);
            const float num1 = acc[3] * 2.0f;
            const float cs = (num1 - num0 + c2) / (acc[2] - den0 + c2);
            ssim = lum * cs;
        } else {
#pragma unroll
            for (int q = 0; q < 4; ++q) dst[(size_t)q * total + i] = acc[q];
        }
    }
    if (LAST) {
        // per-channel block sums, fixed order: thread 0 walks the 256 slots
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

__global__ void ssim_final_kernel(const double* __restrict__ partial, int nblocks, int C, double inv_n,
                                  double* __restrict__ out) {
    __shared__ double s[256];
    for (int c = 0; c < C; ++c) {
        double acc = 0.0;
        for (int bI = threadIdx.x; bI < nblocks; bI += 256) acc += partial[(size_t)bI * 4 + c];
        s[threadIdx.x] = acc;
        __syncthreads();
        if (threadIdx.x == 0) {
            double t = 0.0;
            for (int q = 0; q < 256; ++q) t += s[q];
            out[c] = t * inv_n;
        }
        __syncthreads();
    }
}

__global__ void __launch_bounds__(256) sqerr_kernel(const float* __restrict__ a, const float* __restrict__ b, size_t n,
                                                    double* __restrict__ partial) {
    // float32 products summed in short float runs, runs and everything above them in double.  16-byte loads, four of
    // them per array in flight per thread (HBM-bound: 2 * n * 4 bytes read).
    double acc = 0.0;
    const size_t stride = (size_t)gridDim.x * 256;
    const size_t tid0 = (size_t)blockIdx.x * 256 + threadIdx.x;
    const bool vec = ((reinterpret_cast<uintptr_t>(a) | reinterpret_cast<uintptr_t>(b)) & 15) == 0;
    const size_t n4 = vec ? n / 4 : 0;
    const float4* a4 = reinterpret_cast<const float4*>(a);
    const float4* b4 = reinterpret_cast<const float4*>(b);
    for (size_t i = tid0; i < n4; i += 4 * stride) {
        float4 x[4], y[4];
#pragma unroll
        for (int u = 0; u < 4; ++u) {
            const size_t j = i + u * stride;
            x[u] = j < n4 ? __ldg(a4 + j) : make_float4(0.f, 0.f, 0.f, 0.f);
            y[u] = j < n4 ? __ldg(b4 + j) : make_float4(0.f, 0.f, 0.f, 0.f);
        }
        float run = 0.f;
#pragma unroll
        for (int u = 0; u < 4; ++u) {
            const float d0 = x[u].x - y[u].x, d1 = x[u].y - y[u].y, d2 = x[u].z - y[u].z, d3 = x[u].w - y[u].w;
            run = fmaf(d0, d0, run); run = fmaf(d1, d1, run); run = fmaf(d2, d2, run); run = fmaf(d3, d3, run);
        }
        acc += (double)run;
    }
    for (size_t i = 4 * n4 + tid0; i < n; i += stride) {           // tail (or everything, when unaligned)
        const float d = a[i] - b[i];
        acc += (double)(d * d);
    }
    __shared__ double s[256];
    s[threadIdx.x] = acc;
    __syncthreads();
    for (int o = 128; o > 0; o >>= 1) {
        if ((int)threadIdx.x < o) s[threadIdx.x] += s[threadIdx.x + o];
        __syncthreads();
    }
    if (threadIdx.x == 0) partial[blockIdx.x] = s[0];
}
__global__ void sqerr_final_kernel(const double* __restrict__ partial, int nb, double* __restrict__ out) {
    // fixed order: lane l sums partials l, l+32, ...; then a fixed shuffle tree
    double t = 0.0;
    for (int i = threadIdx.x; i < nb; i += 32) t += partial[i];
    for (int o = 16; o > 0; o >>= 1) t += __shfl_down_sync(0xffffffffu, t, o);
    if (threadIdx.x == 0) out[0] = t;
}

// ---- quantiser --------------------------------------------------------------------------------
// float32 path: every operation is a separately rounded IEEE op (no FMA contraction), matching
// NumPy: np.round((p - lb) / (ub - lb + 10e-12) * step)
template <typename F>
__global__ void __launch_bounds__(256) quantize_kernel(const float* __restrict__ x, const double* __restrict__ lb,
                                                       const double* __restrict__ ub, size_t n, int cols, double step,
                                                       F* __restrict__ codes);
template <>
__global__ void __launch_bounds__(256) quantize_kernel<float>(const float* __restrict__ x, const double* __restrict__ lb,
                                                              const double* __restrict__ ub, size_t n, int cols,
                                                              double step, float* __restrict__ codes) {
    const size_t i = (size_t)blockIdx.x * 256 + threadIdx.x;
    if (i >= n) return;
    const int c = (int)(i % cols);
    const float l = (float)lb[c], u = (float)ub[c];
    const float den = __fadd_rn(__fsub_rn(u, l), 10e-12f);
    const float nrm = __fdiv_rn(__fsub_rn(x[i], l), den);
    codes[i] = rintf(__fmul_rn(nrm, (float)step));
}
template <>
__global__ void __launch_bounds__(256) quantize_kernel<double>(const float* __restrict__ x, const double* __restrict__ lb,
                                                               const double* __restrict__ ub, size_t n, int cols,
                                                               double step, double* __restrict__ codes) {
    const size_t i = (size_t)blockIdx.x * 256 + threadIdx.x;
    if (i >= n) return;
    const int c = (int)(i % cols);
    const double den = __dadd_rn(__dsub_rn(ub[c], lb[c]), 10e-12);
    const double nrm = __ddiv_rn(__dsub_rn((double)x[i], lb[c]), den);
    codes[i] = rint(__dmul_rn(nrm, step));
}
// r = q / step * (ub - lb) + lb   (quantizer.py:124-130)
__global__ void __launch_bounds__(256) rescale_kernel_f32(const float* __restrict__ q, const double* __restrict__ lb,
                                                          const double* __restrict__ ub, size_t n, int cols, double step,
                                                          float* __restrict__ out) {
    const size_t i = (size_t)blockIdx.x * 256 + threadIdx.x;
    if (i >= n) return;
    const int c = (int)(i % cols);
    const float l = (float)lb[c], u = (float)ub[c];
    out[i] = __fadd_rn(__fmul_rn(__fdiv_rn(q[i], (float)step), __fsub_rn(u, l)), l);
}
__global__ void __launch_bounds__(256) rescale_kernel_f64(const double* __restrict__ q, const double* __restrict__ lb,
                                                          const double* __restrict__ ub, size_t n, int cols, double step,
                                                          double* __restrict__ out) {
    const size_t i = (size_t)blockIdx.x * 256 + threadIdx.x;
    if (i >= n) return;
    const int c = (int)(i % cols);
    out[i] = __dadd_rn(__dmul_rn(__ddiv_rn(q[i], step), __dsub_rn(ub[c], lb[c])), lb[c]);
}
__global__ void __launch_bounds__(256) colminmax_kernel(const float* __restrict__ x, int rows, int cols,
                                                        double* __restrict__ lb, double* __restrict__ ub) {
    const int c = blockIdx.x;
    float mn = INFINITY, mx = -INFINITY;
    for (int r = threadIdx.x; r < rows; r += 256) {
        const float v = x[(size_t)r * cols + c];
        mn = fminf(mn, v);
        mx = fmaxf(mx, v);
    }
    __shared__ float smn[256], smx[256];
    smn[threadIdx.x] = mn;
    smx[threadIdx.x] = mx;
    __syncthreads();
    for (int o = 128; o > 0; o >>= 1) {
        if ((int)threadIdx.x < o) {
            smn[threadIdx.x] = fminf(smn[threadIdx.x], smn[threadIdx.x + o]);
            smx[threadIdx.x] = fmaxf(smx[threadIdx.x], smx[threadIdx.x + o]);
        }
        __syncthreads();
    }
    if (threadIdx.x == 0) { lb[c] = (double)smn[0]; ub[c] = (double)smx[0]; }
}

}  // namespace smoe

using namespace smoe;

extern "C" {

size_t smoe_ssim_workspace_bytes(int d, const int32_t dims[3], int C) {
    (void)d;
    size_t total = (size_t)dims[0] * dims[1] * dims[2] * C;
    size_t nblocks = (total + 255) / 256;       // >= number of 32x32 tiles of the fused 2-D kernel
    return 2 * 4 * total * sizeof(float) + (nblocks + 1024) * 4 * sizeof(double) + 256;
}

int smoe_ssim(int d, const int32_t dims[3], int C, const float* a, const float* b, double* out, void* workspace,
              void* stream) {
    SMOE_REQUIRE(a && b && out && workspace && dims, "null argument");
    SMOE_REQUIRE((d == 2 || d == 3) && C >= 1 && C <= 4, "unsupported d / C");
    cudaStream_t st = (cudaStream_t)stream;
    const int n0 = dims[0], n1 = dims[1], n2 = (d == 3) ? dims[2] : 1;
    const size_t total = (size_t)n0 * n1 * n2 * C;
    const int nblocks = (int)((total + 255) / 256);
    float* p0 = (float*)workspace;
    float* p1 = p0 + 4 * total;
    size_t off = (2 * 4 * total * sizeof(float) + 255) / 256 * 256;
    double* partial = (double*)((char*)workspace + off);
    const float c1 = 0.01f * 0.01f, c2 = 0.03f * 0.03f;
    if (d == 2) {
        // shared-memory tile kernel (ssim_tile.cuh): one launch, HBM traffic = the two images once (+ halo re-reads in L2)
        tile2d::MomentArgs g;
        g.x = a; g.y = b; g.pitch = n1; g.n0 = n0; g.n1 = n1;
        g.clo0 = g.clo1 = 0; g.cn0 = n0; g.cn1 = n1;
        g.c1 = c1; g.c2 = c2; g.maps = nullptr; g.partial = partial;
        int nb = 0;
        cudaError_t e = cudaSuccess;
        switch (C) {
            case 1: e = tile2d::launch_moments<1, false>(g, st, &nb); break;
            case 2: e = tile2d::launch_moments<2, false>(g, st, &nb); break;
            case 3: e = tile2d::launch_moments<3, false>(g, st, &nb); break;
            default: e = tile2d::launch_moments<4, false>(g, st, &nb); break;
        }
        if (e != cudaSuccess) { set_error("smoe_ssim: %s", cudaGetErrorString(e)); return (int)e; }
        ssim_final_kernel<<<1, 256, 0, st>>>(partial, nb, C, 1.0 / (double)((size_t)n0 * n1), out);
        return check_launch("smoe_ssim");
    } else {
        ssim_pass_kernel<true, false><<<nblocks, 256, 0, st>>>(a, b, nullptr, p0, n0, n1, n2, C, 2, c1, c2, nullptr);
        ssim_pass_kernel<false, false><<<nblocks, 256, 0, st>>>(a, b, p0, p1, n0, n1, n2, C, 1, c1, c2, nullptr);
        ssim_pass_kernel<false, true><<<nblocks, 256, 0, st>>>(a, b, p1, nullptr, n0, n1, n2, C, 0, c1, c2, partial);
    }
    ssim_final_kernel<<<1, 256, 0, st>>>(partial, nblocks, C, 1.0 / (double)((size_t)n0 * n1 * n2), out);
    return check_launch("smoe_ssim");
}

int smoe_sqerr(const float* a, const float* b, size_t n, double* out, void* workspace, void* stream) {
    SMOE_REQUIRE(a && b && out && workspace && n > 0, "bad argument");
    int nb = (int)((n / 16 + 255) / 256);       // a thread covers 16 elements per round
    if (nb < 1) nb = 1;
    if (nb > 1024) nb = 1024;          // workspace holds 1024 doubles
    cudaStream_t st = (cudaStream_t)stream;
    sqerr_kernel<<<nb, 256, 0, st>>>(a, b, n, (double*)workspace);
    sqerr_final_kernel<<<1, 32, 0, st>>>((const double*)workspace, nb, out);
    return check_launch("smoe_sqerr");
}

int smoe_quantize(const float* x, const double* lb, const double* ub, int rows, int cols, double step, int f64,
                  void* codes, void* stream) {
    SMOE_REQUIRE(x && lb && ub && codes && rows > 0 && cols > 0, "bad argument");
    size_t n = (size_t)rows * cols;
    int nb = (int)((n + 255) / 256);
    cudaStream_t st = (cudaStream_t)stream;
    if (f64) quantize_kernel<double><<<nb, 256, 0, st>>>(x, lb, ub, n, cols, step, (double*)codes);
    else quantize_kernel<float><<<nb, 256, 0, st>>>(x, lb, ub, n, cols, step, (float*)codes);
    return check_launch("smoe_quantize");
}

int smoe_rescale(const void* codes, const double* lb, const double* ub, int rows, int cols, double step, int f64,
                 void* out, void* stream) {
    SMOE_REQUIRE(codes && lb && ub && out && rows > 0 && cols > 0, "bad argument");
    size_t n = (size_t)rows * cols;
    int nb = (int)((n + 255) / 256);
    cudaStream_t st = (cudaStream_t)stream;
    if (f64) rescale_kernel_f64<<<nb, 256, 0, st>>>((const double*)codes, lb, ub, n, cols, step, (double*)out);
    else rescale_kernel_f32<<<nb, 256, 0, st>>>((const float*)codes, lb, ub, n, cols, step, (float*)out);
    return check_launch("smoe_rescale");
}

int smoe_colminmax(const float* x, int rows, int cols, double* lb, double* ub, void* stream) {
    SMOE_REQUIRE(x && lb && ub && rows > 0 && cols > 0, "bad argument");
    colminmax_kernel<<<cols, 256, 0, (cudaStream_t)stream>>>(x, rows, cols, lb, ub);
    return check_launch("smoe_colminmax");
}

int smoe_suggest_splits(int K_cap, const smoe_batch* b) {
    // Pixel splits of the backward: split s owns tiles s, s+NS, s+2NS, ...  Enough splits that the grid has a
    // few waves of CTAs and a CTA's list stays short (~180 tiles of the whole batch per split measured best on
    // c3), and NS coprime to the tile-grid extents: the kernels of a CTA are spatial neighbours, so a split
    // whose tiles share columns (NS dividing the tiles per row) would give some CTAs all the work and others none.
    int dev = 0, sms = 148;
    if (cudaGetDevice(&dev) == cudaSuccess) cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, dev);
    int nt[3], ntiles = 1;
    for (int i = 0; i < 3; ++i) { nt[i] = (b->extent[i] + b->tile[i] - 1) / b->tile[i]; ntiles *= nt[i]; }
    const int kt = (K_cap + kGroup - 1) / kGroup;
    int want = (2 * 8 * sms + kt - 1) / kt;          // >= ~2 waves of 8 CTAs per SM
    if (want < ntiles / 180) want = ntiles / 180;
    if (want < 8) want = 8;
    if (want >= ntiles) return ntiles < 1 ? 1 : ntiles;      // small problems: one tile per CTA
    for (int ns = want; ns < ntiles; ++ns) {
        bool prime = true;
        for (int q = 2; q * q <= ns; ++q) if (ns % q == 0) { prime = false; break; }
        if (!prime) continue;
        if ((nt[1] > 1 && nt[1] % ns == 0) || (nt[2] > 1 && nt[2] % ns == 0) || (nt[1] * nt[2] > 1 && (nt[1] * nt[2]) % ns == 0)) continue;
        return ns;
    }
    return ntiles;
}

}  // extern "C"
