// pi-mask stream compaction + parameter staging (smoe_pack, smoe_pack_fed, smoe_update_kernel_list).
//
// Replaces smoe.py:474-480 (fake-quant of pis), 732-735 (A assembly), 738-753 (bool_mask,
// indices, 5x boolean_mask) and 1012 (num_pi).  HBM-bound: reads K_all*(P*4+1) bytes, writes
// K*(PK*4+4) bytes.  Two launches: (1) per-block flag counts and regulariser partial sums,
// (2) every block re-derives its exclusive prefix from the <= few-thousand block counts,
// scans its own flags with warp ballots and scatters records in the order of `perm` (a stable
// compaction of the permuted sequence; perm == NULL is ascending kernel index, the order of numpy
// boolean masking).  Smoe passes the Hilbert order of the centres, so that 128 consecutive records
// (a forward chunk) and 64 consecutive records (a backward CTA) are spatial neighbours; the SET of
// surviving indices is what the reference defines and does not depend on the order.  No atomics:
// sums are fixed-order.
#include <math.h>
#include <stdarg.h>
#include "smoe_common.cuh"

namespace smoe {

static thread_local char g_err[512] = "";
void set_error(const char* fmt, ...) {
    va_list ap;
    va_start(ap, fmt);
    vsnprintf(g_err, sizeof(g_err), fmt, ap);
    va_end(ap);
}
int check_launch(const char* what) {
    cudaError_t e = cudaGetLastError();
    if (e != cudaSuccess) {
        set_error("%s: %s", what, cudaGetErrorString(e));
        return (int)e;
    }
    return 0;
}

struct PackBlk { int32_t count, numpi, nonpos, pad; float sum_pi, sum_diag; };

template <int D, int C>
__global__ void __launch_bounds__(256) pack_count_kernel(const float* __restrict__ theta,
                                                         const uint8_t* __restrict__ klist,
                                                         const int32_t* __restrict__ perm, int K_all,
                                                         int quantize_pis, QuantSet qs_in,
                                                         const QuantDyn* __restrict__ qdyn, PackBlk* __restrict__ blk,
                                                         float* __restrict__ grads_clear,
                                                         float* __restrict__ scalars_clear,
                                                         uint8_t* __restrict__ infl_clear) {
    const QuantSet qs = qs_in.mode == 3 ? qdyn->qs : qs_in;
    constexpr int P = nparam(D, C);
    const int j = blockIdx.x * 256 + threadIdx.x;
    // start-of-pass clears folded into this launch (zero_op of smoe.py:1612-1613, scalar block, influence flags)
    if (grads_clear) {
        const size_t n = (size_t)K_all * P;
        for (size_t q = (size_t)j; q < n; q += (size_t)gridDim.x * 256) grads_clear[q] = 0.f;
    }
    if (scalars_clear && j < SMOE_NSCAL) scalars_clear[j] = 0.f;
    if (infl_clear && j < K_all) infl_clear[j] = 0;
    const int i = j < K_all ? (perm ? perm[j] : j) : K_all;
    int flag = 0, numpi = 0;
    float spi = 0.f, sdiag = 0.f;
    if (i < K_all) {
        const float* row = theta + (size_t)i * P;
        float pi = row[off_pi(D, C)];
        if (quantize_pis) pi = fake_quant(pi, qs.g[QG_PI]);
        numpi = pi > 0.f;
        flag = numpi && klist[i];
        if (flag) {
            spi = pi;
#pragma unroll
            for (int l = 0; l < D; ++l) {
                float a = row[off_A(D, C) + lt(l, l)];
                if (qs.mode >= 2) a = fake_quant(a, qs.g[QG_AD]);
                sdiag += a;
            }
        }
    }
    // fixed-order block reduction (shuffle tree, then warp 0 over the 8 warp results)
    __shared__ int s_c[8], s_n[8];
    __shared__ float s_p[8], s_d[8];
    int lane = threadIdx.x & 31, w = threadIdx.x >> 5;
    unsigned bal = __ballot_sync(0xffffffffu, flag);
    unsigned baln = __ballot_sync(0xffffffffu, numpi);
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) {
        spi += __shfl_down_sync(0xffffffffu, spi, o);
        sdiag += __shfl_down_sync(0xffffffffu, sdiag, o);
    }
    if (lane == 0) { s_c[w] = __popc(bal); s_n[w] = __popc(baln); s_p[w] = spi; s_d[w] = sdiag; }
    __syncthreads();
    if (threadIdx.x == 0) {
        PackBlk b = {0, 0, 0, 0, 0.f, 0.f};
        for (int j = 0; j < 8; ++j) { b.count += s_c[j]; b.numpi += s_n[j]; b.sum_pi += s_p[j]; b.sum_diag += s_d[j]; }
        blk[blockIdx.x] = b;
    }
}

// Stage one kernel's compute record.  A is given as a full d x d matrix (row-major).
template <int D, int C>
__device__ __forceinline__ int stage_record(const smoe_cfg& cfg, const float (&A)[D][D], const float* mu, float pi,
                                            const float* nu, const float* ga, float* __restrict__ rec) {
    constexpr int PK = pstride(D, C);
#pragma unroll
    for (int l = 0; l < D; ++l) rec[off_mu(D, C) + l] = mu[l];
#pragma unroll
    for (int l = 0; l < D; ++l)
#pragma unroll
        for (int m = l; m < D; ++m) {
            float q;
            if (cfg.train_inverse_cov) {
                q = 0.5f * (A[l][m] + A[m][l]);          // x^T A x only sees the symmetric part
            } else {
                q = 0.f;
#pragma unroll
                for (int j = 0; j < D; ++j) q = fmaf(A[l][j], A[m][j], q);   // (A A^T)[l][m]
            }
            rec[off_A(D, C) + ut(D, l, m)] = kHalfLog2e * q;
        }
    float coef = pi;
    if (cfg.use_determinant) {
        float det = 1.f;
#pragma unroll
        for (int l = 0; l < D; ++l) det *= A[l][l];
        coef = coef * (det / sqrtf(powf(6.283185307179586f, (float)D)));
    }
    rec[off_pi(D, C)] = log2f(fabsf(coef));            // -inf for coef == 0: the kernel contributes 0
#pragma unroll
    for (int c = 0; c < C; ++c) rec[off_nu(D, C) + c] = nu[c];
#pragma unroll
    for (int l = 0; l < D; ++l)
#pragma unroll
        for (int c = 0; c < C; ++c) {
            float g = ga[l * C + c];
            if (!cfg.train_gammas) g = 0.f;
            if (cfg.use_yuv && cfg.train_gammas && cfg.only_y_gamma && c > 0) g = 0.f;
            rec[off_ga(D, C) + l * C + c] = g;
        }
#pragma unroll
    for (int j = nparam(D, C); j < PK; ++j) rec[j] = 0.f;
    // conservative lower bound of the smallest eigenvalue of Qm (slot P of the record): the culling
    // bound q(x) <= c0 - lam * dist(x, mu)^2.  Negative (indefinite train_inverse_cov forms) = never cull.
    {
        double Q[3][3] = {{0, 0, 0}, {0, 0, 0}, {0, 0, 0}};
#pragma unroll
        for (int l = 0; l < D; ++l)
#pragma unroll
            for (int m = l; m < D; ++m) Q[l][m] = Q[m][l] = (double)rec[off_A(D, C) + ut(D, l, m)];
        double lam;
        if (D == 2) {
            const double hm = 0.5 * (Q[0][0] + Q[1][1]), hd = 0.5 * (Q[0][0] - Q[1][1]);
            lam = hm - sqrt(hd * hd + Q[0][1] * Q[0][1]);
        } else {
            // smallest root of the characteristic polynomial of a symmetric 3x3 (trigonometric form)
            const double p1 = Q[0][1] * Q[0][1] + Q[0][2] * Q[0][2] + Q[1][2] * Q[1][2];
            const double tr = (Q[0][0] + Q[1][1] + Q[2][2]) / 3.0;
            if (p1 == 0.0) {
                lam = fmin(Q[0][0], fmin(Q[1][1], Q[2][2]));
            } else {
                const double a0 = Q[0][0] - tr, a1 = Q[1][1] - tr, a2 = Q[2][2] - tr;
                const double p2 = a0 * a0 + a1 * a1 + a2 * a2 + 2.0 * p1;
                const double pp = sqrt(p2 / 6.0);
                const double b00 = a0 / pp, b11 = a1 / pp, b22 = a2 / pp;
                const double b01 = Q[0][1] / pp, b02 = Q[0][2] / pp, b12 = Q[1][2] / pp;
                double r = 0.5 * (b00 * (b11 * b22 - b12 * b12) - b01 * (b01 * b22 - b12 * b02) + b02 * (b01 * b12 - b11 * b02));
                r = fmin(1.0, fmax(-1.0, r));
                const double phi = acos(r) / 3.0;
                lam = tr + 2.0 * pp * cos(phi + 2.0943951023931953);      // + 2pi/3: the smallest eigenvalue
            }
        }
        const double scale = fabs(Q[0][0]) + fabs(Q[1][1]) + fabs(Q[2][2]);
        lam = lam >= 0.0 ? lam * (1.0 - 1e-3) - 1e-6 * scale : lam;      // keep the bound conservative
        if (lam < 0.0 && !cfg.train_inverse_cov) lam = 0.0;               // A A^T is PSD: rounding only
        rec[nparam(D, C)] = (float)lam;
        // per-axis bounds: d^T Qm d >= kap_l * d_l^2 with kap_l = 1 / (Qm^-1)_ll = det(Qm) / cofactor_ll (the
        // minimum of the form over the other coordinates); much tighter than lam for anisotropic kernels
        double det, cof[3] = {1.0, 1.0, 1.0};
        if (D == 2) {
            det = Q[0][0] * Q[1][1] - Q[0][1] * Q[0][1];
            cof[0] = Q[1][1];
            cof[1] = Q[0][0];
        } else {
            cof[0] = Q[1][1] * Q[2][2] - Q[1][2] * Q[1][2];
            cof[1] = Q[0][0] * Q[2][2] - Q[0][2] * Q[0][2];
            cof[2] = Q[0][0] * Q[1][1] - Q[0][1] * Q[0][1];
            det = Q[0][0] * cof[0] - Q[0][1] * (Q[0][1] * Q[2][2] - Q[1][2] * Q[0][2]) +
                  Q[0][2] * (Q[0][1] * Q[1][2] - Q[1][1] * Q[0][2]);
        }
#pragma unroll
        for (int l = 0; l < D; ++l) {
            double kap = 0.0;
            if (lam > 0.0 && det > 0.0 && cof[l] > 0.0) kap = det / cof[l] * (1.0 - 1e-3) - 1e-6 * scale;
            rec[nparam(D, C) + 1 + l] = (float)fmax(kap, 0.0);
        }
    }
    return coef < 0.f;       // negative weights cannot be carried in the log domain
}

// Per chunk of kChunk consecutive active kernels: bounding box of the centres, smallest lam, largest c0
// (coarse level of the exact culling; layout [mu_min[3] | mu_max[3] | lam_min | c0_max | kap_min[3] | -], stride kCB).
template <int D, int C>
__global__ void __launch_bounds__(kChunk) chunk_bounds_kernel(const float* __restrict__ packed,
                                                              int32_t* __restrict__ counts,
                                                              float* __restrict__ cb,
                                                              const int32_t* __restrict__ nonpos_blk, int nb) {
    constexpr int PK = pstride(D, C);
    if (nonpos_blk && blockIdx.x == 0 && threadIdx.x == 0) {        // counts[2]: kernels with pi * det <= 0
        int s = 0;
        for (int b = 0; b < nb; ++b) s += nonpos_blk[b];
        counts[2] = s;
    }
    const int K = counts[0];
    const int k = blockIdx.x * kChunk + threadIdx.x;
    if ((int)blockIdx.x * kChunk >= K) return;
    float v[kCB];
    const bool on = k < K;
    const float* rec = packed + (size_t)(on ? k : 0) * PK;
#pragma unroll
    for (int l = 0; l < 3; ++l) {
        const float m = (l < D && on) ? rec[off_mu(D, C) + l] : 0.f;
        v[l] = (l < D && on) ? m : INFINITY;        // min
        v[3 + l] = (l < D && on) ? -m : INFINITY;   // max as min of the negation
        v[8 + l] = (l < D && on) ? rec[nparam(D, C) + 1 + l] : INFINITY;
    }
    v[6] = on ? rec[nparam(D, C)] : INFINITY;
    const float c0 = on ? rec[off_pi(D, C)] : -INFINITY;
    v[7] = -c0;
    if (!(c0 == c0)) v[7] = -INFINITY;             // NaN c0: never cull
    v[11] = 0.f;
    __shared__ float s[kChunk / 32][kCB];
#pragma unroll
    for (int q = 0; q < kCB; ++q) {
#pragma unroll
        for (int o = 16; o > 0; o >>= 1) v[q] = fminf(v[q], __shfl_xor_sync(0xffffffffu, v[q], o));
    }
    if ((threadIdx.x & 31) == 0)
#pragma unroll
        for (int q = 0; q < kCB; ++q) s[threadIdx.x >> 5][q] = v[q];
    __syncthreads();
    if (threadIdx.x < kCB) {
        float m = INFINITY;
        for (int w = 0; w < kChunk / 32; ++w) m = fminf(m, s[w][threadIdx.x]);
        const int q = threadIdx.x;
        cb[(size_t)blockIdx.x * kCB + q] = ((q >= 3 && q <= 5) || q == 7) ? -m : m;
    }
}

template <int D, int C>
__global__ void __launch_bounds__(256) pack_scatter_kernel(smoe_cfg cfg, const float* __restrict__ theta,
                                                           const float* __restrict__ mus_grid,
                                                           const uint8_t* __restrict__ klist,
                                                           const int32_t* __restrict__ perm, int K_all, QuantSet qs_in,
                                                           const QuantDyn* __restrict__ qdyn,
                                                           const PackBlk* __restrict__ blk, float* __restrict__ packed,
                                                           int32_t* __restrict__ indices, int32_t* __restrict__ pos,
                                                           int32_t* __restrict__ counts,
                                                           float* __restrict__ regsums, int32_t* __restrict__ nonpos_blk) {
    constexpr int P = nparam(D, C), PK = pstride(D, C);
    const QuantSet qs = qs_in.mode == 3 ? qdyn->qs : qs_in;
    __shared__ int s_red[8];
    __shared__ int s_warp[8];
    int lane = threadIdx.x & 31, w = threadIdx.x >> 5;
    // exclusive prefix of the block counts before this block
    int part = 0;
    for (int b = threadIdx.x; b < (int)blockIdx.x; b += 256) part += blk[b].count;
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) part += __shfl_down_sync(0xffffffffu, part, o);
    if (lane == 0) s_red[w] = part;
    __syncthreads();
    int prefix = 0;
    for (int j = 0; j < 8; ++j) prefix += s_red[j];

    const int j = blockIdx.x * 256 + threadIdx.x;
    const int i = j < K_all ? (perm ? perm[j] : j) : K_all;
    int flag = 0;
    float pi = 0.f;
    const float* row = theta + (size_t)min(i, K_all - 1) * P;
    if (i < K_all) {
        pi = row[off_pi(D, C)];
        if (cfg.quantize_pis) pi = fake_quant(pi, qs.g[QG_PI]);
        flag = (pi > 0.f) && klist[i];
    }
    unsigned bal = __ballot_sync(0xffffffffu, flag);
    if (lane == 0) s_warp[w] = __popc(bal);
    __syncthreads();
    int woff = 0;
    for (int j = 0; j < w; ++j) woff += s_warp[j];
    int dst = prefix + woff + __popc(bal & ((1u << lane) - 1u));
    int neg = 0;
    if (i < K_all) pos[i] = flag ? dst : -1;          // original index -> packed row (the inverse of `indices`)
    if (flag) {
        indices[dst] = i;
        // the variables, fake-quantised when quantization_mode == 2 (smoe.py:482-496)
        float A[D][D], mu[D], nu[C], ga[D * C];
        const bool fq = qs.mode >= 2;
#pragma unroll
        for (int l = 0; l < D; ++l)
#pragma unroll
            for (int m = 0; m < D; ++m) {
                float v = (m <= l) ? row[off_A(D, C) + lt(l, m)] : 0.f;
                if (fq && m <= l) v = fake_quant(v, qs.g[m == l ? QG_AD : QG_AC]);
                A[l][m] = v;
            }
        if (cfg.train_inverse_cov) {
#pragma unroll
            for (int l = 0; l < D; ++l)
#pragma unroll
                for (int m = l + 1; m < D; ++m) A[l][m] = A[m][l];
        }
#pragma unroll
        for (int l = 0; l < D; ++l) {
            float v = row[off_mu(D, C) + l];
            if (fq) v = fake_quant(v, qs.g[QG_MU]);
            if (cfg.use_diff_center) v += mus_grid[(size_t)i * D + l];       // smoe.py:746-747
            mu[l] = v;
        }
#pragma unroll
        for (int c = 0; c < C; ++c) nu[c] = fq ? fake_quant(row[off_nu(D, C) + c], qs.g[QG_NU]) : row[off_nu(D, C) + c];
#pragma unroll
        for (int j = 0; j < D * C; ++j) ga[j] = fq ? fake_quant(row[off_ga(D, C) + j], qs.g[QG_GA]) : row[off_ga(D, C) + j];
        neg = stage_record<D, C>(cfg, A, mu, pi, nu, ga, packed + (size_t)dst * PK);
    }
    int negs = __syncthreads_count(neg);
    if (threadIdx.x == 0) nonpos_blk[blockIdx.x] = negs;
    // the last block publishes the totals (fixed-order sums over the block records)
    if (blockIdx.x == gridDim.x - 1 && threadIdx.x == 0) {
        int K = 0, np_ = 0;
        float sp = 0.f, sd = 0.f;
        for (int b = 0; b < (int)gridDim.x; ++b) { K += blk[b].count; np_ += blk[b].numpi; sp += blk[b].sum_pi; sd += blk[b].sum_diag; }
        counts[0] = K;
        counts[1] = np_;
        counts[3] = 0;
        regsums[0] = sp;
        regsums[1] = sd;
    }
}


template <int D, int C>
__global__ void __launch_bounds__(256) pack_fed_kernel(smoe_cfg cfg, const float* __restrict__ A_, const float* __restrict__ mus,
                                                       const float* __restrict__ nu, const float* __restrict__ ga,
                                                       const float* __restrict__ pis,
                                                       const int32_t* __restrict__ order, int K,
                                                       float* __restrict__ packed, int32_t* __restrict__ indices,
                                                       int32_t* __restrict__ counts, float* __restrict__ scalars_clear,
                                                       uint8_t* __restrict__ infl_clear, int K_all) {
    constexpr int PK = pstride(D, C);
    const int j = blockIdx.x * 256 + threadIdx.x;
    if (scalars_clear && j < SMOE_NSCAL) scalars_clear[j] = 0.f;
    if (infl_clear)
        for (int q = j; q < K_all; q += gridDim.x * 256) infl_clear[q] = 0;
    if (j == 0) { counts[0] = K; counts[1] = K; counts[2] = 0; counts[3] = 0; }
    if (j >= K) return;
    const int i = order ? order[j] : j;            // fed row staged at packed row j
    indices[j] = i;
    float A[D][D];
#pragma unroll
    for (int l = 0; l < D; ++l)
#pragma unroll
        for (int m = 0; m < D; ++m) A[l][m] = A_[(size_t)i * D * D + l * D + m];
    stage_record<D, C>(cfg, A, mus + (size_t)i * D, pis[i], nu + (size_t)i * C, ga + (size_t)i * D * C,
                       packed + (size_t)j * PK);
}

// Hilbert-curve key of each kernel centre on a 2^10 grid per axis (Skilling's transpose algorithm): the packing
// order of smoe_pack.  Unlike Z-order, ANY run of consecutive Hilbert indices is spatially compact, so the 128-record
// chunks of the forward and the 64-record CTAs of the backward stay compact when pruning / kernel lists compact the
// sequence and shift the run boundaries (with Z-order a shifted run straddles the curve's long jumps, its bounding
// box explodes and the tile culling of that chunk / CTA is lost: +25 % backward time on config 3, measured).
// `scale[a]` maps normalised coordinates to a common pixel-isotropic unit (n_a / max n), so that runs are compact
// in PIXELS, the space the tiles live in.
struct KeyScale { float s[3]; };
__global__ void __launch_bounds__(256) spatial_keys_kernel(const float* __restrict__ mu, int K, int d, int stride,
                                                           const float* __restrict__ grid, KeyScale sc,
                                                           long long* __restrict__ keys) {
    const int i = blockIdx.x * 256 + threadIdx.x;
    if (i >= K) return;
    constexpr int B = 10;
    unsigned X[3] = {0, 0, 0};
    for (int a = 0; a < d; ++a) {
        float v = mu[(size_t)i * stride + a];
        if (grid) v += grid[(size_t)i * d + a];
        v = v * sc.s[a] * 1024.f;
        X[a] = v >= 1023.f ? 1023u : (v > 0.f ? (unsigned)v : 0u);      // NaN -> 0
    }
    const unsigned M = 1u << (B - 1);
    for (unsigned Q = M; Q > 1; Q >>= 1) {          // inverse undo
        const unsigned P = Q - 1;
        for (int a = 0; a < d; ++a) {
            if (X[a] & Q) X[0] ^= P;
            else { const unsigned t = (X[0] ^ X[a]) & P; X[0] ^= t; X[a] ^= t; }
        }
    }
    for (int a = 1; a < d; ++a) X[a] ^= X[a - 1];   // Gray encode
    unsigned t = 0;
    for (unsigned Q = M; Q > 1; Q >>= 1)
        if (X[d - 1] & Q) t ^= Q - 1;
    for (int a = 0; a < d; ++a) X[a] ^= t;
    unsigned long long key = 0;
    for (int b = 0; b < B; ++b)
        for (int a = 0; a < d; ++a) key |= (unsigned long long)((X[a] >> b) & 1u) << (b * d + (d - 1 - a));
    keys[i] = (long long)key;
}

// kernel_list[i] = infl[i]: the influence flags are indexed by ORIGINAL kernel index and only active kernels can
// have theirs set, so this is the reference's "clear, then kernel_list[indices] = influential" (smoe.py:1763-1766).
__global__ void __launch_bounds__(256) klist_update_kernel(const uint8_t* __restrict__ infl,
                                                           uint8_t* __restrict__ klist, int K_all) {
    const int i = blockIdx.x * 256 + threadIdx.x;
    if (i < K_all) klist[i] = infl[i] ? 1 : 0;
}

// start of a training / evaluation pass: zero the gradient accumulators (zero_op, smoe.py:1612-1613), the scalar
// blocks of all batches and the influence flags of the first batch -- one launch instead of three fills
__global__ void __launch_bounds__(256) step_begin_kernel(float* __restrict__ grads, size_t n_grads,
                                                         float* __restrict__ scalars, int n_scalars, int scalar_stride,
                                                         int n_rows, uint8_t* __restrict__ infl, int K) {
    const size_t stride = (size_t)gridDim.x * 256;
    const size_t i0 = (size_t)blockIdx.x * 256 + threadIdx.x;
    if (grads)
        for (size_t i = i0; i < n_grads; i += stride) grads[i] = 0.f;
    for (size_t i = i0; i < (size_t)n_rows * n_scalars; i += stride)
        scalars[(i / n_scalars) * scalar_stride + i % n_scalars] = 0.f;
    if (infl)
        for (size_t i = i0; i < (size_t)K; i += stride) infl[i] = 0;
}


// ---- quantization_mode 3: data-dependent fake-quant ranges (smoe.py:497-531) ---------------------------------
// min / max of every parameter group over the kernels whose (fake-quantised) pi is positive, then TF's Nudge().
// One CTA, fixed-order reductions.  Forms (oracle/graph.py:_FakeQuantVarsMasked):
//   A diagonal, nu_e : q = fq(x - min; 0, max - min) + min         (shifted; straight-through)
//   A_corr, musX, gamma_e : q = fq(x; min, max)                    (plain; clipped gradients go to the extremes)
// A_corr's reduce_min / reduce_max run over the whole (d, d) blocks of the variable, whose diagonal and upper
// entries are structural zeros, so 0 always takes part.
// Two stages, so that the scan over the K_all rows uses the whole GPU: (1) every CTA reduces a slice of the rows to
// min / max per group (exact, order-independent), (2) one warp combines the CTA results and applies TF's Nudge().
constexpr int kQBlocks = 256;                 // stage-1 CTAs at most
struct QuantPart { float mn[6], mx[6]; int kept; int pad[3]; };
struct QuantRoutePart { float sum[6]; int cnt[6]; };
struct QuantWork { QuantDyn dyn; QuantPart part[kQBlocks]; QuantRoutePart rpart[kQBlocks]; float share[8]; };

template <int D, int C>
__global__ void __launch_bounds__(256) quant_ranges_stage1(QuantSet qs_static, const float* __restrict__ theta, int K_all,
                                                           QuantWork* __restrict__ wk) {
    constexpr int P = nparam(D, C);
    float mn[6], mx[6];
#pragma unroll
    for (int g = 0; g < 6; ++g) { mn[g] = INFINITY; mx[g] = -INFINITY; }
    int kept = 0;
    for (int i = blockIdx.x * 256 + threadIdx.x; i < K_all; i += gridDim.x * 256) {
        const float* row = theta + (size_t)i * P;
        if (!(fake_quant(row[off_pi(D, C)], qs_static.g[QG_PI]) > 0.f)) continue;          // pis_mask, smoe.py:480
        kept = 1;
        auto upd = [&](int g, float v) { mn[g] = fminf(mn[g], v); mx[g] = fmaxf(mx[g], v); };
#pragma unroll
        for (int l = 0; l < D; ++l) {
            upd(QG_MU, row[off_mu(D, C) + l]);
#pragma unroll
            for (int m = 0; m <= l; ++m) upd(m == l ? QG_AD : QG_AC, row[off_A(D, C) + lt(l, m)]);
        }
        upd(QG_AC, 0.f);
#pragma unroll
        for (int c = 0; c < C; ++c) upd(QG_NU, row[off_nu(D, C) + c]);
#pragma unroll
        for (int j = 0; j < D * C; ++j) upd(QG_GA, row[off_ga(D, C) + j]);
    }
    __shared__ float s_mn[8][6], s_mx[8][6];
    __shared__ int s_kept[8];
#pragma unroll
    for (int g = 0; g < 6; ++g)
#pragma unroll
        for (int o = 16; o > 0; o >>= 1) {
            mn[g] = fminf(mn[g], __shfl_xor_sync(0xffffffffu, mn[g], o));
            mx[g] = fmaxf(mx[g], __shfl_xor_sync(0xffffffffu, mx[g], o));
        }
    kept = __any_sync(0xffffffffu, kept);
    if ((threadIdx.x & 31) == 0) {
#pragma unroll
        for (int g = 0; g < 6; ++g) { s_mn[threadIdx.x >> 5][g] = mn[g]; s_mx[threadIdx.x >> 5][g] = mx[g]; }
        s_kept[threadIdx.x >> 5] = kept;
    }
    __syncthreads();
    if (threadIdx.x != 0) return;
    QuantPart p;
    p.kept = 0;
    for (int g = 0; g < 6; ++g) { p.mn[g] = INFINITY; p.mx[g] = -INFINITY; }
    for (int w = 0; w < 8; ++w) {
        p.kept |= s_kept[w];
#pragma unroll
        for (int g = 0; g < 6; ++g) { p.mn[g] = fminf(p.mn[g], s_mn[w][g]); p.mx[g] = fmaxf(p.mx[g], s_mx[w][g]); }
    }
    wk->part[blockIdx.x] = p;
}

__global__ void __launch_bounds__(32) quant_ranges_stage2(smoe_cfg cfg, QuantSet qs_static, int train_musx, int nblocks,
                                                          QuantWork* __restrict__ wk) {
    float mn[6], mx[6];
#pragma unroll
    for (int g = 0; g < 6; ++g) { mn[g] = INFINITY; mx[g] = -INFINITY; }
    int kept = 0;
    for (int b = threadIdx.x; b < nblocks; b += 32) {
        const QuantPart p = wk->part[b];
        kept |= p.kept;
#pragma unroll
        for (int g = 0; g < 6; ++g) { mn[g] = fminf(mn[g], p.mn[g]); mx[g] = fmaxf(mx[g], p.mx[g]); }
    }
#pragma unroll
    for (int g = 0; g < 6; ++g)
#pragma unroll
        for (int o = 16; o > 0; o >>= 1) {
            mn[g] = fminf(mn[g], __shfl_xor_sync(0xffffffffu, mn[g], o));
            mx[g] = fmaxf(mx[g], __shfl_xor_sync(0xffffffffu, mx[g], o));
        }
    kept = __any_sync(0xffffffffu, kept);
    if (threadIdx.x != 0) return;
    QuantDyn q;
    q.qs = qs_static;
    q.qs.mode = 3;
    const int bits[6] = {cfg.q_bits[0], cfg.q_bits[1], cfg.q_bits[2], cfg.q_bits[3], cfg.q_bits[4], cfg.q_bits[0]};
    for (int g = 0; g < 6; ++g) {
        q.mn[g] = mn[g];
        q.mx[g] = mx[g];
        if (g == QG_PI) continue;                                  // pis keep their fixed bounds (smoe.py:474-478)
        Nudged n = {0.f, 0.f, 1.f, 1.f, 0.f, QF_IDENT};
        const bool shifted = g == QG_AD || g == QG_NU;
        if (kept && !(g == QG_MU && !train_musx)) {
            const float lo = shifted ? 0.f : mn[g], hi = shifted ? __fsub_rn(mx[g], mn[g]) : mx[g];
            if (lo == 0.f && hi == 0.f) {
                n.flags = QF_ZERO;
            } else {
                n = nudge(lo, hi, bits[g]);
                n.flags = shifted ? QF_PASS : QF_ROUTE;
            }
            n.shift = shifted ? mn[g] : 0.f;
        }
        q.qs.g[g] = n;
    }
    wk->dyn = q;
}

// The variables as the graph uses them (the q* tensors, smoe.py:482-538; what get_params returns, smoe.py:1796-1798)
template <int D, int C>
__global__ void __launch_bounds__(256) fake_quant_theta_kernel(smoe_cfg cfg, QuantSet qs_in,
                                                               const QuantDyn* __restrict__ qdyn,
                                                               const float* __restrict__ theta, int K_all,
                                                               float* __restrict__ out,
                                                               float* __restrict__ structural) {
    constexpr int P = nparam(D, C);
    const QuantSet qs = qs_in.mode == 3 ? qdyn->qs : qs_in;
    const int i = blockIdx.x * 256 + threadIdx.x;
    if (i >= K_all) return;
    const float* row = theta + (size_t)i * P;
    float* o = out + (size_t)i * P;
    const bool fq = qs.mode >= 2;
#pragma unroll
    for (int l = 0; l < D; ++l) {
        o[off_mu(D, C) + l] = fq ? fake_quant(row[off_mu(D, C) + l], qs.g[QG_MU]) : row[off_mu(D, C) + l];
#pragma unroll
        for (int m = 0; m <= l; ++m) {
            const float v = row[off_A(D, C) + lt(l, m)];
            o[off_A(D, C) + lt(l, m)] = fq ? fake_quant(v, qs.g[m == l ? QG_AD : QG_AC]) : v;
        }
    }
    o[off_pi(D, C)] = cfg.quantize_pis ? fake_quant(row[off_pi(D, C)], qs.g[QG_PI]) : row[off_pi(D, C)];
#pragma unroll
    for (int c = 0; c < C; ++c) o[off_nu(D, C) + c] = fq ? fake_quant(row[off_nu(D, C) + c], qs.g[QG_NU]) : row[off_nu(D, C) + c];
#pragma unroll
    for (int j = 0; j < D * C; ++j) o[off_ga(D, C) + j] = fq ? fake_quant(row[off_ga(D, C) + j], qs.g[QG_GA]) : row[off_ga(D, C) + j];
    if (i == 0 && structural) {   // what the structural zeros of A_diagonal (off-diagonal) / A_corr (diagonal, upper) become
        structural[0] = fq ? fake_quant(0.f, qs.g[QG_AD]) : 0.f;
        structural[1] = fq ? fake_quant(0.f, qs.g[QG_AC]) : 0.f;
    }
}

// Mode-3 gradient routing for the plain groups (A_corr, musX, gamma_e), applied once to the accumulated gradient
// before Adam: in-range elements keep their gradient; the gradients of elements below nudged_min / above
// nudged_max go to the `min` / `max` inputs of fake_quant_with_min_max_vars and from there, through
// reduce_min / reduce_max (equal shares among ties), to the extreme elements of the kept kernels.  When the
// extreme of A_corr is one of its structural zeros, that share lands on a variable entry the graph never reads;
// it is dropped here (deviation from HEAD, which would start moving that unused entry).
template <int D, int C, typename FN>
__device__ __forceinline__ void route_visit(const float* __restrict__ row, float* __restrict__ gr, FN&& fn) {
#pragma unroll
    for (int l = 0; l < D; ++l) {
        fn(1, row[off_mu(D, C) + l], gr[off_mu(D, C) + l]);
#pragma unroll
        for (int m = 0; m < l; ++m) fn(0, row[off_A(D, C) + lt(l, m)], gr[off_A(D, C) + lt(l, m)]);
    }
#pragma unroll
    for (int j = 0; j < D * C; ++j) fn(2, row[off_ga(D, C) + j], gr[off_ga(D, C) + j]);
}

// stage 1: per CTA, the clipped gradient mass below / above the nudged range of each plain group and the number of
// elements tied at the group's min / max (fixed-order reductions)
template <int D, int C>
__global__ void __launch_bounds__(256) quant_route_stage1(QuantSet qs_static, const float* __restrict__ theta, int K_all,
                                                          float* __restrict__ grads, QuantWork* __restrict__ wk) {
    constexpr int P = nparam(D, C);
    const QuantDyn* qdyn = &wk->dyn;
    const QuantSet qs = qdyn->qs;
    const int groups[3] = {QG_AC, QG_MU, QG_GA};
    float sum[6] = {0.f, 0.f, 0.f, 0.f, 0.f, 0.f};      // [2*j]: below, [2*j+1]: above
    int cnt[6] = {0, 0, 0, 0, 0, 0};                    // ties at min / max
    for (int i = blockIdx.x * 256 + threadIdx.x; i < K_all; i += gridDim.x * 256) {
        const float* row = theta + (size_t)i * P;
        if (!(fake_quant(row[off_pi(D, C)], qs_static.g[QG_PI]) > 0.f)) continue;
        route_visit<D, C>(row, grads + (size_t)i * P, [&](int j, float x, float& g) {
            const Nudged n = qs.g[groups[j]];
            if (!(n.flags & QF_ROUTE)) return;
            if (x < n.nmin) sum[2 * j] += g;
            if (x > n.nmax) sum[2 * j + 1] += g;
            if (x == qdyn->mn[groups[j]]) cnt[2 * j] += 1;
            if (x == qdyn->mx[groups[j]]) cnt[2 * j + 1] += 1;
        });
        // structural zeros of A_corr take part in the ties at min / max
        if (qdyn->mn[QG_AC] == 0.f) cnt[0] += D * D - D * (D - 1) / 2;
        if (qdyn->mx[QG_AC] == 0.f) cnt[1] += D * D - D * (D - 1) / 2;
    }
    __shared__ float s_sum[8][6];
    __shared__ int s_cnt[8][6];
#pragma unroll
    for (int q = 0; q < 6; ++q)
#pragma unroll
        for (int o = 16; o > 0; o >>= 1) {
            sum[q] += __shfl_down_sync(0xffffffffu, sum[q], o);
            cnt[q] += __shfl_down_sync(0xffffffffu, cnt[q], o);
        }
    if ((threadIdx.x & 31) == 0)
#pragma unroll
        for (int q = 0; q < 6; ++q) { s_sum[threadIdx.x >> 5][q] = sum[q]; s_cnt[threadIdx.x >> 5][q] = cnt[q]; }
    __syncthreads();
    if (threadIdx.x < 6) {
        float t = 0.f;
        int c = 0;
        for (int w = 0; w < 8; ++w) { t += s_sum[w][threadIdx.x]; c += s_cnt[w][threadIdx.x]; }
        wk->rpart[blockIdx.x].sum[threadIdx.x] = t;
        wk->rpart[blockIdx.x].cnt[threadIdx.x] = c;
    }
}

// stage 2: equal shares of the clipped mass for the tied extreme elements
__global__ void __launch_bounds__(32) quant_route_stage2(int nblocks, QuantWork* __restrict__ wk) {
    if (threadIdx.x >= 6) return;
    float t = 0.f;
    int c = 0;
    for (int b = 0; b < nblocks; ++b) { t += wk->rpart[b].sum[threadIdx.x]; c += wk->rpart[b].cnt[threadIdx.x]; }
    wk->share[threadIdx.x] = c > 0 ? t / (float)c : 0.f;
}

// stage 3: in-range mask + shares.  Mode-3 gradient routing for the plain groups (A_corr, musX, gamma_e), applied once
// to the accumulated gradient before Adam: in-range elements keep their gradient; the gradients of elements below
// nudged_min / above nudged_max go to the `min` / `max` inputs of fake_quant_with_min_max_vars and from there, through
// reduce_min / reduce_max (equal shares among ties), to the extreme elements of the kept kernels.  When the extreme
// of A_corr is one of its structural zeros, that share lands on a variable entry the graph never reads; it is
// dropped here (deviation from HEAD, which would start moving that unused entry).
template <int D, int C>
__global__ void __launch_bounds__(256) quant_route_stage3(QuantSet qs_static, const float* __restrict__ theta, int K_all,
                                                          float* __restrict__ grads, const QuantWork* __restrict__ wk) {
    constexpr int P = nparam(D, C);
    const QuantDyn* qdyn = &wk->dyn;
    const QuantSet qs = qdyn->qs;
    const int groups[3] = {QG_AC, QG_MU, QG_GA};
    for (int i = blockIdx.x * 256 + threadIdx.x; i < K_all; i += gridDim.x * 256) {
        const float* row = theta + (size_t)i * P;
        if (!(fake_quant(row[off_pi(D, C)], qs_static.g[QG_PI]) > 0.f)) continue;
        route_visit<D, C>(row, grads + (size_t)i * P, [&](int j, float x, float& g) {
            const Nudged n = qs.g[groups[j]];
            if (!(n.flags & QF_ROUTE)) return;
            float v = (x >= n.nmin && x <= n.nmax) ? g : 0.f;
            if (x == qdyn->mn[groups[j]]) v += wk->share[2 * j];
            if (x == qdyn->mx[groups[j]]) v += wk->share[2 * j + 1];
            g = v;
        });
    }
}

}  // namespace smoe

using namespace smoe;

extern "C" {

int smoe_abi_version(void) { return SMOE_ABI_VERSION; }
const char* smoe_last_error(void) { return smoe::g_err; }
int smoe_param_count(int d, int C) { return nparam(d, C); }
int smoe_packed_stride(int d, int C) { return pstride(d, C); }
int smoe_pix_stride(int d, int C, const smoe_batch* b) { return pix_stride(d, C, b->tile[d - 1]); }
int smoe_num_tiles(const smoe_batch* b) {
    int n = 1;
    for (int i = 0; i < 3; ++i) n *= (b->extent[i] + b->tile[i] - 1) / b->tile[i];
    return n;
}
size_t smoe_pack_workspace_bytes(int K_all) {
    size_t nb = (size_t)(K_all + 255) / 256;
    return nb * sizeof(PackBlk) + nb * sizeof(int32_t) + 256;
}

int smoe_pack(const smoe_cfg* cfg, const float* theta, const float* mus_grid, const void* quant_ranges,
              const uint8_t* kernel_list, const int32_t* perm, int K_all, float* packed, int32_t* indices, int32_t* pos,
              int32_t* counts, float* regsums, float* chunk_bounds, void* workspace, float* grads_clear,
              float* scalars_clear, uint8_t* infl_clear, void* stream) {
    SMOE_REQUIRE(!cfg || !cfg->use_diff_center || mus_grid, "use_diff_center needs mus_grid");
    SMOE_REQUIRE(!cfg || cfg->quantization_mode != 3 || quant_ranges, "quantization_mode 3 needs quant_ranges");
    const QuantDyn* qdyn = (const QuantDyn*)quant_ranges;
    SMOE_REQUIRE(cfg && theta && kernel_list && packed && indices && pos && counts && regsums && chunk_bounds && workspace,
                 "null argument");
    SMOE_REQUIRE(K_all > 0, "K_all must be positive");
    cudaStream_t st = (cudaStream_t)stream;
    int nb = (K_all + 255) / 256;
    PackBlk* blk = (PackBlk*)workspace;
    int32_t* nonpos_blk = (int32_t*)((char*)workspace + (size_t)nb * sizeof(PackBlk));
    const QuantSet qs = make_quantset(cfg);
#define CALL(D, C)                                                                                              \
    pack_count_kernel<D, C><<<nb, 256, 0, st>>>(theta, kernel_list, perm, K_all, cfg->quantize_pis, qs, qdyn, blk,  \
                                                grads_clear, scalars_clear, infl_clear);                          \
    pack_scatter_kernel<D, C><<<nb, 256, 0, st>>>(*cfg, theta, mus_grid, kernel_list, perm, K_all, qs, qdyn, blk, \
                                                  packed, indices, pos, counts, regsums, nonpos_blk);
    SMOE_DISPATCH_DC(cfg->d, cfg->C, CALL)
#undef CALL
    const int nchunks = (K_all + kChunk - 1) / kChunk;
#define CALL(D, C) chunk_bounds_kernel<D, C><<<nchunks, kChunk, 0, st>>>(packed, counts, chunk_bounds, nonpos_blk, nb);
    SMOE_DISPATCH_DC(cfg->d, cfg->C, CALL)
#undef CALL
    return check_launch("smoe_pack");
}

size_t smoe_quant_ranges_bytes(void) { return sizeof(QuantWork); }      // QuantDyn first, then the two-stage scratch

int smoe_quant_ranges(const smoe_cfg* cfg, const float* theta, int K_all, int train_musx, void* quant_ranges,
                      void* stream) {
    SMOE_REQUIRE(cfg && theta && quant_ranges && K_all > 0, "bad argument");
    SMOE_REQUIRE(cfg->quantization_mode == 3, "only meaningful for quantization_mode 3");
    QuantSet qs = make_quantset(cfg);
    qs.g[QG_PI] = nudge(cfg->pis_lb, cfg->pis_ub, cfg->pis_bits);
    cudaStream_t st = (cudaStream_t)stream;
    int nb = (K_all + 255) / 256;
    if (nb > kQBlocks) nb = kQBlocks;
#define CALL(D, C) quant_ranges_stage1<D, C><<<nb, 256, 0, st>>>(qs, theta, K_all, (QuantWork*)quant_ranges);
    SMOE_DISPATCH_DC(cfg->d, cfg->C, CALL)
#undef CALL
    quant_ranges_stage2<<<1, 32, 0, st>>>(*cfg, qs, train_musx, nb, (QuantWork*)quant_ranges);
    return check_launch("smoe_quant_ranges");
}

int smoe_fake_quant_theta(const smoe_cfg* cfg, const float* theta, const void* quant_ranges, int K_all, float* out,
                          float* structural, void* stream) {
    SMOE_REQUIRE(cfg && theta && out && K_all > 0, "bad argument");
    SMOE_REQUIRE(cfg->quantization_mode != 3 || quant_ranges, "quantization_mode 3 needs quant_ranges");
    const QuantSet qs = make_quantset(cfg);
    cudaStream_t st = (cudaStream_t)stream;
#define CALL(D, C)                                                                                              \
    fake_quant_theta_kernel<D, C><<<(K_all + 255) / 256, 256, 0, st>>>(*cfg, qs, (const QuantDyn*)quant_ranges, \
                                                                       theta, K_all, out, structural);
    SMOE_DISPATCH_DC(cfg->d, cfg->C, CALL)
#undef CALL
    return check_launch("smoe_fake_quant_theta");
}

int smoe_quant_route(const smoe_cfg* cfg, const float* theta, const void* quant_ranges, int K_all, float* grads,
                     void* stream) {
    SMOE_REQUIRE(cfg && theta && quant_ranges && grads && K_all > 0, "bad argument");
    SMOE_REQUIRE(cfg->quantization_mode == 3, "only meaningful for quantization_mode 3");
    QuantSet qs = make_quantset(cfg);
    qs.g[QG_PI] = nudge(cfg->pis_lb, cfg->pis_ub, cfg->pis_bits);
    cudaStream_t st = (cudaStream_t)stream;
    int nb = (K_all + 255) / 256;
    if (nb > kQBlocks) nb = kQBlocks;
    QuantWork* wk = (QuantWork*)const_cast<void*>(quant_ranges);       // the scratch part of the opaque block
#define CALL(D, C)                                                                 \
    quant_route_stage1<D, C><<<nb, 256, 0, st>>>(qs, theta, K_all, grads, wk);     \
    quant_route_stage2<<<1, 32, 0, st>>>(nb, wk);                                  \
    quant_route_stage3<D, C><<<nb, 256, 0, st>>>(qs, theta, K_all, grads, wk);
    SMOE_DISPATCH_DC(cfg->d, cfg->C, CALL)
#undef CALL
    return check_launch("smoe_quant_route");
}

int smoe_pack_fed(const smoe_cfg* cfg, const float* A, const float* musX, const float* nu_e, const float* gamma_e,
                  const float* pis, const int32_t* order, int K, float* packed, int32_t* indices, int32_t* counts,
                  float* chunk_bounds, float* scalars_clear, uint8_t* infl_clear, int K_all, void* stream) {
    SMOE_REQUIRE(cfg && A && musX && nu_e && gamma_e && pis && packed && indices && counts && chunk_bounds, "null argument");
    SMOE_REQUIRE(K > 0, "K must be positive");
    cudaStream_t st = (cudaStream_t)stream;
    int nb = (K + 255) / 256;
#define CALL(D, C) pack_fed_kernel<D, C><<<nb, 256, 0, st>>>(*cfg, A, musX, nu_e, gamma_e, pis, order, K, packed, indices, counts, \
                                                              scalars_clear, infl_clear, infl_clear ? K_all : 0);
    SMOE_DISPATCH_DC(cfg->d, cfg->C, CALL)
#undef CALL
    const int nchunks = (K + kChunk - 1) / kChunk;
#define CALL(D, C) chunk_bounds_kernel<D, C><<<nchunks, kChunk, 0, st>>>(packed, counts, chunk_bounds, nullptr, 0);
    SMOE_DISPATCH_DC(cfg->d, cfg->C, CALL)
#undef CALL
    return check_launch("smoe_pack_fed");
}

int smoe_update_kernel_list(const uint8_t* infl, uint8_t* kernel_list, int K_all, void* stream) {
    SMOE_REQUIRE(infl && kernel_list && K_all > 0, "bad argument");
    cudaStream_t st = (cudaStream_t)stream;
    int nb = (K_all + 255) / 256;
    klist_update_kernel<<<nb, 256, 0, st>>>(infl, kernel_list, K_all);
    return check_launch("smoe_update_kernel_list");
}

int smoe_spatial_keys(const float* centres, int K, int d, int row_stride, const float* grid, const float* scale,
                      long long* keys, void* stream) {
    SMOE_REQUIRE(centres && keys && K > 0 && (d == 2 || d == 3) && row_stride >= d, "bad argument");
    KeyScale sc = {{1.f, 1.f, 1.f}};
    if (scale)
        for (int a = 0; a < d; ++a) sc.s[a] = scale[a];
    spatial_keys_kernel<<<(K + 255) / 256, 256, 0, (cudaStream_t)stream>>>(centres, K, d, row_stride, grid, sc, keys);
    return check_launch("smoe_spatial_keys");
}

int smoe_step_begin(float* grads, size_t n_grads, float* scalars, int n_rows, int row_stride, uint8_t* infl, int K,
                    void* stream) {
    SMOE_REQUIRE(scalars && n_rows > 0 && row_stride >= SMOE_NSCAL, "bad argument");
    size_t n = n_grads > (size_t)K ? n_grads : (size_t)K;
    if (n < (size_t)n_rows * SMOE_NSCAL) n = (size_t)n_rows * SMOE_NSCAL;
    int nb = (int)((n + 255) / 256);
    if (nb > 148 * 8) nb = 148 * 8;
    step_begin_kernel<<<nb, 256, 0, (cudaStream_t)stream>>>(grads, grads ? n_grads : 0, scalars, SMOE_NSCAL, row_stride,
                                                           n_rows, infl, infl ? K : 0);
    return check_launch("smoe_step_begin");
}

}  // extern "C"
