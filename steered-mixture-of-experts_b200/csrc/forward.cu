// Fused SMoE forward (smoe_forward).  Replaces smoe.py:777-858, 899-937, 1053.
//
// Pixel-stationary: a CTA (128 threads, 4 resident per SM) owns a spatially compact tile of
// SMOE_TPIX = 512 pixels (4 per thread, in registers) and streams ALL active kernels past it twice:
//   sweep A  S_n = sum_k 2^{q_k(x_n)}                       (the normaliser of smoe.py:819-821)
//   sweep B  w = 2^{q}/S, m = w > tau, r_c += m*w*E_kc(x)   (smoe.py:823-848; needs the FINAL S,
//            and gates are not renormalised after thresholding, so one sweep is not enough)
// Kernel records arrive in shared memory by TMA bulk copies (cp.async.bulk + mbarrier, double
// buffered) and are re-expressed per tile in tile-centred coordinates, where the whole gate
// logit is one quadratic  q(x') = qc + ql.x' + x'^T qq x'  evaluated by Horner in T+d FFMA
// (5 for d=2, 9 for d=3) + one ex2.approx: -1/2 maha * log2(e) + log2(pi*det/(2pi)^{d/2}).
// The quadratic form covers both ||A^T(x-mu)||^2 and the train_inverse_cov branch x^T A x.
// The N x K gate matrix is never materialised.  FP32 FFMA + MUFU.EX2 bound; no tensor cores
// (inner dimensions are d = 2..3 and C = 1..3).
//
// Exact-zero skipping (results are bit-identical to dense execution, cfg.dense_exec = 1):
//   sweep A skips ex2 + add for a thread-iteration whose 4 logits are all < -126, where
//     ex2.approx.ftz returns exactly +0;
//   sweep B tests the threshold in the log domain, q > log2(tau*S), so it needs no ex2 at all
//     unless a gate passes (<0.2 % of the pairs at the benchmark shapes), and skips the expert
//     part otherwise.
#include "smoe_common.cuh"

namespace smoe {

template <int D, int C>
struct Rec {
    static constexpr int T = tri(D);
    static constexpr int GN = T + D + 1;          // qq (upper-tri, off-diagonals doubled) | ql | qc
    static constexpr int EN = C + D * C;          // nu' | gamma
    static constexpr int RC = (GN + EN + 3) / 4 * 4;
    static constexpr int NG4 = (GN + 3) / 4;      // float4 loads that cover the geometry part
    static constexpr int OQ = 0, OL = T, OC = T + D, ONU = GN, OGA = GN + C;
};

// raw packed record + tile centre -> tile-centred compute record
template <int D, int C>
__device__ __forceinline__ void transform_record(const float* __restrict__ raw, const float (&ctr)[3],
                                                 float* __restrict__ out) {
    using R = Rec<D, C>;
    float mu[D], Qm[D][D], v[D];
#pragma unroll
    for (int l = 0; l < D; ++l) mu[l] = raw[off_mu(D, C) + l] - ctr[l];
#pragma unroll
    for (int l = 0; l < D; ++l)
#pragma unroll
        for (int m = l; m < D; ++m) Qm[l][m] = Qm[m][l] = raw[off_A(D, C) + ut(D, l, m)];
    float qc = raw[off_pi(D, C)];
#pragma unroll
    for (int l = 0; l < D; ++l) {
        v[l] = 0.f;
#pragma unroll
        for (int m = 0; m < D; ++m) v[l] = fmaf(Qm[l][m], mu[m], v[l]);
    }
#pragma unroll
    for (int l = 0; l < D; ++l) qc = fmaf(-mu[l], v[l], qc);
#pragma unroll
    for (int l = 0; l < D; ++l) {
        out[R::OL + l] = 2.f * v[l];
#pragma unroll
        for (int m = l; m < D; ++m) out[R::OQ + ut(D, l, m)] = (l == m) ? -Qm[l][m] : -2.f * Qm[l][m];
    }
    out[R::OC] = qc;
#pragma unroll
    for (int c = 0; c < C; ++c) {
        float nu = raw[off_nu(D, C) + c];
#pragma unroll
        for (int l = 0; l < D; ++l) {
            float g = raw[off_ga(D, C) + l * C + c];
            nu = fmaf(g, ctr[l], nu);
            out[R::OGA + l * C + c] = g;
        }
        out[R::ONU + c] = nu;
    }
#pragma unroll
    for (int j = R::GN + R::EN; j < R::RC; ++j) out[j] = 0.f;
}

// q(x') by Horner: T + D FFMA
template <int D, int C>
__device__ __forceinline__ float logit(const float* __restrict__ f, const float (&x)[D]) {
    using R = Rec<D, C>;
    float q = f[R::OC];
#pragma unroll
    for (int l = 0; l < D; ++l) {
        float t = f[R::OL + l];
#pragma unroll
        for (int m = l; m < D; ++m) t = fmaf(f[R::OQ + ut(D, l, m)], x[m], t);
        q = fmaf(t, x[l], q);
    }
    return q;
}

struct FwdArgs {
    smoe_cfg cfg;
    smoe_batch b;
    const float* packed;
    const int32_t* indices;
    const int32_t* counts;
    const float* image;
    const float* ax[3];
    float* res;
    float* res_pre;
    int32_t* argmax;
    uint8_t* infl;
    float* pix;
    float* scalars;
    float* partials;
    int32_t* ticket;
    int ntiles, nt1, nt2;
    float tau, eps, q_scale, q_inv_scale;
};

template <int D, int C>
__global__ void __launch_bounds__(kThreadsF, 4) forward_kernel(const FwdArgs a) {
    using R = Rec<D, C>;
    constexpr int PK = pstride(D, C);
    constexpr int PPT = kPixPerThread;
    extern __shared__ __align__(128) unsigned char smem_raw[];
    float* raw0 = reinterpret_cast<float*>(smem_raw);
    float* raw1 = raw0 + kChunk * PK;
    float* crec = raw1 + kChunk * PK;
    uint64_t* bar = reinterpret_cast<uint64_t*>(crec + kChunk * R::RC);
    float* red = reinterpret_cast<float*>(bar + 2);          // [warps][8]

    const int tid = threadIdx.x;
    const int K = a.counts[0];
    const int nchunks = (K + kChunk - 1) / kChunk;
    if (tid == 0) {
        mbar_init(&bar[0], 1);
        mbar_init(&bar[1], 1);
        fence_barrier_init();
    }
    __syncthreads();
    uint32_t phase0 = 0, phase1 = 0;

    auto issue = [&](int ci, int buf) {
        int nk = min(kChunk, K - ci * kChunk);
        uint32_t bytes = (uint32_t)nk * PK * 4u;
        fence_proxy_async();
        mbar_expect_tx(&bar[buf], bytes);
        tma_load_1d(buf ? raw1 : raw0, a.packed + (size_t)ci * kChunk * PK, bytes, &bar[buf]);
    };

    float lsum[C];
#pragma unroll
    for (int c = 0; c < C; ++c) lsum[c] = 0.f;
    float sqsum = 0.f;
    int nonfinite = 0;

    const int e0 = a.b.tile[0], e1 = a.b.tile[1], e2 = a.b.tile[2];
    (void)e0;
    for (int tile = blockIdx.x; tile < a.ntiles; tile += gridDim.x) {
        // ---- tile geometry ------------------------------------------------------------
        int tt[3];
        tt[2] = tile % a.nt2;
        tt[1] = (tile / a.nt2) % a.nt1;
        tt[0] = tile / (a.nt2 * a.nt1);
        int lo[3], hi[3];
        float ctr[3];
#pragma unroll
        for (int i = 0; i < 3; ++i) {
            lo[i] = a.b.origin[i] + tt[i] * a.b.tile[i];
            hi[i] = min(lo[i] + a.b.tile[i], a.b.origin[i] + a.b.extent[i]) - 1;
            ctr[i] = (i < D) ? 0.5f * (a.ax[i][lo[i]] + a.ax[i][hi[i]]) : 0.f;
        }
        float x[PPT][D];
        long long gidx[PPT];     // linear pixel index in the image buffer, -1 when outside
#pragma unroll
        for (int p = 0; p < PPT; ++p) {
            int j = p * kThreadsF + tid;
            int i2 = j % e2, i1 = (j / e2) % e1, i0 = j / (e2 * e1);
            int g0 = lo[0] + i0, g1 = lo[1] + i1, g2 = lo[2] + i2;
            bool ok = g0 <= hi[0] && g1 <= hi[1] && g2 <= hi[2];
            gidx[p] = ok ? ((long long)g0 * a.b.dims[1] + g1) * a.b.dims[2] + g2 : -1;
            int gg[3] = {min(g0, hi[0]), min(g1, hi[1]), min(g2, hi[2])};
#pragma unroll
            for (int l = 0; l < D; ++l) x[p][l] = a.ax[l][gg[l]] - ctr[l];
        }

        // ---- sweep A: normaliser ------------------------------------------------------
        float S[PPT];
#pragma unroll
        for (int p = 0; p < PPT; ++p) S[p] = 0.f;
        if (tid == 0 && nchunks > 0) {
            issue(0, 0);
            if (nchunks > 1) issue(1, 1);
        }
        for (int ci = 0; ci < nchunks; ++ci) {
            const int buf = ci & 1;
            const int nk = min(kChunk, K - ci * kChunk);
            if (buf) { mbar_wait(&bar[1], phase1); phase1 ^= 1; } else { mbar_wait(&bar[0], phase0); phase0 ^= 1; }
            for (int kt = tid; kt < nk; kt += kThreadsF)
                transform_record<D, C>((buf ? raw1 : raw0) + kt * PK, ctr, crec + kt * R::RC);
            __syncthreads();
            if (tid == 0 && ci + 2 < nchunks) issue(ci + 2, buf);
#pragma unroll 2
            for (int kk = 0; kk < nk; ++kk) {
                float f[4 * R::NG4];
                const float4* r4 = reinterpret_cast<const float4*>(crec + kk * R::RC);
#pragma unroll
                for (int j = 0; j < R::NG4; ++j) {
                    float4 v = r4[j];
                    f[4 * j] = v.x; f[4 * j + 1] = v.y; f[4 * j + 2] = v.z; f[4 * j + 3] = v.w;
                }
                float q[PPT];
                float qmax = -INFINITY;
#pragma unroll
                for (int p = 0; p < PPT; ++p) {
                    q[p] = logit<D, C>(f, x[p]);
                    qmax = fmaxf(qmax, q[p]);
                }
                // ex2.approx.ftz(q) == +0 exactly for q < -126: adding it would not change S
                if (__builtin_expect(__any_sync(0xffffffffu, qmax >= -126.0f) || a.cfg.dense_exec, 0)) {
#pragma unroll
                    for (int p = 0; p < PPT; ++p) S[p] += ex2f(q[p]);
                }
            }
            __syncthreads();
        }

        // ---- sweep B: thresholded gates, experts ----------------------------------------
        float qthr[PPT], r[PPT][C], bestw[PPT];
        int bestk[PPT];
#pragma unroll
        for (int p = 0; p < PPT; ++p) {
            float Sc = fmaxf(S[p], kSFloor);
            // w = e/S > tau  <=>  q > log2(tau * S); +inf disables pixels outside the batch
            qthr[p] = gidx[p] >= 0 ? log2f(a.tau * Sc) : INFINITY;
            bestw[p] = 0.f;
            bestk[p] = -1;
#pragma unroll
            for (int c = 0; c < C; ++c) r[p][c] = 0.f;
        }
        if (tid == 0 && nchunks > 0) {
            issue(0, 0);
            if (nchunks > 1) issue(1, 1);
        }
        for (int ci = 0; ci < nchunks; ++ci) {
            const int buf = ci & 1;
            const int nk = min(kChunk, K - ci * kChunk);
            if (buf) { mbar_wait(&bar[1], phase1); phase1 ^= 1; } else { mbar_wait(&bar[0], phase0); phase0 ^= 1; }
            for (int kt = tid; kt < nk; kt += kThreadsF)
                transform_record<D, C>((buf ? raw1 : raw0) + kt * PK, ctr, crec + kt * R::RC);
            __syncthreads();
            if (tid == 0 && ci + 2 < nchunks) issue(ci + 2, buf);
#pragma unroll 2
            for (int kk = 0; kk < nk; ++kk) {
                float f[R::RC];
                const float4* r4 = reinterpret_cast<const float4*>(crec + kk * R::RC);
#pragma unroll
                for (int j = 0; j < R::NG4; ++j) {
                    float4 v = r4[j];
                    f[4 * j] = v.x; f[4 * j + 1] = v.y; f[4 * j + 2] = v.z; f[4 * j + 3] = v.w;
                }
                float q[PPT];
                bool any = false;
#pragma unroll
                for (int p = 0; p < PPT; ++p) {
                    q[p] = logit<D, C>(f, x[p]);
                    any |= (q[p] > qthr[p]);
                }
                if (__builtin_expect(__any_sync(0xffffffffu, any) || a.cfg.dense_exec, 0)) {
                    float w[PPT];       // w = e/S = tau * 2^(q - log2(tau*S)), the form the backward recomputes
#pragma unroll
                    for (int p = 0; p < PPT; ++p) w[p] = a.tau * ex2f(q[p] - qthr[p]);
#pragma unroll
                    for (int j = R::NG4; j < R::RC / 4; ++j) {
                        float4 v = r4[j];
                        f[4 * j] = v.x; f[4 * j + 1] = v.y; f[4 * j + 2] = v.z; f[4 * j + 3] = v.w;
                    }
                    const int kglob = ci * kChunk + kk;
#pragma unroll
                    for (int p = 0; p < PPT; ++p) {
                        const bool pass = q[p] > qthr[p];
                        const float wm = pass ? w[p] : 0.f;
#pragma unroll
                        for (int c = 0; c < C; ++c) {
                            float E = f[R::ONU + c];
#pragma unroll
                            for (int l = 0; l < D; ++l) E = fmaf(f[R::OGA + l * C + c], x[p][l], E);
                            r[p][c] = fmaf(wm, E, r[p][c]);
                        }
                        if (pass && w[p] > bestw[p]) { bestw[p] = w[p]; bestk[p] = kglob; }
                    }
                    if (any && a.infl) a.infl[kglob] = 1;
                }
            }
            __syncthreads();
        }

        // ---- epilogue: clip, output fake-quant, loss, backward state -----------------------
#pragma unroll
        for (int p = 0; p < PPT; ++p) {
            const int j = p * kThreadsF + tid;
            float g[C], gr = 0.f;
            if (gidx[p] >= 0) {
#pragma unroll
                for (int c = 0; c < C; ++c) {
                    const float rv = r[p][c];
                    if (!(fabsf(rv) <= 3.0e38f)) nonfinite = 1;
                    const float rc = fminf(fmaxf(rv, 0.f), 1.f);                        // smoe.py:857
                    const float kq = floorf(__fadd_rn(__fmul_rn(rc, a.q_inv_scale), 0.5f));
                    const float rq = __fmul_rn(kq, a.q_scale);                          // smoe.py:899
                    const float tgt = a.image[gidx[p] * C + c];
                    const float diff = __fsub_rn(rq, tgt);                              // smoe.py:905
                    sqsum = fmaf(diff, diff, sqsum);
                    const float ad = fabsf(diff) - a.eps;                               // smoe.py:932
                    lsum[c] = fmaf(ad, ad, lsum[c]);
                    const float cw = a.cfg.use_yuv ? (c == 0 ? 0.75f : 0.125f) : (1.0f / C);
                    const float sgn = diff > 0.f ? 1.f : (diff < 0.f ? -1.f : 0.f);
                    const bool ste = (rv >= 0.f) && (rv <= 1.f);                        // clip + fake-quant STE
                    g[c] = ste ? 2.f * ad * sgn * cw * a.b.inv_count : 0.f;
                    gr = fmaf(g[c], rv, gr);
                    a.res[gidx[p] * C + c] = rq;
                    if (a.res_pre) a.res_pre[gidx[p] * C + c] = rv;
                }
                if (a.argmax) a.argmax[gidx[p]] = bestk[p] >= 0 ? a.indices[bestk[p]] : -1;
            } else {
#pragma unroll
                for (int c = 0; c < C; ++c) g[c] = 0.f;
            }
            if (a.pix) {
                float rec[SMOE_PIXREC];
#pragma unroll
                for (int q = 0; q < SMOE_PIXREC; ++q) rec[q] = 0.f;
                rec[PR_QTHR] = INFINITY;          // pixels outside the batch: w = tau * 2^(-inf) = 0
                if (gidx[p] >= 0) {
#pragma unroll
                    for (int l = 0; l < D; ++l) rec[PR_X + l] = x[p][l];
                    const bool live = S[p] > kSFloor;                                    // smoe.py:821
                    rec[PR_QTHR] = qthr[p];
                    rec[PR_GR] = live ? gr : 0.f;
#pragma unroll
                    for (int c = 0; c < C; ++c) rec[PR_G + c] = g[c];
                }
                float4* dst = reinterpret_cast<float4*>(a.pix + ((size_t)tile * SMOE_TPIX + j) * SMOE_PIXREC);
                dst[0] = make_float4(rec[0], rec[1], rec[2], rec[3]);
                dst[1] = make_float4(rec[4], rec[5], rec[6], rec[7]);
            }
        }
    }

    // ---- loss partials: warp shuffle -> CTA -> fixed-order sum by the last CTA -------------
    float vals[8];
#pragma unroll
    for (int q = 0; q < 8; ++q) vals[q] = 0.f;
#pragma unroll
    for (int c = 0; c < C; ++c) vals[c] = lsum[c];
    vals[4] = sqsum;
    vals[5] = (float)nonfinite;
#pragma unroll
    for (int q = 0; q < 6; ++q)
#pragma unroll
        for (int o = 16; o > 0; o >>= 1) vals[q] += __shfl_down_sync(0xffffffffu, vals[q], o);
    if ((tid & 31) == 0)
#pragma unroll
        for (int q = 0; q < 8; ++q) red[(tid >> 5) * 8 + q] = vals[q];
    __syncthreads();
    __shared__ int s_last;
    if (tid < 8) {
        float s = 0.f;
        for (int wv = 0; wv < kThreadsF / 32; ++wv) s += red[wv * 8 + tid];
        a.partials[(size_t)blockIdx.x * 8 + tid] = s;
    }
    __threadfence();
    __syncthreads();
    if (tid == 0) s_last = (atomicAdd(a.ticket, 1) == (int)gridDim.x - 1);
    __syncthreads();
    if (s_last) {
        __threadfence();
        if (tid < 8) {
            float s = 0.f;
            for (int bq = 0; bq < (int)gridDim.x; ++bq) s += __ldcg(&a.partials[(size_t)bq * 8 + tid]);
            a.scalars[tid] += s;
        }
        if (tid == 0) *a.ticket = 0;
    }
}

template <int D, int C>
static size_t fwd_smem_bytes() {
    return (size_t)(2 * kChunk * pstride(D, C) + kChunk * Rec<D, C>::RC) * 4 + 2 * 8 + 8 * 8 * 4 + 64;
}

}  // namespace smoe

using namespace smoe;

extern "C" int smoe_forward(const smoe_cfg* cfg, const smoe_batch* batch, const float* packed, const int32_t* indices,
                            const int32_t* counts, const float* image, const float* ax0, const float* ax1,
                            const float* ax2, float* res, float* res_pre, int32_t* argmax, uint8_t* infl, float* pix,
                            float* scalars, float* partials, int32_t* ticket, void* stream) {
    SMOE_REQUIRE(cfg && batch && packed && indices && counts && image && ax0 && ax1 && res && scalars && partials &&
                     ticket,
                 "null argument");
    SMOE_REQUIRE(cfg->d == 2 || ax2, "ax2 required for d == 3");
    SMOE_REQUIRE(batch->tile[0] * batch->tile[1] * batch->tile[2] == SMOE_TPIX, "tile product must be SMOE_TPIX");
    for (int i = 0; i < 3; ++i)
        SMOE_REQUIRE(batch->extent[i] > 0 && batch->origin[i] >= 0 && batch->origin[i] + batch->extent[i] <= batch->dims[i],
                     "batch rectangle outside the image");
    SMOE_REQUIRE(cfg->d == 3 || (batch->dims[2] == 1 && batch->tile[2] == 1), "d == 2 needs dims[2] == tile[2] == 1");
    FwdArgs a;
    a.cfg = *cfg;
    a.b = *batch;
    a.packed = packed; a.indices = indices; a.counts = counts; a.image = image;
    a.ax[0] = ax0; a.ax[1] = ax1; a.ax[2] = ax2 ? ax2 : ax0;
    a.res = res; a.res_pre = res_pre; a.argmax = argmax; a.infl = infl; a.pix = pix;
    a.scalars = scalars; a.partials = partials; a.ticket = ticket;
    a.nt1 = (batch->extent[1] + batch->tile[1] - 1) / batch->tile[1];
    a.nt2 = (batch->extent[2] + batch->tile[2] - 1) / batch->tile[2];
    a.ntiles = smoe_num_tiles(batch);
    const float two_p = (float)(1 << cfg->precision);
    a.tau = 0.5f / two_p;
    a.eps = cfg->margin / two_p;
    a.q_scale = 1.0f / (two_p - 1.0f);
    a.q_inv_scale = 1.0f / a.q_scale;
    int dev = 0, sms = 0;
    cudaGetDevice(&dev);
    cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, dev);
    int grid = a.ntiles < 4 * sms ? a.ntiles : 4 * sms;
    cudaStream_t st = (cudaStream_t)stream;
#define CALL(D, C)                                                                                           \
    {                                                                                                        \
        size_t sm = fwd_smem_bytes<D, C>();                                                                  \
        cudaFuncSetAttribute(forward_kernel<D, C>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)sm);    \
        forward_kernel<D, C><<<grid, kThreadsF, sm, st>>>(a);                                                 \
    }
    SMOE_DISPATCH_DC(cfg->d, cfg->C, CALL)
#undef CALL
    return check_launch("smoe_forward");
}
