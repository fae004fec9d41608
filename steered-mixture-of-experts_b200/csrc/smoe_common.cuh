// Shared device/host helpers of libsmoe_b200 (sm_100a only).
#pragma once
#include <cuda_runtime.h>
#include <stdint.h>
#include <stdio.h>
#include <math.h>
#include "../../include/smoe_b200.h"

#if defined(__CUDA_ARCH__) && (__CUDA_ARCH__ < 1000)
#error "libsmoe_b200 is written for sm_100a (B200) only"
#endif

namespace smoe {

#ifndef SMOE_BWD_THREADS
#define SMOE_BWD_THREADS 64
#endif
#ifndef SMOE_BWD_HALVES
#define SMOE_BWD_HALVES 2
#endif
constexpr int kThreads = SMOE_BWD_THREADS;                        // threads per CTA of the backward
constexpr int kHalves = SMOE_BWD_HALVES;                          // threads per kernel in the backward: each takes every other tile row
constexpr int kGroup = kThreads / kHalves;          // kernels per backward CTA (= planning group)
constexpr int kFin = 64;                            // kernels per CTA of grad_finalize (256 threads gather, 64 apply the chain rule)
constexpr int kThreadsF = 128;                       // threads per CTA of the forward
constexpr int kPixPerThread = SMOE_TPIX / kThreadsF; // 4 pixels per thread in the forward
constexpr int kChunk = 128;                          // kernels staged per shared-memory chunk
constexpr float kHalfLog2e = 0.72134752044448170368f;   // log2(e)/2
constexpr float kSFloor = 10e-12f;                  // the literal of smoe.py:821

__host__ __device__ constexpr int tri(int d) { return d * (d + 1) / 2; }
__host__ __device__ constexpr int nparam(int d, int C) { return d + tri(d) + 1 + C + d * C; }
// packed record = P parameters | lam (eigenvalue bound) | kap[d] (per-axis bounds) | pad to a multiple of 4
__host__ __device__ constexpr int pstride(int d, int C) { return (nparam(d, C) + 1 + d + 3) / 4 * 4; }
constexpr int kCB = 12;                              // floats per chunk-bounds entry
// Groups of kernels that reach no tile of the batch are skipped by every backward CTA; their slabs of raw_part are
// never written and never read.  A group that reaches some tile has all its num_splits slabs written.
__host__ __device__ inline int segments_used(int n, int num_splits) { return n > 0 ? num_splits : 0; }
// offsets inside a theta / grads row and inside a packed record (same order)
__host__ __device__ constexpr int off_mu(int, int) { return 0; }
__host__ __device__ constexpr int off_A(int d, int) { return d; }
__host__ __device__ constexpr int off_pi(int d, int) { return d + tri(d); }
__host__ __device__ constexpr int off_nu(int d, int) { return d + tri(d) + 1; }
__host__ __device__ constexpr int off_ga(int d, int C) { return d + tri(d) + 1 + C; }
// lower-triangular (l >= m) row-major index; upper-triangular (l <= m) row-major index
__host__ __device__ constexpr int lt(int l, int m) { return l * (l + 1) / 2 + m; }
__host__ __device__ constexpr int ut(int d, int l, int m) { return l * d - l * (l - 1) / 2 + (m - l); }

// Pixel state written by the forward and streamed by the backward, per tile:
//   planes [3 + C][SMOE_TPIX]: z = tile-centred LAST coordinate | qthr = log2(max(S, 1e-11)), +inf
//   outside the batch | gr = sum_c g_c r_c (0 where S is clamped, smoe.py:821) | g_c = dL/dr_c
//   then rowc [d-1][SMOE_TPIX / RL]: the other tile-centred coordinates, constant along a row of RL pixels
constexpr int PL_Z = 0, PL_QTHR = 1, PL_GR = 2, PL_G = 3;
__host__ __device__ constexpr int pix_rowc_offset(int C) { return (3 + C) * SMOE_TPIX; }
__host__ __device__ inline int pix_stride(int d, int C, int RL) {
    return pix_rowc_offset(C) + ((d - 1) * (SMOE_TPIX / RL) + 3) / 4 * 4;
}

// TF FakeQuantWithMinMaxArgs / ...Vars (float32): nudged range and scale, see oracle/graph.py:_nudge.
//   value = fq(x - shift; nmin, nmax) + shift   (shift = 0 except for the "shifted" groups of mode 3)
//   flags: QF_ZERO  min == max == 0: the op returns zeros and passes the gradient (TF kernel special case)
//          QF_PASS  straight-through for every element (the shifted form of mode 3, smoe.py:506-511, 524-527)
//          QF_IDENT the group is not quantised at all (musX with train_musx == False in mode 3, smoe.py:516-522)
//          QF_ROUTE mode-3 plain group: the in-range mask is applied by smoe_quant_route, not by grad_finalize
struct Nudged { float nmin, nmax, scale, inv_scale, shift; int flags; };
constexpr int QF_ZERO = 1, QF_PASS = 2, QF_IDENT = 4, QF_ROUTE = 8;
constexpr int QG_AD = 0, QG_MU = 1, QG_NU = 2, QG_PI = 3, QG_GA = 4, QG_AC = 5;   // A diagonal | musX | nu_e | pis | gamma_e | A_corr
struct QuantSet { Nudged g[6]; int mode; int pad; };
struct QuantDyn { QuantSet qs; float mn[6], mx[6]; };       // mode 3: ranges computed on the device (smoe_quant_ranges)
__host__ __device__ inline Nudged nudge(float mn, float mx, int bits) {
    float qmin = 0.f, qmax = (float)((1 << bits) - 1);
    float scale = (mx - mn) / (qmax - qmin);
    float zp = qmin - mn / scale;
    float nzp = zp < qmin ? qmin : (zp > qmax ? qmax : floorf(zp + 0.5f));
    Nudged n;
    n.nmin = (qmin - nzp) * scale;
    n.nmax = (qmax - nzp) * scale;
    n.scale = scale;
    n.inv_scale = 1.0f / scale;
    n.shift = 0.f;
    n.flags = 0;
    return n;
}
static inline QuantSet make_quantset(const smoe_cfg* cfg) {
    QuantSet q;
    q.mode = cfg->quantization_mode;
    q.pad = 0;
    for (int i = 0; i < 6; ++i) {
        q.g[i].nmin = 0.f; q.g[i].nmax = 0.f; q.g[i].scale = 1.f; q.g[i].inv_scale = 1.f; q.g[i].shift = 0.f;
        q.g[i].flags = QF_IDENT;
        const int b = i == QG_AC ? 0 : i;                    // A_corr shares the bounds of A (smoe.py:483-486)
        if (cfg->quantization_mode == 2) q.g[i] = nudge(cfg->q_lb[b], cfg->q_ub[b], cfg->q_bits[b]);
    }
    if (cfg->quantization_mode != 2 && cfg->quantize_pis) q.g[QG_PI] = nudge(cfg->pis_lb, cfg->pis_ub, cfg->pis_bits);
    return q;
}

void set_error(const char* fmt, ...);
int check_launch(const char* what);

#define SMOE_REQUIRE(cond, msg)                           \
    do {                                                  \
        if (!(cond)) {                                    \
            smoe::set_error("%s: %s", __func__, msg);     \
            return SMOE_E_BADARG;                         \
        }                                                 \
    } while (0)

// dispatch on (d, C): instantiations exist for d in {2,3}, C in {1,3}
#define SMOE_DISPATCH_DC(d, C, CALL)                                          \
    if ((d) == 2 && (C) == 1) { CALL(2, 1); }                                 \
    else if ((d) == 2 && (C) == 3) { CALL(2, 3); }                            \
    else if ((d) == 3 && (C) == 1) { CALL(3, 1); }                            \
    else if ((d) == 3 && (C) == 3) { CALL(3, 3); }                            \
    else { smoe::set_error("%s: unsupported (d,C)=(%d,%d)", __func__, (int)(d), (int)(C)); return SMOE_E_UNSUPPORTED; }

#ifdef __CUDACC__
__device__ __forceinline__ float fake_quant(float x, Nudged n) {
    if (n.flags & QF_IDENT) return x;
    if (n.flags & QF_ZERO) return n.shift;
    float c = fminf(fmaxf(__fsub_rn(x, n.shift), n.nmin), n.nmax);
    float k = floorf(__fadd_rn(__fmul_rn(__fsub_rn(c, n.nmin), n.inv_scale), 0.5f));
    return __fadd_rn(__fadd_rn(__fmul_rn(k, n.scale), n.nmin), n.shift);
}
__device__ __forceinline__ bool quant_in_range(float x, Nudged n) {
    const float xs = __fsub_rn(x, n.shift);
    return xs >= n.nmin && xs <= n.nmax;
}
__device__ __forceinline__ float ste_mask(float x, Nudged n) {
    if (n.flags & (QF_IDENT | QF_ZERO | QF_PASS | QF_ROUTE)) return 1.f;
    return quant_in_range(x, n) ? 1.f : 0.f;
}
// sum_{sp < used} p[sp * stride] in split order; the loads of 8 slabs are issued before the first add (a dependent
// load-add chain would serialise one memory latency per slab)
__device__ __forceinline__ float slab_sum(const float* __restrict__ p, size_t stride, int used) {
    float acc = 0.f;
    for (int sp = 0; sp < used; sp += 8) {
        float v[8];
#pragma unroll
        for (int u = 0; u < 8; ++u) v[u] = (sp + u < used) ? p[(size_t)(sp + u) * stride] : 0.f;
#pragma unroll
        for (int u = 0; u < 8; ++u)
            if (sp + u < used) acc += v[u];
    }
    return acc;
}
__device__ __forceinline__ float ex2f(float x) {
    float y;
    asm("ex2.approx.ftz.f32 %0, %1;" : "=f"(y) : "f"(x));
    return y;
}
__device__ __forceinline__ uint32_t smem_u32(const void* p) {
    return (uint32_t)__cvta_generic_to_shared(p);
}
// ---- mbarrier + TMA bulk copy (cp.async.bulk, SASS UBLKCP) ---------------------------------
__device__ __forceinline__ void mbar_init(uint64_t* bar, uint32_t count) {
    asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(smem_u32(bar)), "r"(count));
}
__device__ __forceinline__ void fence_barrier_init() {
    asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
}
__device__ __forceinline__ void fence_proxy_async() {
    asm volatile("fence.proxy.async.shared::cta;" ::: "memory");
}
__device__ __forceinline__ void mbar_expect_tx(uint64_t* bar, uint32_t bytes) {
    asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(smem_u32(bar)), "r"(bytes)
                 : "memory");
}
__device__ __forceinline__ void mbar_wait(uint64_t* bar, uint32_t parity) {
    asm volatile(
        "{\n"
        ".reg .pred p;\n"
        "WAIT_LOOP:\n"
        "mbarrier.try_wait.parity.shared::cta.b64 p, [%0], %1;\n"
        "@p bra WAIT_DONE;\n"
        "bra WAIT_LOOP;\n"
        "WAIT_DONE:\n"
        "}\n" ::"r"(smem_u32(bar)),
        "r"(parity)
        : "memory");
}
// global -> shared bulk copy; bytes % 16 == 0, both addresses 16-byte aligned
__device__ __forceinline__ void tma_load_1d(void* smem_dst, const void* gsrc, uint32_t bytes, uint64_t* bar) {
    asm volatile(
        "cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1], %2, [%3];" ::"r"(
            smem_u32(smem_dst)),
        "l"(gsrc), "r"(bytes), "r"(smem_u32(bar))
        : "memory");
}
#endif  // __CUDACC__

}  // namespace smoe
