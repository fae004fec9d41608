// The one exchange step of the pixel-sharded path, written against NVLink peer memory instead of a
// library collective: every rank PUBLISHES its per-kernel sufficient statistics, loss scalars and influence flags in
// a window of its own device memory that all peers have mapped (cudaIpc), and the consumer kernels -- the gradient
// finalisation of a training pass, or a small tail reduction of an evaluation pass -- read the R windows directly
// and sum them in fixed rank order while they work.  One-shot all-reduce fused into its consumer:
//   * no NCCL launch, no pack / unpack kernels, no host synchronisation between the halves of a step: the sharded
//     step is ONE capturable stream of kernels (one CUDA graph), like the single-GPU step;
//   * every rank adds the same R rows in the same order, so gradients, Adam updates and kernel lists are
//     bit-identical on all ranks -- replicas cannot drift apart (the desync guard of SURVEY.md 8e by construction).
// Synchronisation is a flag barrier through the same windows: after publishing epoch e a rank stores e into slot
// [rank] of every peer's flag block (st.release.sys after __threadfence_system), and a consumer CTA spins until all R
// slots of its OWN flag block have reached e (ld.acquire.sys).  Two payload buffers alternate with the parity of e:
// a rank can only publish e+2 after it consumed e+1, which needs every peer to have published e+1, which a peer
// does only after it finished reading everybody's buffer e -- so no second barrier is needed.  A spin that sees no
// progress for ~5 s sets an error flag in the window instead of hanging the GPU.
#pragma once
#include "smoe_common.cuh"

namespace smoe {

// window = [XW_HDR ints: barrier slot 0 {flags[0..8) | epoch 16 | ticket 18}, error 17,
//                       barrier slot 1 {flags[32..40) | epoch 20 | ticket 21}, pad] [payload 0] [payload 1]
// slot 0 is the per-pass statistics exchange, slot 1 the halo pull of a pixel-sharded SSIM loss.
constexpr int XW_HDR = 64;            // ints (256 B)
constexpr int XW_EPOCH = 16, XW_ERROR = 17, XW_TICKET = 18;
__device__ __forceinline__ int xw_flags(int slot) { return slot ? 32 : 0; }
__device__ __forceinline__ int xw_epoch(int slot) { return slot ? 20 : XW_EPOCH; }
__device__ __forceinline__ int xw_ticket(int slot) { return slot ? 21 : XW_TICKET; }
// payload = [K_all*P statistics | SMOE_NSCAL scalars | K_all influence flags | groups reach flags], groups of kGroup
__host__ __device__ inline size_t xw_groups(int K_all) { return ((size_t)K_all + kGroup - 1) / kGroup; }
__host__ __device__ inline size_t xw_payload_floats(int K_all, int P) {
    return ((size_t)K_all * P + SMOE_NSCAL + (size_t)K_all + xw_groups(K_all) + 3) / 4 * 4;
}
__device__ __forceinline__ const float* xw_payload(const smoe_peers& pr, int r, int e, int K_all, int P) {
    return reinterpret_cast<const float*>(reinterpret_cast<const int*>(pr.win[r]) + XW_HDR) +
           (size_t)(e & 1) * xw_payload_floats(K_all, P);
}

__device__ __forceinline__ void st_release_sys(int* p, int v) {
    asm volatile("st.release.sys.global.s32 [%0], %1;" ::"l"(p), "r"(v) : "memory");
}
__device__ __forceinline__ int ld_acquire_sys(const int* p) {
    int v;
    asm volatile("ld.acquire.sys.global.s32 %0, [%1];" : "=r"(v) : "l"(p) : "memory");
    return v;
}
__device__ __forceinline__ float ld_peer(const float* p) {
    float v;
    asm volatile("ld.relaxed.sys.global.f32 %0, [%1];" : "=f"(v) : "l"(p) : "memory");
    return v;
}
__device__ __forceinline__ float4 ld_peer4(const float* p) {
    float4 v;
    asm volatile("ld.relaxed.sys.global.v4.f32 {%0, %1, %2, %3}, [%4];"
                 : "=f"(v.x), "=f"(v.y), "=f"(v.z), "=f"(v.w)
                 : "l"(p)
                 : "memory");
    return v;
}

// Called by every CTA of a consumer kernel.  Block 0 announces this rank's epoch to the peers; thread 0 of every
// block waits for all ranks.  Returns the epoch (its parity selects the payload buffer).
__device__ __forceinline__ int peer_barrier(const smoe_peers& pr, int slot = 0) {
    int* own = reinterpret_cast<int*>(pr.win[pr.rank]);
    const int e = own[xw_epoch(slot)] + 1;
    if (blockIdx.x == 0 && (int)threadIdx.x < pr.world) {
        __threadfence_system();
        st_release_sys(reinterpret_cast<int*>(pr.win[threadIdx.x]) + xw_flags(slot) + pr.rank, e);
    }
    if (threadIdx.x == 0) {
        const long long t0 = clock64();
        for (int r = 0; r < pr.world; ++r) {
            while (ld_acquire_sys(own + xw_flags(slot) + r) - e < 0) {
                if (clock64() - t0 > 10000000000ll) {          // ~5 s at 2 GHz: a peer died; fail loudly, do not hang
                    own[XW_ERROR] = 1;
                    break;
                }
            }
        }
    }
    __syncthreads();
    return e;
}

// the last CTA of a consumer kernel closes the epoch
__device__ __forceinline__ void peer_epoch_end(const smoe_peers& pr, int e, int slot = 0) {
    int* own = reinterpret_cast<int*>(pr.win[pr.rank]);
    __syncthreads();
    if (threadIdx.x == 0) {
        __threadfence();
        if (atomicAdd(own + xw_ticket(slot), 1) == (int)gridDim.x - 1) {
            own[xw_ticket(slot)] = 0;
            own[xw_epoch(slot)] = e;
            __threadfence();
        }
    }
}

// scalars = sum over ranks (rank order), infl = any rank; grid-stride over the tail.  The R remote loads of an
// element are issued back to back before the first use (a peer load is an NVLink round trip of a few microseconds:
// dependent loads would serialise R of them).
__device__ __forceinline__ void reduce_tail(const smoe_peers& pr, int e, int K_all, int P, float* __restrict__ scalars,
                                            uint8_t* __restrict__ infl) {
    const size_t stride = (size_t)K_all * P;
    const size_t step = (size_t)gridDim.x * blockDim.x, i0 = (size_t)blockIdx.x * blockDim.x + threadIdx.x;
    for (size_t i = i0; i < (size_t)SMOE_NSCAL + K_all; i += step) {
        float v[SMOE_MAX_PEERS];
#pragma unroll
        for (int r = 0; r < SMOE_MAX_PEERS; ++r) v[r] = r < pr.world ? ld_peer(xw_payload(pr, r, e, K_all, P) + stride + i) : 0.f;
        float s = 0.f;
#pragma unroll
        for (int r = 0; r < SMOE_MAX_PEERS; ++r) s += v[r];
        if (i < SMOE_NSCAL) scalars[i] = s; else infl[i - SMOE_NSCAL] = s > 0.f ? 1 : 0;
    }
}

// Sum of the R published statistics rows of kernels [k0, k0 + nk) into shared memory (k0 a multiple of kFin), fixed
// rank order, coalesced 16-byte peer loads issued R at a time.  A rank whose pixel block is not reached by a group
// of kernels published a reach flag of 0 for it and its rows are not read at all (they hold nothing defined).
template <int P>
__device__ __forceinline__ void gather_stats(const smoe_peers& pr, int e, int K_all, int k0, int nk,
                                             float* __restrict__ s_stats) {
    constexpr int F4G = kGroup * P / 4;                       // float4 per group: kGroup * P floats is a multiple of 4
    static_assert((kGroup * P) % 4 == 0, "group rows must be float4-aligned");
    __shared__ float s_reach[SMOE_MAX_PEERS][kFin / kGroup];
    const size_t tail = (size_t)K_all * P + SMOE_NSCAL + K_all;
    constexpr int GPB = kFin / kGroup;                         // groups per finalize block
    if (threadIdx.x < SMOE_MAX_PEERS * GPB) {
        const int r = threadIdx.x / GPB, gl = threadIdx.x % GPB;
        const size_t g = (size_t)k0 / kGroup + gl;
        s_reach[r][gl] = (r < pr.world && g < xw_groups(K_all)) ? ld_peer(xw_payload(pr, r, e, K_all, P) + tail + g) : 0.f;
    }
    __syncthreads();
    const size_t beg = (size_t)k0 * P;
    const int n4 = (nk * P + 3) / 4;
    for (int q = threadIdx.x; q < n4; q += blockDim.x) {
        const int gl = q / F4G;
        float4 v[SMOE_MAX_PEERS];
#pragma unroll
        for (int r = 0; r < SMOE_MAX_PEERS; ++r)
            v[r] = (r < pr.world && s_reach[r][gl] != 0.f) ? ld_peer4(xw_payload(pr, r, e, K_all, P) + beg + 4 * (size_t)q)
                                                          : make_float4(0.f, 0.f, 0.f, 0.f);
        float4 s = make_float4(0.f, 0.f, 0.f, 0.f);
#pragma unroll
        for (int r = 0; r < SMOE_MAX_PEERS; ++r) { s.x += v[r].x; s.y += v[r].y; s.z += v[r].z; s.w += v[r].w; }
        const float sv[4] = {s.x, s.y, s.z, s.w};
#pragma unroll
        for (int j = 0; j < 4; ++j)
            if (4 * q + j < nk * P) s_stats[4 * q + j] = sv[j];
    }
    __syncthreads();
}

}  // namespace smoe
