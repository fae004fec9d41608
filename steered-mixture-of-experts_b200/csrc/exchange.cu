// Host entry points and stand-alone kernels of the peer-memory exchange (see exchange.cuh for the protocol).
#include "exchange.cuh"

namespace smoe {

// payload[e & 1] = [sum_s raw_part[s] (K*P, fixed order) | scalars | influence flags as 0/1 floats]
__global__ void __launch_bounds__(256) xchg_publish_kernel(smoe_peers pr, const int32_t* __restrict__ counts, int K_all,
                                                           int P, int num_splits, const float* __restrict__ part,
                                                           const int32_t* __restrict__ plan_cnt,
                                                           const float* __restrict__ scalars,
                                                           const uint8_t* __restrict__ infl) {
    int* own = reinterpret_cast<int*>(pr.win[pr.rank]);
    const int e = own[XW_EPOCH] + 1;
    float* pay = reinterpret_cast<float*>(own + XW_HDR) + (size_t)(e & 1) * xw_payload_floats(K_all, P);
    const size_t stride = (size_t)K_all * P;
    const size_t step = (size_t)gridDim.x * 256, i0 = (size_t)blockIdx.x * 256 + threadIdx.x;
    if (part) {
        const size_t n = (size_t)counts[0] * P;
        for (size_t i = i0; i < n; i += step) {
            const int used = plan_cnt ? segments_used(plan_cnt[(i / P) / kGroup], num_splits) : num_splits;
            if (used == 0) continue;                       // unreached group: reach flag 0, rows never read
            pay[i] = slab_sum(part + i, stride, used);
        }
    }
    float* tail = pay + stride;
    for (size_t i = i0; i < (size_t)SMOE_NSCAL + K_all; i += step)
        tail[i] = i < SMOE_NSCAL ? scalars[i] : (infl[i - SMOE_NSCAL] ? 1.f : 0.f);
    // reach flags: which groups of kernels this rank's backward had any tile for (their rows above are defined)
    float* reach = tail + SMOE_NSCAL + K_all;
    for (size_t g = i0; g < xw_groups(K_all); g += step)
        reach[g] = (part && (!plan_cnt || plan_cnt[g] > 0) && (int)(g * kGroup) < counts[0]) ? 1.f : 0.f;
}

__global__ void __launch_bounds__(256) xchg_reduce_tail_kernel(smoe_peers pr, int K_all, int P,
                                                               float* __restrict__ scalars, uint8_t* __restrict__ infl) {
    const int e = peer_barrier(pr);
    reduce_tail(pr, e, K_all, P, scalars, infl);
    peer_epoch_end(pr, e);
}

// Halo pull of a pixel-sharded SSIM loss: after every rank's loss stage has written the quantised reconstruction of
// its own block, each rank copies the ring of pixels around its block from the ranks that own them (peer loads over
// NVLink), so that SSIM windows which straddle a block border see the neighbour's pixels.  One flag barrier (slot 1)
// in front; the next writer of any res buffer is the next pass's loss stage, which every rank reaches only after the
// statistics exchange of this pass -- i.e. after everybody's pull -- so no second barrier is needed.
__global__ void __launch_bounds__(256) halo_pull_kernel(smoe_peers pr, smoe_halo_map hm) {
    const int e = peer_barrier(pr, 1);
    const int me = hm.rank;
    const size_t n = (size_t)hm.buf_dims[me][0] * hm.buf_dims[me][1] * hm.buf_dims[me][2];
    float* own = reinterpret_cast<float*>(hm.res[me]);
    for (size_t i = (size_t)blockIdx.x * 256 + threadIdx.x; i < n; i += (size_t)gridDim.x * 256) {
        int g[3];
        g[2] = (int)(i % hm.buf_dims[me][2]) + hm.buf_lo[me][2];
        g[1] = (int)((i / hm.buf_dims[me][2]) % hm.buf_dims[me][1]) + hm.buf_lo[me][1];
        g[0] = (int)(i / ((size_t)hm.buf_dims[me][2] * hm.buf_dims[me][1])) + hm.buf_lo[me][0];
        bool mine = true;
        for (int a = 0; a < 3; ++a) mine = mine && g[a] >= hm.blk_lo[me][a] && g[a] < hm.blk_hi[me][a];
        if (mine) continue;
        for (int r = 0; r < hm.world; ++r) {
            bool in = true;
            for (int a = 0; a < 3; ++a) in = in && g[a] >= hm.blk_lo[r][a] && g[a] < hm.blk_hi[r][a];
            if (!in) continue;
            const size_t src = (((size_t)(g[0] - hm.buf_lo[r][0]) * hm.buf_dims[r][1] + (g[1] - hm.buf_lo[r][1])) *
                                hm.buf_dims[r][2] + (g[2] - hm.buf_lo[r][2])) * hm.C;
            const float* peer = reinterpret_cast<const float*>(hm.res[r]);
            for (int c = 0; c < hm.C; ++c) own[i * hm.C + c] = ld_peer(peer + src + c);
            break;
        }
    }
    peer_epoch_end(pr, e, 1);
}

}  // namespace smoe

using namespace smoe;

extern "C" {

size_t smoe_xchg_window_bytes(int K_all, int P) {
    return (size_t)XW_HDR * sizeof(int32_t) + 2 * xw_payload_floats(K_all, P) * sizeof(float);
}

// The window must be a dedicated cudaMalloc allocation (cudaIpc exports whole allocations), so the library
// allocates it: the one documented exception to "the library never allocates".
int smoe_peer_alloc(size_t bytes, void** ptr) {
    SMOE_REQUIRE(ptr && bytes > 0, "bad argument");
    cudaError_t e = cudaMalloc(ptr, bytes);
    if (e == cudaSuccess) e = cudaMemset(*ptr, 0, bytes);
    if (e == cudaSuccess) e = cudaDeviceSynchronize();
    if (e != cudaSuccess) { set_error("smoe_peer_alloc: %s", cudaGetErrorString(e)); return (int)e; }
    return 0;
}
int smoe_peer_free(void* ptr) {
    cudaError_t e = cudaFree(ptr);
    if (e != cudaSuccess) { set_error("smoe_peer_free: %s", cudaGetErrorString(e)); return (int)e; }
    return 0;
}
int smoe_peer_export(const void* ptr, void* handle64) {
    SMOE_REQUIRE(ptr && handle64, "bad argument");
    static_assert(sizeof(cudaIpcMemHandle_t) == 64, "handle size");
    cudaError_t e = cudaIpcGetMemHandle((cudaIpcMemHandle_t*)handle64, const_cast<void*>(ptr));
    if (e != cudaSuccess) { set_error("smoe_peer_export: %s", cudaGetErrorString(e)); return (int)e; }
    return 0;
}
int smoe_peer_open(const void* handle64, void** ptr) {
    SMOE_REQUIRE(ptr && handle64, "bad argument");
    cudaIpcMemHandle_t h;
    memcpy(&h, handle64, sizeof(h));
    cudaError_t e = cudaIpcOpenMemHandle(ptr, h, cudaIpcMemLazyEnablePeerAccess);
    if (e != cudaSuccess) { set_error("smoe_peer_open: %s", cudaGetErrorString(e)); return (int)e; }
    return 0;
}
int smoe_peer_close(void* ptr) {
    cudaError_t e = cudaIpcCloseMemHandle(ptr);
    if (e != cudaSuccess) { set_error("smoe_peer_close: %s", cudaGetErrorString(e)); return (int)e; }
    return 0;
}

static int check_peers(const smoe_peers* pr) {
    if (!pr || pr->world < 1 || pr->world > SMOE_MAX_PEERS || pr->rank < 0 || pr->rank >= pr->world) return 0;
    for (int r = 0; r < pr->world; ++r)
        if (!pr->win[r]) return 0;
    return 1;
}

int smoe_xchg_publish(const smoe_cfg* cfg, const smoe_peers* peers, const int32_t* counts, int K_all, int num_splits,
                      const float* raw_part, const int32_t* plan, const float* scalars, const uint8_t* infl,
                      void* stream) {
    SMOE_REQUIRE(cfg && counts && scalars && infl && K_all > 0 && num_splits > 0, "bad argument");
    SMOE_REQUIRE(check_peers(peers), "bad peer set");
    const int P = nparam(cfg->d, cfg->C);
    size_t n = (size_t)K_all * P;
    int nb = (int)((n + 255) / 256);
    if (nb > 148 * 8) nb = 148 * 8;
    xchg_publish_kernel<<<nb, 256, 0, (cudaStream_t)stream>>>(*peers, counts, K_all, P, num_splits, raw_part, plan, scalars, infl);
    return check_launch("smoe_xchg_publish");
}

int smoe_xchg_reduce_tail(const smoe_cfg* cfg, const smoe_peers* peers, int K_all, float* scalars, uint8_t* infl,
                          void* stream) {
    SMOE_REQUIRE(cfg && scalars && infl && K_all > 0, "bad argument");
    SMOE_REQUIRE(check_peers(peers), "bad peer set");
    int nb = (K_all + SMOE_NSCAL + 255) / 256;
    if (nb > 148) nb = 148;             // every CTA spins in the barrier: all of them must be resident
    xchg_reduce_tail_kernel<<<nb, 256, 0, (cudaStream_t)stream>>>(*peers, K_all, nparam(cfg->d, cfg->C), scalars, infl);
    return check_launch("smoe_xchg_reduce_tail");
}

// Host -> device feed of a pass's target pixels on a dedicated copy stream: the copy is ordered after everything
// already enqueued on `main_stream` (earlier readers of dst) and `done_event` fires when it has landed; the consumer
// stream waits for that event right before smoe_loss.  Four runtime calls, no synchronisation.
int smoe_feed(void* dst, const void* src_host, size_t bytes, void* main_stream, void* copy_stream, void* order_event,
              void* done_event) {
    SMOE_REQUIRE(dst && src_host && bytes > 0 && copy_stream && order_event && done_event, "bad argument");
    cudaError_t e = cudaEventRecord((cudaEvent_t)order_event, (cudaStream_t)main_stream);
    if (e == cudaSuccess) e = cudaStreamWaitEvent((cudaStream_t)copy_stream, (cudaEvent_t)order_event, 0);
    if (e == cudaSuccess) e = cudaMemcpyAsync(dst, src_host, bytes, cudaMemcpyHostToDevice, (cudaStream_t)copy_stream);
    if (e == cudaSuccess) e = cudaEventRecord((cudaEvent_t)done_event, (cudaStream_t)copy_stream);
    if (e != cudaSuccess) { set_error("smoe_feed: %s", cudaGetErrorString(e)); return (int)e; }
    return 0;
}

int smoe_halo_pull(const smoe_peers* peers, const smoe_halo_map* map, void* stream) {
    SMOE_REQUIRE(check_peers(peers) && map, "bad argument");
    SMOE_REQUIRE(map->world == peers->world && map->rank == peers->rank && (map->C == 1 || map->C == 3), "bad halo map");
    for (int r = 0; r < map->world; ++r) SMOE_REQUIRE(map->res[r], "null res buffer in the halo map");
    const size_t n = (size_t)map->buf_dims[map->rank][0] * map->buf_dims[map->rank][1] * map->buf_dims[map->rank][2];
    int nb = (int)((n + 255) / 256);
    if (nb > 148 * 4) nb = 148 * 4;
    halo_pull_kernel<<<nb, 256, 0, (cudaStream_t)stream>>>(*peers, *map);
    return check_launch("smoe_halo_pull");
}

int smoe_xchg_status(const smoe_peers* peers, int32_t* epoch_and_error /*[2], host*/) {
    SMOE_REQUIRE(check_peers(peers) && epoch_and_error, "bad argument");
    cudaError_t e = cudaMemcpy(epoch_and_error, (const int32_t*)peers->win[peers->rank] + XW_EPOCH, 2 * sizeof(int32_t),
                               cudaMemcpyDeviceToHost);
    if (e != cudaSuccess) { set_error("smoe_xchg_status: %s", cudaGetErrorString(e)); return (int)e; }
    return 0;
}

}  // extern "C"
