// peer_sync.cuh -- device-side building blocks of the peer-memory exchanges (csrc/peer.cu): window layout, the epoch
// flags (signal with st.release.sys, bounded wait with ld.acquire.sys) and uncached peer loads.  Included by the kernels
// that read other ranks' memory directly (exchange kernels, the gathering pre-pass of the sharded matmul).
#pragma once
#include "feo_common.cuh"

struct feo_comm {
    feo_ctx *ctx = nullptr;
    int rank = 0, world = 1;
    void *windows[FEO_MAX_PEERS] = {nullptr};
    size_t window_bytes = 0, slot_bytes = 0;
    unsigned long long epoch = 0;  // exchanges issued so far
};

namespace feo_peer {

constexpr int kMaxPeers = FEO_MAX_PEERS;
constexpr size_t kFlagStride = 128;
constexpr size_t kFlagBytes = kMaxPeers * kFlagStride;
// How long a kernel waits for a peer is ONE setting for every device-side wait (exchange flags, the sharded matmul's
// k-block flags, and -- with a margin on top -- the matmul's own pipeline barriers): feo_ctx::peer_timeout_ns, from the
// environment variable FEO_PEER_TIMEOUT_S or feo_tune(FEO_KNOB_PEER_TIMEOUT_S) (default 600 s; 0 = wait for ever, like a
// collective library).  A wait that expires raises FEO_STATUS_PEER_TIMEOUT in the context's status word and the kernel goes
// on with whatever data is there: the context stays usable, and the host finds out at feo_peer_status / feo_sync.
constexpr int kStatusPeerTimeout = 1;
constexpr int kXThreads = 256;

struct XchgParams {
    char *win[kMaxPeers];
    int rank, world;
    unsigned long long epoch;
    size_t slot_off;
    long long n;  // outputs
    unsigned long long timeout_ns;  // 0 = no bound
    int *status;                    // device status word of the context (kStatusPeerTimeout on expiry)
};

__device__ __forceinline__ void st_release_sys(unsigned long long *p, unsigned long long v) {
    asm volatile("st.release.sys.global.u64 [%0], %1;" ::"l"(p), "l"(v) : "memory");
}
__device__ __forceinline__ unsigned long long ld_acquire_sys(const unsigned long long *p) {
    unsigned long long v;
    asm volatile("ld.acquire.sys.global.u64 %0, [%1];" : "=l"(v) : "l"(p) : "memory");
    return v;
}
__device__ __forceinline__ unsigned long long global_ns() {
    unsigned long long t;
    asm volatile("mov.u64 %0, %globaltimer;" : "=l"(t));
    return t;
}
// Peer data is read exactly once per exchange; .relaxed.sys keeps it out of the (non-coherent) L1.
template <typename V> __device__ __forceinline__ V ld_peer(const V *p) {
    V r;
    if constexpr (sizeof(V) == 16) {
        uint32_t *w = reinterpret_cast<uint32_t *>(&r);
        asm volatile("ld.relaxed.sys.global.v4.u32 {%0,%1,%2,%3}, [%4];" : "=r"(w[0]), "=r"(w[1]), "=r"(w[2]), "=r"(w[3]) : "l"(p) : "memory");
    } else if constexpr (sizeof(V) == 8) {
        unsigned long long *w = reinterpret_cast<unsigned long long *>(&r);
        asm volatile("ld.relaxed.sys.global.u64 %0, [%1];" : "=l"(w[0]) : "l"(p) : "memory");
    } else {
        static_assert(sizeof(V) == 4, "peer load width");
        uint32_t *w = reinterpret_cast<uint32_t *>(&r);
        asm volatile("ld.relaxed.sys.global.u32 %0, [%1];" : "=r"(w[0]) : "l"(p) : "memory");
    }
    return r;
}

// Steps 1-2 of the header comment.  Every CTA waits on the local flags; only one CTA signals (CTA 0 of the exchange
// kernels; `signals` = true for the one CTA that runs the exchange at the tail of a reduction kernel).
__device__ __forceinline__ void signal_and_wait(const XchgParams &p, bool signals) {
    if (p.epoch == 0) return;  // nothing to wait for (single rank / local data)
    if (threadIdx.x < p.world) {
        if (signals) {
            __threadfence_system();
            st_release_sys(reinterpret_cast<unsigned long long *>(p.win[threadIdx.x] + p.rank * kFlagStride), p.epoch);
        }
        const unsigned long long *mine = reinterpret_cast<const unsigned long long *>(p.win[p.rank] + threadIdx.x * kFlagStride);
        const unsigned long long t0 = global_ns();
        while (ld_acquire_sys(mine) < p.epoch) {
            if (p.timeout_ns != 0 && global_ns() - t0 > p.timeout_ns) {
                if (p.status && atomicOr(p.status, kStatusPeerTimeout) == 0)
                    printf("feo peer exchange: rank %d gave up waiting for rank %d to reach exchange %llu (it is at %llu); result invalid\n",
                           p.rank, (int)threadIdx.x, p.epoch, ld_acquire_sys(mine));
                break;
            }
            __nanosleep(32);
        }
    }
    __syncthreads();
}

__device__ __forceinline__ void signal_and_wait(const XchgParams &p) { signal_and_wait(p, blockIdx.x == 0 && blockIdx.y == 0); }

// A sharded reduction asks the local reduction kernel to run the exchange itself (feo_reduce_sharded -> ctx->xtail_req):
// when the plan is one launch whose outputs are final (rows kernel, un-split or fused combine) and few, the last CTA to
// finish signals the peers, waits, pulls and combines -- one launch instead of two.  `consumed` tells the caller whether the
// planner took the offer; if not it launches the exchange kernel as before.
struct XTailReq {
    bool consumed;
    int mean;                                  // Tensor::mean: divide the sum by `divisor`
    XchgParams p;
    void *out;                                 // the replicated result
    const void *slot;                          // where the final local pass must write its partial (this rank's window slot)
    unsigned long long divisor_bits, count_bits;  // values of the element type, as bit patterns
};

// B[K,N] of the sharded matmul as `world` row panels (panel r = rows [r * rows_per_part, ...)), each a peer-mapped pointer.
struct BParts {
    const float *part[kMaxPeers];
    int world;
    int rows_per_part;
};

}  // namespace feo_peer

// matmul_tf32x3.cu: C = A * B with B given as row panels in peer memory; the pre-pass waits on `sync` (epoch 0 = no wait)
// and pulls the panels over NVLink while it splits and transposes them.
int feo_impl_matmul_tf32x3_parts(feo_ctx *ctx, void *C, const void *A, const feo_peer::BParts &parts, const feo_peer::XchgParams &sync,
                                 size_t M, size_t K, size_t N);
