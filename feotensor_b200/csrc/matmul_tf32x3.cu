// matmul_tf32x3.cu -- K7: DynamicMatrix::matmul (matrix.rs:76-90) for f32 on the 5th-generation tensor cores.
//
//   C[M,N] = A[M,K] * B[K,N]  (all dense row-major f32), computed as split-TF32 ("3xTF32"):
//       a = a_hi + a_lo,  a_hi = rna_tf32(a),  a_lo = a - a_hi   (exact in f32; the MMA reads a_lo's top 19 bits)
//       C = A_lo*B_hi + A_hi*B_lo + A_hi*B_hi                    (small terms first; A_lo*B_lo ~ 2^-22 is dropped)
//   with f32 accumulation in tensor memory.  Error ~ 1e-6 * sum|a||b|: inside the f32 reassociation bound the
//   reference's own sequential sum obeys, at 3 tensor-core passes instead of the FP32 SIMT pipe.
//
// Structure (hand-written PTX; no CUTLASS).  Two main kernels share the pre-pass over A and the accumulation scheme:
//   pre-pass   split_rows_kernel     A  -> A_hi, A_lo          [M,Kp] (K-major as is, zero padded to whole k-blocks)
//   default    matmul_tf32x3_pair_kernel (further down): tcgen05 cta_group::2, two CTAs per 256 x BN tile, persistent,
//              B read as raw f32 tiles (MN-major operand) and split hi/lo inside the kernel -- no pre-pass over B;
//              for a row-sharded B the panels arrive while the kernel runs (copy engines / the kernel's own TMA bulk
//              copies / a gather kernel, see matmul_tf32x3_run).
//   fallback   matmul_tf32x3_kernel (round 1; knob MATMUL_TF32_PAIR = 1): one 128 x 256 output tile per CTA, with
//              split_transpose_kernel B -> Bt_hi, Bt_lo [N,Kp] (transposed so both operands are K-major) in front
//              (the pair kernel takes this pre-split form too when N % 4 != 0: TMA needs a 16-byte row pitch):
//              warp 0   TMA producer: cp.async.bulk.tensor (128B swizzle) of A_hi/A_lo/Bt_hi/Bt_lo k-blocks into a
//                       kStages-deep shared-memory ring, completion on mbarriers (expect_tx)
//              warp 1   allocates all 512 TMEM columns (two 128x256 f32 accumulators); one lane issues
//                       tcgen05.mma.cta_group::1.kind::tf32 (M=128, N=256, K=8; 12 per k-block) from shared-memory
//                       descriptors; tcgen05.commit frees the ring slot / publishes a finished chunk
//              warps 2-9 accumulate + epilogue.  The tensor core TRUNCATES when it adds into the TMEM accumulator
//                       (measured: 83 % of the errors point toward zero and the bias grows linearly with K,
//                       -1.7e-4 relative at K = 16384 on positive data), so K is cut into chunks of kc k-blocks that
//                       accumulate into alternating TMEM buffers; these warps drain each finished chunk with
//                       tcgen05.ld and add it to f32 registers with round-to-nearest, overlapped with
//                       the MMAs of the next chunk, then write C.
// Bound: the tensor pipe (3 x 2MNK TF32 flop; 92.7 % active in the pair kernel at 8192^3, at the clock the power cap
// allows); see DESIGN.md 4.6 for the measurements.
#include <cuda.h>

#include "feo_common.cuh"
#include "peer_sync.cuh"

namespace {

constexpr int BM = 128, BN = 256, BK = 32;      // CTA tile; BK floats = 128 bytes = one swizzle atom row
constexpr int UMMA_K = 8;                        // tf32: 32 bytes of K per instruction
constexpr int kStages = 2;
constexpr int kTileABytes = BM * BK * 4;         // 16 KiB
constexpr int kTileBBytes = BN * BK * 4;         // 32 KiB
constexpr int kStageBytes = 2 * kTileABytes + 2 * kTileBBytes;  // hi + lo of both operands: 96 KiB
constexpr int kSmemBytes = kStages * kStageBytes + 1024 /* alignment slack */ + 256 /* barriers */;
constexpr int kThreads = 320;                    // warp 0 TMA, warp 1 MMA, warps 2..9 accumulate + epilogue
constexpr int kTmemCols = 512;                   // two accumulator buffers of BN columns
constexpr int kEpiWarps = 8;
constexpr int kGroupM = 16;                      // rasterisation group (tiles along M)

__device__ __forceinline__ uint32_t smem_u32(const void *p) { return (uint32_t)__cvta_generic_to_shared(p); }

__device__ __forceinline__ void mbar_init(uint64_t *bar, uint32_t count) {
    asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(smem_u32(bar)), "r"(count));
}
__device__ __forceinline__ void mbar_arrive(uint64_t *bar) {
    asm volatile("mbarrier.arrive.shared::cta.b64 _, [%0];" ::"r"(smem_u32(bar)) : "memory");
}
__device__ __forceinline__ void mbar_expect_tx(uint64_t *bar, uint32_t bytes) {
    asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(smem_u32(bar)), "r"(bytes) : "memory");
}
// Bounded wait (limit_ns = 0: unbounded): the pipeline's barriers complete within microseconds unless the producer is itself
// waiting for a peer's k-blocks, and that wait gives up after the context's peer timeout -- so a barrier still closed after
// twice that plus a margin is a pipeline bug, and the only sane reaction is to report and trap instead of hanging the GPU.
__device__ __forceinline__ void mbar_wait(uint64_t *bar, uint32_t parity, int what, unsigned long long limit_ns) {
    const uint32_t addr = smem_u32(bar);
    unsigned long long t_start = 0;
    for (uint32_t spin = 0;; ++spin) {
        uint32_t done;
        asm volatile(
            "{\n\t.reg .pred p;\n\t"
            "mbarrier.try_wait.parity.shared::cta.b64 p, [%1], %2;\n\t"
            "selp.u32 %0, 1, 0, p;\n\t}"
            : "=r"(done)
            : "r"(addr), "r"(parity)
            : "memory");
        if (done) return;
        // wall time, sampled every 2^20 polls
        if (limit_ns != 0 && (spin & 0xFFFFFu) == 0xFFFFFu) {
            if (t_start == 0) t_start = feo_peer::global_ns();
            else if (feo_peer::global_ns() - t_start > limit_ns) {
                printf("matmul_tf32x3: mbarrier wait timed out (what=%d block=%d thread=%d parity=%u)\n", what, blockIdx.x, threadIdx.x, parity);
                __trap();
            }
        }
    }
}
__device__ __forceinline__ void tma_load_2d(void *smem_dst, const CUtensorMap *map, uint64_t *bar, int c0, int c1) {
    asm volatile(
        "cp.async.bulk.tensor.2d.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1, {%3, %4}], [%2];"
        ::"r"(smem_u32(smem_dst)), "l"(map), "r"(smem_u32(bar)), "r"(c0), "r"(c1)
        : "memory");
}
// K-major, 128-byte swizzle shared-memory matrix descriptor (rows of 128 bytes, 8-row groups 1024 bytes apart).
__device__ __forceinline__ uint64_t make_desc(uint32_t smem_addr) {
    uint64_t d = 0;
    d |= (uint64_t)((smem_addr & 0x3FFFF) >> 4);        // start address, bits [0,14)
    d |= (uint64_t)1 << 16;                              // leading byte offset (unused for swizzled K-major)
    d |= (uint64_t)(1024 >> 4) << 32;                    // stride byte offset: next 8-row group
    d |= (uint64_t)1 << 46;                              // descriptor version (Blackwell)
    d |= (uint64_t)2 << 61;                              // SWIZZLE_128B
    return d;
}
// MN-major descriptor for a raw tile of B[K,N].  For 32-bit operands the tensor core transposes MN-major data only from the
// "128-byte swizzle with 32-byte atoms" layout (UMMA layout type 1, TMA's CU_TENSOR_MAP_SWIZZLE_128B_ATOM_32B): rows of 32
// consecutive n (128 bytes), the four 32-byte chunks of a row XOR-ed with (k mod 4), i.e. a 4-row, 512-byte swizzle atom.
// Canonical layout ((8,n),(4,k)):((1,LBO),(8,SBO)) in 16-byte units: leading byte offset = stride between 32-column groups
// (one TMA box, P_BOX_BYTES), stride byte offset = stride between 4-row k groups (512); one UMMA_K step (8 k) = two atoms,
// so the start address advances 1024 bytes per step.
__device__ __forceinline__ uint64_t make_desc_mn(uint32_t smem_addr, uint32_t group_stride_bytes) {
    uint64_t d = 0;
    d |= (uint64_t)((smem_addr & 0x3FFFF) >> 4);
    d |= (uint64_t)(group_stride_bytes >> 4) << 16;      // leading byte offset: next 32 columns
    d |= (uint64_t)(512 >> 4) << 32;                     // stride byte offset: next 4 k-rows
    d |= (uint64_t)1 << 46;
    d |= (uint64_t)1 << 61;                              // SWIZZLE_128B_BASE32B
    return d;
}
// kind::tf32, f32 accumulate, both operands K-major, M = 128, N = 256.
constexpr uint32_t kIdesc = (1u << 4) | (2u << 7) | (2u << 10) | ((uint32_t)(BN >> 3) << 17) | ((uint32_t)(BM >> 4) << 24);

__device__ __forceinline__ void umma_tf32(uint32_t tmem_d, uint64_t desc_a, uint64_t desc_b, uint32_t accumulate) {
    asm volatile(
        "{\n\t.reg .pred p;\n\t"
        "setp.ne.b32 p, %4, 0;\n\t"
        "tcgen05.mma.cta_group::1.kind::tf32 [%0], %1, %2, %3, p;\n\t}"
        ::"r"(tmem_d), "l"(desc_a), "l"(desc_b), "r"(kIdesc), "r"(accumulate)
        : "memory");
}
__device__ __forceinline__ void umma_commit(uint64_t *bar) {
    asm volatile("tcgen05.commit.cta_group::1.mbarrier::arrive::one.shared::cluster.b64 [%0];" ::"r"(smem_u32(bar)) : "memory");
}

__global__ void __launch_bounds__(kThreads, 1)
matmul_tf32x3_kernel(const __grid_constant__ CUtensorMap map_a_hi, const __grid_constant__ CUtensorMap map_a_lo,
                     const __grid_constant__ CUtensorMap map_b_hi, const __grid_constant__ CUtensorMap map_b_lo,
                     float *__restrict__ C, int M, int N, int num_k_blocks, int kc, const unsigned int *__restrict__ ready,
                     unsigned int ready_need, int kb_rot, unsigned long long peer_timeout_ns, int *status) {
    const unsigned long long bar_limit = peer_timeout_ns ? 2 * peer_timeout_ns + 30000000000ull : 0;
    // `ready` != nullptr: Bt's k-blocks are still being produced by the gathering pre-pass running concurrently on another
    // stream; ready[kb] counts its finished CTAs (ready_need per k-block).  Iteration i works on k-block (i + kb_rot) % n,
    // the order the pre-pass fills them in (the local panel first, then the peers' in ring order).
    extern __shared__ uint8_t smem_raw[];
    uint8_t *smem = reinterpret_cast<uint8_t *>((reinterpret_cast<uintptr_t>(smem_raw) + 1023) & ~(uintptr_t)1023);
    uint64_t *full_bar = reinterpret_cast<uint64_t *>(smem + kStages * kStageBytes);
    uint64_t *empty_bar = full_bar + kStages;
    uint64_t *tmem_full_bar = empty_bar + kStages;   // [2]: chunk accumulated in buffer b
    uint64_t *tmem_empty_bar = tmem_full_bar + 2;    // [2]: buffer b drained by all epilogue warps
    uint32_t *tmem_ptr_smem = reinterpret_cast<uint32_t *>(tmem_empty_bar + 2);

    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    // grouped rasterisation: consecutive CTAs walk kGroupM tiles down M before moving along N, so one wave of 148 CTAs
    // touches a near-square (16 x ~9 tiles) region: half the DRAM operand traffic of row-major order (ncu: 8.2 -> ~4 GB)
    const int tiles_m = (M + BM - 1) / BM, tiles_n = (N + BN - 1) / BN;
    int pid_m, pid_n;
    {
        const int pid = blockIdx.x, in_group = kGroupM * tiles_n;
        const int first_m = (pid / in_group) * kGroupM;
        const int group_m = (tiles_m - first_m < kGroupM) ? tiles_m - first_m : kGroupM;
        pid_m = first_m + (pid % in_group) % group_m;
        pid_n = (pid % in_group) / group_m;
    }
    const int m0 = pid_m * BM, n0 = pid_n * BN;
    const int num_chunks = (num_k_blocks + kc - 1) / kc;

    if (warp == 0 && lane == 0) {
        for (int s = 0; s < kStages; ++s) { mbar_init(&full_bar[s], 1); mbar_init(&empty_bar[s], 1); }
        for (int b = 0; b < 2; ++b) { mbar_init(&tmem_full_bar[b], 1); mbar_init(&tmem_empty_bar[b], kEpiWarps); }
        asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
        asm volatile("prefetch.tensormap [%0];" ::"l"(&map_a_hi) : "memory");
        asm volatile("prefetch.tensormap [%0];" ::"l"(&map_a_lo) : "memory");
        asm volatile("prefetch.tensormap [%0];" ::"l"(&map_b_hi) : "memory");
        asm volatile("prefetch.tensormap [%0];" ::"l"(&map_b_lo) : "memory");
    }
    if (warp == 1) {  // whole warp: allocate the accumulator columns
        asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(smem_u32(tmem_ptr_smem)), "r"(kTmemCols) : "memory");
        asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;" ::: "memory");
    }
    asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
    __syncthreads();
    asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
    const uint32_t tmem_base = *tmem_ptr_smem;

    if (warp == 0) {  // ===== TMA producer (lane 0 issues; the whole warp looks ahead on the readiness flags) =====
        int ready_upto = ready ? 0 : num_k_blocks;  // iterations [0, ready_upto) are known to be complete
        for (int it = 0; it < num_k_blocks; ++it) {
            if (it >= ready_upto) {
                // every lane polls the flag of iteration it + lane; the contiguous ready prefix advances the watermark
                const unsigned long long t0 = feo_peer::global_ns();
                for (;;) {
                    bool ok = false;
                    if (it + lane < num_k_blocks) {
                        int kbl = it + lane + kb_rot;
                        if (kbl >= num_k_blocks) kbl -= num_k_blocks;
                        unsigned int v;
                        asm volatile("ld.acquire.gpu.global.u32 %0, [%1];" : "=r"(v) : "l"(ready + kbl) : "memory");
                        ok = v >= ready_need;
                    }
                    const unsigned int mask = __ballot_sync(0xffffffffu, ok);
                    const int prefix = __ffs(~mask) - 1;  // number of leading ready iterations (32 if all)
                    if (prefix != 0) { ready_upto = it + (prefix < 0 ? 32 : prefix); break; }
                    if (peer_timeout_ns != 0 && feo_peer::global_ns() - t0 > peer_timeout_ns) {
                        // a peer never delivered its panel: raise the status flag and finish on whatever is there
                        if (lane == 0 && atomicOr(status, feo_peer::kStatusPeerTimeout) == 0)
                            printf("matmul_tf32x3: k-block %d of B never arrived (block %d); result invalid\n", it, blockIdx.x);
                        ready_upto = num_k_blocks;
                        break;
                    }
                    __nanosleep(100);
                }
                asm volatile("fence.proxy.async;" ::: "memory");  // the pre-pass wrote through the generic proxy; TMA reads through the async one
            }
            if (lane == 0) {
                int kb = it + kb_rot;
                if (kb >= num_k_blocks) kb -= num_k_blocks;
                const int s = it % kStages;
                const uint32_t ph = (it / kStages) & 1;
                mbar_wait(&empty_bar[s], ph ^ 1, 0, bar_limit);
                uint8_t *st = smem + s * kStageBytes;
                mbar_expect_tx(&full_bar[s], kStageBytes);
                tma_load_2d(st, &map_a_hi, &full_bar[s], kb * BK, m0);
                tma_load_2d(st + kTileABytes, &map_a_lo, &full_bar[s], kb * BK, m0);
                tma_load_2d(st + 2 * kTileABytes, &map_b_hi, &full_bar[s], kb * BK, n0);
                tma_load_2d(st + 2 * kTileABytes + kTileBBytes, &map_b_lo, &full_bar[s], kb * BK, n0);
            }
            __syncwarp();
        }
    } else if (warp == 1) {
        if (lane == 0) {  // ===== MMA issuer =====
            int kb = 0;
            for (int c = 0; c < num_chunks; ++c) {
                const int buf = c & 1;
                mbar_wait(&tmem_empty_bar[buf], ((c >> 1) & 1) ^ 1, 3, bar_limit);  // the epilogue has drained this buffer
                asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
                const uint32_t tmem_d = tmem_base + (uint32_t)(buf * BN);
                const int kb_end = (kb + kc < num_k_blocks) ? kb + kc : num_k_blocks;
                for (int first = 1; kb < kb_end; ++kb) {
                    const int s = kb % kStages;
                    const uint32_t ph = (kb / kStages) & 1;
                    mbar_wait(&full_bar[s], ph, 1, bar_limit);
                    asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
                    const uint32_t st = smem_u32(smem + s * kStageBytes);
                    const uint64_t da_hi = make_desc(st), da_lo = make_desc(st + kTileABytes);
                    const uint64_t db_hi = make_desc(st + 2 * kTileABytes), db_lo = make_desc(st + 2 * kTileABytes + kTileBBytes);
#pragma unroll
                    for (int k = 0; k < BK / UMMA_K; ++k) {
                        const uint64_t adv = (uint64_t)((k * UMMA_K * 4) >> 4);  // 32 bytes per step, in 16-byte units
                        umma_tf32(tmem_d, da_lo + adv, db_hi + adv, first ? 0u : 1u);  // the chunk's first MMA overwrites
                        first = 0;
                        umma_tf32(tmem_d, da_hi + adv, db_lo + adv, 1);
                        umma_tf32(tmem_d, da_hi + adv, db_hi + adv, 1);
                    }
                    umma_commit(&empty_bar[s]);  // frees the ring slot once these MMAs have read it
                }
                umma_commit(&tmem_full_bar[buf]);  // chunk complete in TMEM
            }
        }
    } else {  // ===== accumulate + epilogue: warps 2..9; a warp may only touch TMEM lanes 32*(warp%4) .. +31 =====
        const int q = warp & 3;                 // lane quarter
        const int half = (warp - 2) >> 2;       // which 128 of the 256 accumulator columns
        float acc[BN / 2];
#pragma unroll
        for (int i = 0; i < BN / 2; ++i) acc[i] = 0.f;
        for (int c = 0; c < num_chunks; ++c) {
            const int buf = c & 1;
            mbar_wait(&tmem_full_bar[buf], (c >> 1) & 1, 2, bar_limit);
            asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
#pragma unroll
            for (int j = 0; j < BN / 2; j += 32) {
                uint32_t v[32];
                const uint32_t taddr = tmem_base + ((uint32_t)(q * 32) << 16) + (uint32_t)(buf * BN + half * (BN / 2) + j);
                asm volatile(
                    "tcgen05.ld.sync.aligned.32x32b.x32.b32 "
                    "{%0,%1,%2,%3,%4,%5,%6,%7,%8,%9,%10,%11,%12,%13,%14,%15,%16,%17,%18,%19,%20,%21,%22,%23,%24,%25,%26,%27,%28,%29,%30,%31}, [%32];"
                    : "=r"(v[0]), "=r"(v[1]), "=r"(v[2]), "=r"(v[3]), "=r"(v[4]), "=r"(v[5]), "=r"(v[6]), "=r"(v[7]), "=r"(v[8]), "=r"(v[9]),
                      "=r"(v[10]), "=r"(v[11]), "=r"(v[12]), "=r"(v[13]), "=r"(v[14]), "=r"(v[15]), "=r"(v[16]), "=r"(v[17]), "=r"(v[18]),
                      "=r"(v[19]), "=r"(v[20]), "=r"(v[21]), "=r"(v[22]), "=r"(v[23]), "=r"(v[24]), "=r"(v[25]), "=r"(v[26]), "=r"(v[27]),
                      "=r"(v[28]), "=r"(v[29]), "=r"(v[30]), "=r"(v[31])
                    : "r"(taddr));
                asm volatile("tcgen05.wait::ld.sync.aligned;" ::: "memory");
#pragma unroll
                for (int i = 0; i < 32; ++i) acc[j + i] += __uint_as_float(v[i]);  // round-to-nearest f32 add
            }
            asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
            __syncwarp();
            if (lane == 0) mbar_arrive(&tmem_empty_bar[buf]);
        }
        // ragged edges: TMA zero-filled the out-of-range rows of A / Bt, so acc is well defined everywhere; only
        // the stores are guarded (128-bit when the row start is 16-byte aligned, i.e. N % 4 == 0)
        const int row = m0 + q * 32 + lane;
        const int col0 = n0 + half * (BN / 2);
        if (row < M) {
            float *crow = C + (size_t)row * N + col0;
            if ((N & 3) == 0) {
#pragma unroll
                for (int j = 0; j < BN / 2; j += 4)
                    if (col0 + j < N) *reinterpret_cast<float4 *>(crow + j) = make_float4(acc[j], acc[j + 1], acc[j + 2], acc[j + 3]);
            } else {
#pragma unroll
                for (int j = 0; j < BN / 2; ++j)
                    if (col0 + j < N) crow[j] = acc[j];
            }
        }
    }
    asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
    __syncthreads();
    if (warp == 1) {
        asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" ::"r"(tmem_base), "r"(kTmemCols) : "memory");
    }
}

// =====================================================================================================================
// The CTA-pair kernel (default): tcgen05.mma.cta_group::2 -- two CTAs of one cluster (the two SMs of a TPC) work on ONE
// 256 x BN output tile.  CTA r of the pair holds rows [128 r, 128 r + 128) of the A tile and columns [r BN/2, (r+1) BN/2) of the
// B tile in ITS shared memory and rows 128 r .. of the accumulator in ITS tensor memory; the MMA is issued by one thread of
// the even CTA and reads both shared memories.  Against one 128 x 256 tile per CTA this cuts the operand bytes every SM
// pulls from L2 and re-reads from shared memory by a third, which was what bounded the single-CTA kernel (tensor pipe 78 %
// active, L2 -> SM at ~11 TB/s), and the smaller stage leaves room for a 3-deep ring.  K is cut into chunks of kc k-blocks
// that alternate between two TMEM buffers and are added up in f32 registers with round-to-nearest (the tensor core truncates
// inside a chunk).
//   * persistent: the grid is min(tiles, SMs / 2) pairs; pair p works on tiles p, p + pairs, ... (grouped rasterisation), the
//     barrier phases run on across tiles, and the next tile's MMAs start while the epilogue warps still store the last one.
//   * BN is a run-time value (a multiple of 32, <= 256; it only enters the instruction descriptor, the TMA box and loop
//     bounds): the host picks the BN that minimises waves x BN, e.g. 8192 x 8192 with M = 1024 rows per GPU runs as
//     4 x 37 tiles of 256 x 224 = exactly two waves of 74 pairs instead of 128 tiles of 256 x 256 = 1.73 waves run as two.
//   * RAW_B (default when N % 4 == 0): B is never pre-processed.  TMA brings raw f32 tiles of B[K,N] -- boxes of 32 columns
//     x 32 k-rows with the 128-byte / 32-byte-atom swizzle, i.e. the MN-major ("transposed") tf32 operand layout of tcgen05 -- into shared
//     memory, two warps of each CTA split them in place (hi over the raw tile, lo into a second buffer; the swizzle is
//     irrelevant to an elementwise pass), fence them for the async proxy and arrive on the stage's barrier; the MMA reads B
//     MN-major (instruction-descriptor bit 16).  This removes the split + transpose pre-pass over B (768 MB of traffic and
//     ~115 us at 8192 x 8192 -- a fifth of the whole call for a 1024-row share of A at 8 GPUs) and halves B's L2 -> SM
//     stream.  A keeps its cheap row pre-pass (A_hi | A_lo, K-major): splitting it in the kernel too would put the shared-
//     memory pipe at ~95 % (the MMAs alone read 64 B/cycle of its 128).
//   * barriers: full[s] lives in the even CTA and collects the bytes of BOTH CTAs' A loads (cta_group::2 loads may signal
//     the peer's barrier) plus, with RAW_B, one arrival per splitting warp of the pair; raw[s] (RAW_B) is per CTA and
//     collects that CTA's raw-B bytes; empty[s] and tmem_full[b] exist in both CTAs and are signalled by multicast
//     tcgen05.commit; tmem_empty[b] lives in the even CTA and counts the 16 epilogue warps of the pair.
constexpr int P_BM = 128;                         // rows of A (and of the accumulator) per CTA; the pair tile is 2 * P_BM x BN
constexpr int P_BN_MAX = 256;
constexpr int P_STAGES = 3;
constexpr int P_TILE_A = P_BM * BK * 4;           // 16 KiB
constexpr int P_TILE_B = (P_BN_MAX / 2) * BK * 4; // 16 KiB at BN = 256 (fixed slot size; smaller BN use a prefix)
constexpr int P_STAGE = 2 * P_TILE_A + 2 * P_TILE_B;  // 64 KiB: A_hi | A_lo | B_hi (raw B lands here) | B_lo
constexpr int P_SMEM = P_STAGES * P_STAGE + 1024 + 256;
constexpr int P_GROUP_M = 8;                      // rasterisation group, in pair tiles along M
constexpr int P_THREADS = 384;                    // warp 0 TMA, warp 1 MMA, warps 2-3 split raw B, warps 4-11 accumulate + epilogue
constexpr int P_SPLIT_WARPS = 2;
constexpr int P_EPI_WARP0 = 4;
constexpr int P_BOX_N = 32;                       // raw-B TMA box: 32 columns (one 128-byte swizzle row) x BK k-rows
constexpr int P_BOX_BYTES = P_BOX_N * BK * 4;     // 4 KiB = four 8-row swizzle atoms, one per UMMA_K step

__device__ __forceinline__ uint32_t cluster_ctarank() {
    uint32_t r;
    asm volatile("mov.u32 %0, %%cluster_ctarank;" : "=r"(r));
    return r;
}
__device__ __forceinline__ void cluster_sync_all() {
    asm volatile("barrier.cluster.arrive.release.aligned;" ::: "memory");
    asm volatile("barrier.cluster.wait.acquire.aligned;" ::: "memory");
}
// shared::cluster address of `local` (a shared address of this CTA) in CTA `rank` of the cluster
__device__ __forceinline__ uint32_t map_to_cta(uint32_t local, uint32_t rank) {
    uint32_t r;
    asm volatile("mapa.shared::cluster.u32 %0, %1, %2;" : "=r"(r) : "r"(local), "r"(rank));
    return r;
}
// Arrive on a barrier in (possibly) the peer CTA.  Deliberately NOT .release.cluster: that form compiles to MEMBAR.ALL.GPU +
// ERRBAR in front of the arrive -- ~1700 cycles per hand-off, measured (profiles/r2_matmul_drain.md): it made the chunk drain
// the critical path (85 % of the tensor pipe's cycles at kc = 2, 97 % without it).  What is handed over here is tensor memory
// (loads completed by tcgen05.wait::ld, ordered by the tcgen05 fences on both sides) or shared memory already fenced for the
// async proxy, so the default form -- the one CUTLASS uses for the same hand-offs -- is enough.
__device__ __forceinline__ void mbar_arrive_cluster(uint32_t cluster_addr) {
    asm volatile("mbarrier.arrive.shared::cluster.b64 _, [%0];" ::"r"(cluster_addr) : "memory");
}
// TMA load into THIS CTA's shared memory whose bytes are counted by a barrier that may live in the peer CTA
__device__ __forceinline__ void tma_load_2d_pair(void *smem_dst, const CUtensorMap *map, uint32_t bar_cluster_addr, int c0, int c1) {
    asm volatile(
        "cp.async.bulk.tensor.2d.cta_group::2.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1, {%3, %4}], [%2];"
        ::"r"(smem_u32(smem_dst)), "l"(map), "r"(bar_cluster_addr), "r"(c0), "r"(c1)
        : "memory");
}
__device__ __forceinline__ void umma_tf32_pair(uint32_t tmem_d, uint64_t desc_a, uint64_t desc_b, uint32_t idesc, uint32_t accumulate) {
    asm volatile(
        "{\n\t.reg .pred p;\n\t"
        "setp.ne.b32 p, %4, 0;\n\t"
        "tcgen05.mma.cta_group::2.kind::tf32 [%0], %1, %2, %3, p;\n\t}"
        ::"r"(tmem_d), "l"(desc_a), "l"(desc_b), "r"(idesc), "r"(accumulate)
        : "memory");
}
// arrives (once the MMAs issued so far have completed) on the barrier at this offset in BOTH CTAs of the pair
__device__ __forceinline__ void umma_commit_pair(uint64_t *bar) {
    const uint16_t mask = 3;
    asm volatile("tcgen05.commit.cta_group::2.mbarrier::arrive::one.shared::cluster.multicast::cluster.b64 [%0], %1;" ::"r"(smem_u32(bar)), "h"(mask)
                 : "memory");
}
__device__ __forceinline__ void lds128(uint32_t addr, uint32_t (&v)[4]) {
    asm volatile("ld.shared.v4.b32 {%0,%1,%2,%3}, [%4];" : "=r"(v[0]), "=r"(v[1]), "=r"(v[2]), "=r"(v[3]) : "r"(addr) : "memory");
}
__device__ __forceinline__ void sts128(uint32_t addr, const uint32_t (&v)[4]) {
    asm volatile("st.shared.v4.b32 [%0], {%1,%2,%3,%4};" ::"r"(addr), "r"(v[0]), "r"(v[1]), "r"(v[2]), "r"(v[3]) : "memory");
}
__device__ __forceinline__ void tmem_ld16(uint32_t taddr, uint32_t (&v)[16]) {
    asm volatile(
        "tcgen05.ld.sync.aligned.32x32b.x16.b32 {%0,%1,%2,%3,%4,%5,%6,%7,%8,%9,%10,%11,%12,%13,%14,%15}, [%16];"
        : "=r"(v[0]), "=r"(v[1]), "=r"(v[2]), "=r"(v[3]), "=r"(v[4]), "=r"(v[5]), "=r"(v[6]), "=r"(v[7]), "=r"(v[8]), "=r"(v[9]), "=r"(v[10]),
          "=r"(v[11]), "=r"(v[12]), "=r"(v[13]), "=r"(v[14]), "=r"(v[15])
        : "r"(taddr));
}

// x = hi + lo exactly: hi = x rounded to TF32 (ptxas: add 0x1000, clear 13 bits), lo = the remainder (<= 13 significant bits)
__device__ __forceinline__ void split1(float x, float &hi, float &lo) {
    uint32_t h;
    asm("cvt.rna.tf32.f32 %0, %1;" : "=r"(h) : "f"(x));
    hi = __uint_as_float(h);
    lo = x - hi;
}

struct PairArgs {
    float *C;
    int M, N, BN;               // BN: pair tile width (multiple of 32, <= 256)
    int num_k_blocks, kc;
    int tiles_m, tiles_n;       // pair tiles: ceil(M / 256), ceil(N / BN)
    const unsigned int *ready;  // see matmul_tf32x3_kernel; the counter of k-block kb is ready[kb >> ready_shift]
    unsigned int ready_need;
    int ready_shift;
    int kb_rot;
    unsigned long long peer_timeout_ns;
    int *status;
    // PULL: B's row panels live in peer memory; one thread per CTA copies them into the local raw B (pull_dst) with TMA bulk
    // copies through two 16 KiB shared-memory slots, while the MMAs run on the panels already here
    feo_peer::BParts parts;
    feo_peer::XchgParams sync;
    float *pull_dst;            // [K, N]
    unsigned int *ready_w;      // the counters `ready` points at, writable
    int K;
    int own_kb;                 // > 0: the first own_kb k-block iterations are read straight from this rank's panel through map_b_lo
};

constexpr int P_PULL_CHUNK = 2048;                       // floats per bulk copy: 8 KiB of one row of B
constexpr int P_PULL_SLOTS = 4;                          // ring of staging slots: two loads and up to two stores in flight
constexpr int P_PULL_SMEM = P_PULL_SLOTS * P_PULL_CHUNK * 4;  // staged after the operand ring and the barriers

__device__ __forceinline__ void bulk_load(uint32_t smem_dst, const void *gsrc, uint32_t bytes, uint64_t *bar) {
    asm volatile("cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1], %2, [%3];" ::"r"(smem_dst), "l"(gsrc), "r"(bytes),
                 "r"(smem_u32(bar))
                 : "memory");
}
__device__ __forceinline__ void bulk_store(void *gdst, uint32_t smem_src, uint32_t bytes) {
    asm volatile("cp.async.bulk.global.shared::cta.bulk_group [%0], [%1], %2;" ::"l"(gdst), "r"(smem_src), "r"(bytes) : "memory");
    asm volatile("cp.async.bulk.commit_group;" ::: "memory");
}

template <bool RAW_B, bool PULL = false>
__global__ void __cluster_dims__(2, 1, 1) __launch_bounds__(P_THREADS, 1)
matmul_tf32x3_pair_kernel(const __grid_constant__ CUtensorMap map_a_hi, const __grid_constant__ CUtensorMap map_a_lo,
                          const __grid_constant__ CUtensorMap map_b_hi /* RAW_B: the raw B[K,N] */, const __grid_constant__ CUtensorMap map_b_lo,
                          const __grid_constant__ PairArgs args) {
    extern __shared__ uint8_t smem_raw[];
    uint8_t *smem = reinterpret_cast<uint8_t *>((reinterpret_cast<uintptr_t>(smem_raw) + 1023) & ~(uintptr_t)1023);
    uint64_t *full_bar = reinterpret_cast<uint64_t *>(smem + P_STAGES * P_STAGE);   // used in the even CTA
    uint64_t *empty_bar = full_bar + P_STAGES;                                       // both CTAs (multicast commit)
    uint64_t *raw_bar = empty_bar + P_STAGES;                                        // RAW_B: this CTA's raw-B bytes have landed
    uint64_t *tmem_full_bar = raw_bar + P_STAGES;                                    // [2], both CTAs (multicast commit)
    uint64_t *tmem_empty_bar = tmem_full_bar + 2;                                    // [2], used in the even CTA: 2 * kEpiWarps arrivals
    uint64_t *pull_bar = tmem_empty_bar + 2;                                         // [P_PULL_SLOTS], PULL: a peer chunk has landed
    uint32_t *tmem_ptr_smem = reinterpret_cast<uint32_t *>(pull_bar + P_PULL_SLOTS);
    uint8_t *pull_slots = smem + P_STAGES * P_STAGE + 256;                           // PULL: 2 x 16 KiB behind the barrier block

    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    const uint32_t cta = cluster_ctarank();          // 0 = even CTA (issues the MMAs), 1 = odd
    const int pair = blockIdx.x >> 1, num_pairs = gridDim.x >> 1;
    const int BN = args.BN, num_k_blocks = args.num_k_blocks, kc = args.kc;
    const int num_tiles = args.tiles_m * args.tiles_n;
    const int num_chunks = (num_k_blocks + kc - 1) / kc;
    const unsigned long long bar_limit = args.peer_timeout_ns ? 2 * args.peer_timeout_ns + 30000000000ull : 0;
    const int nbox = (BN / 2 + P_BOX_N - 1) / P_BOX_N;  // RAW_B: 32-column boxes per CTA and k-block (the last may overhang the tile)
    // bytes the even CTA's full[s] collects per stage: both CTAs' A tiles, and both CTAs' pre-split B tiles unless RAW_B
    const uint32_t full_tx_bytes = RAW_B ? 2u * (2u * P_TILE_A) : 2u * (2u * P_TILE_A + 2u * (uint32_t)(BN / 2) * BK * 4u);

    if (warp == 0 && lane == 0) {
        for (int s = 0; s < P_STAGES; ++s) {
            mbar_init(&full_bar[s], RAW_B ? 1 + 2 * P_SPLIT_WARPS : 1);
            mbar_init(&empty_bar[s], 1);
            mbar_init(&raw_bar[s], 1);
        }
        for (int b = 0; b < 2; ++b) { mbar_init(&tmem_full_bar[b], 1); mbar_init(&tmem_empty_bar[b], 2 * kEpiWarps); }
        for (int b = 0; b < P_PULL_SLOTS; ++b) mbar_init(&pull_bar[b], 1);
        asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
        asm volatile("prefetch.tensormap [%0];" ::"l"(&map_a_hi) : "memory");
        asm volatile("prefetch.tensormap [%0];" ::"l"(&map_a_lo) : "memory");
        asm volatile("prefetch.tensormap [%0];" ::"l"(&map_b_hi) : "memory");
        asm volatile("prefetch.tensormap [%0];" ::"l"(&map_b_lo) : "memory");
    }
    if (warp == 1) {  // the same warp in both CTAs: one allocation of all 512 columns in both tensor memories
        asm volatile("tcgen05.alloc.cta_group::2.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(smem_u32(tmem_ptr_smem)), "r"(kTmemCols) : "memory");
        asm volatile("tcgen05.relinquish_alloc_permit.cta_group::2.sync.aligned;" ::: "memory");
    }
    asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
    cluster_sync_all();  // barrier inits and the allocation are visible to both CTAs before anyone signals the peer
    asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
    const uint32_t tmem_base = *tmem_ptr_smem;

    // tile t -> (pair-tile row, column): P_GROUP_M tiles down M before moving along N (a wave touches a near-square region)
    auto tile_coords = [&](int t, int &tm, int &tn) {
        const int in_group = P_GROUP_M * args.tiles_n;
        const int first_m = (t / in_group) * P_GROUP_M;
        const int group_m = (args.tiles_m - first_m < P_GROUP_M) ? args.tiles_m - first_m : P_GROUP_M;
        tm = first_m + (t % in_group) % group_m;
        tn = (t % in_group) / group_m;
    };

    if (warp == 0) {  // ===== TMA producer (one in each CTA; lane 0 issues, the warp looks ahead on the readiness flags) =====
        uint32_t full_addr[P_STAGES];
#pragma unroll
        for (int s = 0; s < P_STAGES; ++s) full_addr[s] = map_to_cta(smem_u32(&full_bar[s]), 0);
        int g = 0;  // k-block iterations issued so far, across tiles: ring slot and phase
        for (int t = pair; t < num_tiles; t += num_pairs) {
            int tm, tn;
            tile_coords(t, tm, tn);
            const int m0 = tm * 2 * P_BM + (int)cta * P_BM, n0 = tn * BN + (int)cta * (BN / 2);
            const int own_kb = RAW_B ? args.own_kb : 0;
            int ready_upto = args.ready ? own_kb : num_k_blocks;
            for (int it = 0; it < num_k_blocks; ++it, ++g) {
                if (it >= ready_upto) {
                    const unsigned long long t0 = feo_peer::global_ns();
                    for (;;) {
                        bool ok = false;
                        if (it + lane < num_k_blocks) {
                            int kbl = it + lane + args.kb_rot;
                            if (kbl >= num_k_blocks) kbl -= num_k_blocks;
                            unsigned int v;
                            asm volatile("ld.acquire.gpu.global.u32 %0, [%1];" : "=r"(v) : "l"(args.ready + (kbl >> args.ready_shift)) : "memory");
                            ok = v >= args.ready_need;
                        }
                        const unsigned int mask = __ballot_sync(0xffffffffu, ok);
                        const int prefix = __ffs(~mask) - 1;
                        if (prefix != 0) { ready_upto = it + (prefix < 0 ? 32 : prefix); break; }
                        if (args.peer_timeout_ns != 0 && feo_peer::global_ns() - t0 > args.peer_timeout_ns) {
                            if (lane == 0 && atomicOr(args.status, feo_peer::kStatusPeerTimeout) == 0)
                                printf("matmul_tf32x3: k-block %d of B never arrived (block %d); result invalid\n", it, blockIdx.x);
                            ready_upto = num_k_blocks;
                            break;
                        }
                        __nanosleep(100);
                    }
                    asm volatile("fence.proxy.async;" ::: "memory");
                }
                if (lane == 0) {
                    int kb = it + args.kb_rot;
                    if (kb >= num_k_blocks) kb -= num_k_blocks;
                    const int s = g % P_STAGES;
                    const uint32_t ph = (g / P_STAGES) & 1;
                    mbar_wait(&empty_bar[s], ph ^ 1, 0, bar_limit);
                    uint8_t *st = smem + s * P_STAGE;
                    if (cta == 0) mbar_expect_tx(&full_bar[s], full_tx_bytes);
                    tma_load_2d_pair(st, &map_a_hi, full_addr[s], kb * BK, m0);
                    tma_load_2d_pair(st + P_TILE_A, &map_a_lo, full_addr[s], kb * BK, m0);
                    if (RAW_B) {
                        mbar_expect_tx(&raw_bar[s], (uint32_t)nbox * P_BOX_BYTES);
                        // this rank's own panel is read where it lies (map_b_lo, k relative to the panel); the rest from the
                        // local raw B the pullers fill
                        const bool own = it < own_kb;
                        const CUtensorMap *bmap = own ? &map_b_lo : &map_b_hi;
                        const int krow = own ? it * BK : kb * BK;
                        for (int i = 0; i < nbox; ++i)  // coordinates: (column, k-row) of B[K,N]
                            tma_load_2d(st + 2 * P_TILE_A + i * P_BOX_BYTES, bmap, &raw_bar[s], n0 + i * P_BOX_N, krow);
                    } else {
                        tma_load_2d_pair(st + 2 * P_TILE_A, &map_b_hi, full_addr[s], kb * BK, n0);
                        tma_load_2d_pair(st + 2 * P_TILE_A + P_TILE_B, &map_b_lo, full_addr[s], kb * BK, n0);
                    }
                }
                __syncwarp();
            }
        }
    } else if (warp == 1) {
        if (PULL && lane == (cta == 0 ? 1 : 0)) {
            // ===== puller: one otherwise idle thread per CTA (beside the MMA issuer in the even CTA: measured harmless) =====
            // All-gather of B fused into the matmul kernel: the chunks of B[K,N] -- 16 KiB row segments, numbered in the rotated
            // k order the MMAs walk (this rank's own panel first, then the peers' in ring order) -- are dealt round robin to the
            // CTAs of the grid; each is copied peer memory -> shared memory -> local raw B with TMA bulk copies (no registers,
            // a ring of four 8 KiB slots per CTA), and ready[k-block] counts the chunks that have landed.
            const feo_peer::XchgParams &sy = args.sync;
            if (sy.epoch != 0) {  // the peers' panels are complete: exchange number `epoch` (CTA 0 signals, every puller waits)
                if (blockIdx.x == 0) {
                    __threadfence_system();
                    for (int r = 0; r < sy.world; ++r)
                        feo_peer::st_release_sys(reinterpret_cast<unsigned long long *>(sy.win[r] + sy.rank * feo_peer::kFlagStride), sy.epoch);
                }
                const unsigned long long t0 = feo_peer::global_ns();
                for (int r = 0; r < sy.world; ++r) {
                    const unsigned long long *mine = reinterpret_cast<const unsigned long long *>(sy.win[sy.rank] + r * feo_peer::kFlagStride);
                    while (feo_peer::ld_acquire_sys(mine) < sy.epoch) {
                        if (sy.timeout_ns != 0 && feo_peer::global_ns() - t0 > sy.timeout_ns) {
                            if (sy.status && atomicOr(sy.status, feo_peer::kStatusPeerTimeout) == 0)
                                printf("matmul_tf32x3: rank %d gave up waiting for rank %d's panel of B; result invalid\n", sy.rank, r);
                            break;
                        }
                        __nanosleep(64);
                    }
                }
            }
            const int N = args.N, K = args.K;
            const int segs = (N + P_PULL_CHUNK - 1) / P_PULL_CHUNK;          // chunks per row of B
            // chunks in rotated order: (k-block iteration, row in block, segment); the first own_kb iterations are this rank's own
            // panel, which the TMA producer reads where it lies (args.own_kb > 0), so they are not copied at all
            const long long total = (long long)(num_k_blocks - args.own_kb) * BK * segs;
            const long long first = blockIdx.x, stride = gridDim.x;
            auto locate = [&](long long c, int &kb, const float *&src, float *&dst, uint32_t &bytes) {
                const int seg = (int)(c % segs);
                const long long rowi = c / segs;
                int it = (int)(rowi / BK) + args.own_kb;
                kb = it + args.kb_rot;
                if (kb >= num_k_blocks) kb -= num_k_blocks;
                const int k = kb * BK + (int)(rowi % BK);
                const int n0 = seg * P_PULL_CHUNK;
                const int len = (N - n0 < P_PULL_CHUNK) ? N - n0 : P_PULL_CHUNK;
                bytes = k < K ? (uint32_t)len * 4u : 0u;                      // rows past K (padding of the last k-block): nothing to copy
                const int r = k < K ? k / args.parts.rows_per_part : 0;
                src = args.parts.part[r] + (size_t)(k - r * args.parts.rows_per_part) * N + n0;
                dst = args.pull_dst + (size_t)k * N + n0;
            };
            constexpr int S = P_PULL_SLOTS, AHEAD = S - 2;
            uint32_t phase = 0;                       // bit s: parity the next wait on pull_bar[s] uses
            int ring_kb[S]; float *ring_dst[S]; uint32_t ring_bytes[S];
            const long long n_my = first < total ? (total - first + stride - 1) / stride : 0;
            auto issue = [&](long long j) {           // j-th of this CTA's chunks: peer memory -> slot j % S
                const int sl = (int)(j % S);
                const float *src;
                locate(first + j * stride, ring_kb[sl], src, ring_dst[sl], ring_bytes[sl]);
                if (ring_bytes[sl]) {
                    mbar_expect_tx(&pull_bar[sl], ring_bytes[sl]);
                    bulk_load(smem_u32(pull_slots + sl * (P_PULL_CHUNK * 4)), src, ring_bytes[sl], &pull_bar[sl]);
                }
            };
            for (long long j = 0; j < AHEAD && j < n_my; ++j) issue(j);
            int kb_prev = -1;                         // k-block of the chunk whose store may still be in flight
            for (long long i = 0; i < n_my; ++i) {
                const int sl = (int)(i % S);
                if (i + AHEAD < n_my) {
                    // slot (i + AHEAD) % S last held chunk i - 2: every store group but the newest has finished reading
                    asm volatile("cp.async.bulk.wait_group.read 1;" ::: "memory");
                    issue(i + AHEAD);
                }
                const uint32_t bytes = ring_bytes[sl];
                if (bytes) {
                    mbar_wait(&pull_bar[sl], (phase >> sl) & 1u, 5, bar_limit);
                    phase ^= 1u << sl;
                    bulk_store(ring_dst[sl], smem_u32(pull_slots + sl * (P_PULL_CHUNK * 4)), bytes);
                }
                if (kb_prev >= 0) {  // the store before this one is complete (all groups but the newest): publish its chunk
                    if (bytes) asm volatile("cp.async.bulk.wait_group 1;" ::: "memory");
                    else asm volatile("cp.async.bulk.wait_group 0;" ::: "memory");  // no new group was committed above
                    asm volatile("fence.proxy.async;" ::: "memory");
                    __threadfence();
                    atomicAdd(args.ready_w + kb_prev, 1u);
                }
                kb_prev = ring_kb[sl];
            }
            if (kb_prev >= 0) {
                asm volatile("cp.async.bulk.wait_group 0;" ::: "memory");
                asm volatile("fence.proxy.async;" ::: "memory");
                __threadfence();
                atomicAdd(args.ready_w + kb_prev, 1u);
            }
        } else if (cta == 0 && lane == 0) {  // ===== MMA issuer: one thread of the even CTA, for the pair =====
            // kind::tf32, f32 accumulate, A K-major, B K-major (pre-split) or MN-major (RAW_B), M = 256 (the pair), N = BN
            const uint32_t idesc = (1u << 4) | (2u << 7) | (2u << 10) | (RAW_B ? (1u << 16) : 0u) | ((uint32_t)(BN >> 3) << 17) |
                                   ((uint32_t)((2 * P_BM) >> 4) << 24);
            // descriptor step per UMMA_K: 32 bytes along a K-major row; one 1024-byte swizzle atom of an MN-major tile
            const uint64_t a_step = (uint64_t)((UMMA_K * 4) >> 4), b_step = RAW_B ? (uint64_t)(1024 >> 4) : a_step;
            int g = 0, cc = 0;  // k-block iterations / chunks so far, across tiles
            for (int t = pair; t < num_tiles; t += num_pairs) {
                int kb = 0;
                for (int c = 0; c < num_chunks; ++c, ++cc) {
                    const int buf = cc & 1;
                    mbar_wait(&tmem_empty_bar[buf], ((cc >> 1) & 1) ^ 1, 3, bar_limit);  // both CTAs' epilogue warps have drained it
                    asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
                    const uint32_t tmem_d = tmem_base + (uint32_t)(buf * P_BN_MAX);
                    const int kb_end = (kb + kc < num_k_blocks) ? kb + kc : num_k_blocks;
                    for (int first = 1; kb < kb_end; ++kb, ++g) {
                        const int s = g % P_STAGES;
                        const uint32_t ph = (g / P_STAGES) & 1;
                        mbar_wait(&full_bar[s], ph, 1, bar_limit);
                        asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
                        const uint32_t st = smem_u32(smem + s * P_STAGE);
                        const uint64_t da_hi = make_desc(st), da_lo = make_desc(st + P_TILE_A);
                        const uint64_t db_hi = RAW_B ? make_desc_mn(st + 2 * P_TILE_A, P_BOX_BYTES) : make_desc(st + 2 * P_TILE_A);
                        const uint64_t db_lo = RAW_B ? make_desc_mn(st + 2 * P_TILE_A + P_TILE_B, P_BOX_BYTES) : make_desc(st + 2 * P_TILE_A + P_TILE_B);
#pragma unroll
                        for (int k = 0; k < BK / UMMA_K; ++k) {
                            const uint64_t adv_a = (uint64_t)k * a_step, adv_b = (uint64_t)k * b_step;
                            umma_tf32_pair(tmem_d, da_lo + adv_a, db_hi + adv_b, idesc, first ? 0u : 1u);
                            first = 0;
                            umma_tf32_pair(tmem_d, da_hi + adv_a, db_lo + adv_b, idesc, 1);
                            umma_tf32_pair(tmem_d, da_hi + adv_a, db_hi + adv_b, idesc, 1);
                        }
                        umma_commit_pair(&empty_bar[s]);  // frees the ring slot in both CTAs
                    }
                    umma_commit_pair(&tmem_full_bar[buf]);  // chunk complete in both tensor memories
                }
            }
        }
    } else if (warp < P_EPI_WARP0) {
        if (RAW_B) {  // ===== split the raw B tile in place: hi = rna_tf32(b) over the raw values, lo = b - hi beside it =====
            const int tid = (int)threadIdx.x - 64;  // 0 .. 63
            uint32_t full_addr[P_STAGES];
#pragma unroll
            for (int s = 0; s < P_STAGES; ++s) full_addr[s] = map_to_cta(smem_u32(&full_bar[s]), 0);
            int g = 0;
            for (int t = pair; t < num_tiles; t += num_pairs) {
                for (int it = 0; it < num_k_blocks; ++it, ++g) {
                    const int s = g % P_STAGES;
                    const uint32_t ph = (g / P_STAGES) & 1;
                    mbar_wait(&raw_bar[s], ph, 4, bar_limit);
                    const uint32_t hi_addr = smem_u32(smem + s * P_STAGE + 2 * P_TILE_A) + (uint32_t)tid * 16u, lo_addr = hi_addr + P_TILE_B;
                    for (int b = 0; b < nbox; ++b) {
                        uint32_t v[4][4];
#pragma unroll
                        for (int j = 0; j < 4; ++j) lds128(hi_addr + (uint32_t)(b * P_BOX_BYTES + j * 1024), v[j]);
#pragma unroll
                        for (int j = 0; j < 4; ++j) {
                            uint32_t h[4], l[4];
#pragma unroll
                            for (int e = 0; e < 4; ++e) {
                                h[e] = (v[j][e] + 0x1000u) & 0xffffe000u;  // cvt.rna.tf32.f32 for finite values (what ptxas emits for it)
                                l[e] = __float_as_uint(__uint_as_float(v[j][e]) - __uint_as_float(h[e]));
                            }
                            sts128(hi_addr + (uint32_t)(b * P_BOX_BYTES + j * 1024), h);
                            sts128(lo_addr + (uint32_t)(b * P_BOX_BYTES + j * 1024), l);
                        }
                    }
                    asm volatile("fence.proxy.async.shared::cta;" ::: "memory");  // generic-proxy writes -> visible to the tensor core's reads
                    __syncwarp();
                    if (lane == 0) mbar_arrive_cluster(full_addr[s]);
                }
            }
        }
    } else {  // ===== accumulate + epilogue: 8 warps of each CTA on that CTA's 128 accumulator rows =====
        const int q = warp & 3;                           // TMEM lane quarter this warp may touch
        const int half = (warp - P_EPI_WARP0) >> 2;       // which BN / 2 of the BN accumulator columns
        const int hcols = BN / 2;               // a multiple of 16
        const uint32_t empty_addr0 = map_to_cta(smem_u32(&tmem_empty_bar[0]), 0), empty_addr1 = map_to_cta(smem_u32(&tmem_empty_bar[1]), 0);
        float acc[P_BN_MAX / 2];
        int cc = 0;
        for (int t = pair; t < num_tiles; t += num_pairs) {
            int tm, tn;
            tile_coords(t, tm, tn);
#pragma unroll
            for (int i = 0; i < P_BN_MAX / 2; ++i) acc[i] = 0.f;
            for (int c = 0; c < num_chunks; ++c, ++cc) {
                const int buf = cc & 1;
                mbar_wait(&tmem_full_bar[buf], (cc >> 1) & 1, 2, bar_limit);
                asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
                const uint32_t trow = tmem_base + ((uint32_t)(q * 32) << 16) + (uint32_t)(buf * P_BN_MAX + half * hcols);
#pragma unroll
                for (int j = 0; j < P_BN_MAX / 2; j += 16) {
                    if (j < hcols) {  // warp-uniform; the eight warps' loads share one 64 B/cycle TMEM read port, so one in flight each is enough
                        uint32_t v[16];
                        tmem_ld16(trow + (uint32_t)j, v);
                        asm volatile("tcgen05.wait::ld.sync.aligned;" ::: "memory");
#pragma unroll
                        for (int i = 0; i < 16; ++i) acc[j + i] += __uint_as_float(v[i]);  // round-to-nearest f32 add
                    }
                }
                asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
                __syncwarp();
                if (lane == 0) mbar_arrive_cluster(buf ? empty_addr1 : empty_addr0);
            }
            // TMA zero-filled the out-of-range rows of A / B, so acc is well defined everywhere; only the stores are guarded
            const int row = tm * 2 * P_BM + (int)cta * P_BM + q * 32 + lane;
            const int col0 = tn * BN + half * hcols;
            if (row < args.M) {
                float *crow = args.C + (size_t)row * args.N + col0;
                if ((args.N & 3) == 0) {
#pragma unroll
                    for (int j = 0; j < P_BN_MAX / 2; j += 4)
                        if (j < hcols && col0 + j < args.N) *reinterpret_cast<float4 *>(crow + j) = make_float4(acc[j], acc[j + 1], acc[j + 2], acc[j + 3]);
                } else {
#pragma unroll
                    for (int j = 0; j < P_BN_MAX / 2; ++j)
                        if (j < hcols && col0 + j < args.N) crow[j] = acc[j];
                }
            }
        }
    }
    // neither CTA may leave (or free its tensor memory) while the peer can still signal its barriers or read its tiles
    asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
    cluster_sync_all();
    if (warp == 1) {
        asm volatile("tcgen05.dealloc.cta_group::2.sync.aligned.b32 %0, %1;" ::"r"(tmem_base), "r"(kTmemCols) : "memory");
    }
}

// ---- pre-pass: hi/lo split --------------------------------------------------------------------------------
// A[M,K] -> hi/lo [M,Kp] (Kp = K rounded up to BK, zero padded).
__global__ void __launch_bounds__(256) split_rows_kernel(float *__restrict__ hi, float *__restrict__ lo, const float *__restrict__ x, int M, int K, int Kp, bool vec) {
    const size_t nvec = (size_t)M * (Kp / 4);
    for (size_t i = (size_t)blockIdx.x * 256 + threadIdx.x; i < nvec; i += (size_t)gridDim.x * 256) {
        const size_t m = i / (Kp / 4);
        const int k = (int)(i % (Kp / 4)) * 4;
        float v[4] = {0.f, 0.f, 0.f, 0.f};
        if (vec && k + 3 < K) {
            const float4 t = *reinterpret_cast<const float4 *>(x + m * K + k);
            v[0] = t.x; v[1] = t.y; v[2] = t.z; v[3] = t.w;
        } else {
#pragma unroll
            for (int j = 0; j < 4; ++j)
                if (k + j < K) v[j] = x[m * K + k + j];
        }
        float4 h, l;
        split1(v[0], h.x, l.x); split1(v[1], h.y, l.y); split1(v[2], h.z, l.z); split1(v[3], h.w, l.w);
        reinterpret_cast<float4 *>(hi)[i] = h;
        reinterpret_cast<float4 *>(lo)[i] = l;
    }
}
// B[K,N] -> hi/lo [N,Kp] through a padded 32x32 shared tile (coalesced on both sides); zero beyond K.
__global__ void __launch_bounds__(256) split_transpose_kernel(float *__restrict__ hi, float *__restrict__ lo, const float *__restrict__ b, int K, int N, int Kp) {
    __shared__ float tile[32][33];
    const int k0 = blockIdx.y * 32, n0 = blockIdx.x * 32;
    const int tx = threadIdx.x & 31, ty = threadIdx.x >> 5;  // 32 x 8
#pragma unroll
    for (int r = ty; r < 32; r += 8) tile[r][tx] = (k0 + r < K && n0 + tx < N) ? b[(size_t)(k0 + r) * N + n0 + tx] : 0.f;
    __syncthreads();
#pragma unroll
    for (int r = ty; r < 32; r += 8) {
        if (n0 + r < N) {  // k0 + tx < Kp always (Kp is a multiple of 32)
            float h, l;
            split1(tile[tx][r], h, l);
            hi[(size_t)(n0 + r) * Kp + k0 + tx] = h;
            lo[(size_t)(n0 + r) * Kp + k0 + tx] = l;
        }
    }
}

// The same for a B that is row-sharded across GPUs (all-gather of B -> matmul, fused): the panels stay where
// they are and this kernel pulls them over NVLink while it splits and transposes -- no gathered copy of B is ever written.
// Tile = 32 k-rows x 128 n-columns; every lane loads 16 bytes along n (512 contiguous bytes per warp request, all four of
// a thread's requests in flight before the first use) and stores 16 bytes along k.
constexpr int GT_K = 32, GT_N = 128;
// k-blocks are produced in rotated order (blockIdx.y = 0 is the first k-block of THIS rank's own panel, then the peers' in
// ring order): every rank pulls from a different owner at any time, instead of all eight draining one GPU's links.
// `ready` (optional): ready[k-block] += 1 when a CTA has published its tile -- the matmul kernel, already running on the
// main stream, waits on these per k-block, so the NVLink pull overlaps the tensor-core work on the panels already here.
// (<= 42 registers so that these CTAs fit on an SM beside one 168-register matmul CTA.)
__global__ void __launch_bounds__(256, 6) split_transpose_gather_kernel(float *__restrict__ hi, float *__restrict__ lo,
                                                                        const __grid_constant__ feo_peer::BParts parts,
                                                                        const __grid_constant__ feo_peer::XchgParams sync, int K, int N, int Kp,
                                                                        bool vec, int kb_rot, unsigned int *__restrict__ ready) {
    __shared__ float tile[GT_K][GT_N + 1];
    feo_peer::signal_and_wait(sync);
    int kblock = blockIdx.y + kb_rot;
    if (kblock >= (int)gridDim.y) kblock -= gridDim.y;
    const int k0 = kblock * GT_K, n0 = blockIdx.x * GT_N;
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
    float4 v[4];
#pragma unroll
    for (int j = 0; j < 4; ++j) {
        const int k = k0 + warp * 4 + j, n = n0 + lane * 4;
        v[j] = make_float4(0.f, 0.f, 0.f, 0.f);
        if (k < K) {
            const int r = k / parts.rows_per_part;
            const float *row = parts.part[r] + (size_t)(k - r * parts.rows_per_part) * N;
            if (vec && n + 3 < N) {
                v[j] = feo_peer::ld_peer(reinterpret_cast<const float4 *>(row + n));
            } else {
                float t[4] = {0.f, 0.f, 0.f, 0.f};
#pragma unroll
                for (int c = 0; c < 4; ++c)
                    if (n + c < N) t[c] = feo_peer::ld_peer(row + n + c);
                v[j] = make_float4(t[0], t[1], t[2], t[3]);
            }
        }
    }
#pragma unroll
    for (int j = 0; j < 4; ++j) {
        float *dst = &tile[warp * 4 + j][lane * 4];
        dst[0] = v[j].x; dst[1] = v[j].y; dst[2] = v[j].z; dst[3] = v[j].w;
    }
    __syncthreads();
    // thread -> (n = pass * 32 + tid / 8, four k at (tid % 8) * 4): 128-byte rows of Bt, conflict-free column reads
    const int kq = (threadIdx.x & 7) * 4;
#pragma unroll
    for (int pass = 0; pass < 4; ++pass) {
        const int nl = pass * 32 + (threadIdx.x >> 3);
        if (n0 + nl < N) {
            float4 h, l;
            split1(tile[kq + 0][nl], h.x, l.x); split1(tile[kq + 1][nl], h.y, l.y);
            split1(tile[kq + 2][nl], h.z, l.z); split1(tile[kq + 3][nl], h.w, l.w);
            const size_t o = (size_t)(n0 + nl) * Kp + k0 + kq;
            *reinterpret_cast<float4 *>(hi + o) = h;
            *reinterpret_cast<float4 *>(lo + o) = l;
        }
    }
    if (ready) {
        asm volatile("fence.proxy.async;" ::: "memory");
        __threadfence();
        __syncthreads();
        if (threadIdx.x == 0) atomicAdd(ready + kblock, 1u);
    }
}

// RAW_B with a row-sharded B: the panels are pulled over NVLink into ONE local raw copy of B[K,N] (no split, no transpose --
// the matmul kernel does that on the tiles), k-blocks in the same rotated order, one ready[] count per k-block and CTA.
// A CTA moves 32 k-rows x 1024 columns; every thread keeps 8 sixteen-byte peer loads in flight.
constexpr int GR_COLS = 1024;
__global__ void __launch_bounds__(256) gather_rows_kernel(float *__restrict__ braw, const __grid_constant__ feo_peer::BParts parts,
                                                          const __grid_constant__ feo_peer::XchgParams sync, int K, int N, int kb_rot,
                                                          unsigned int *__restrict__ ready) {
    feo_peer::signal_and_wait(sync);
    int kblock = blockIdx.y + kb_rot;
    if (kblock >= (int)gridDim.y) kblock -= gridDim.y;
    const int k0 = kblock * GT_K, n = blockIdx.x * GR_COLS + threadIdx.x * 4;
    if (n < N) {  // N % 4 == 0: whole vectors
#pragma unroll
        for (int r0 = 0; r0 < GT_K; r0 += 8) {
            float4 v[8];
#pragma unroll
            for (int j = 0; j < 8; ++j) {
                const int k = k0 + r0 + j;
                if (k < K) {
                    const int r = k / parts.rows_per_part;
                    v[j] = feo_peer::ld_peer(reinterpret_cast<const float4 *>(parts.part[r] + (size_t)(k - r * parts.rows_per_part) * N + n));
                }
            }
#pragma unroll
            for (int j = 0; j < 8; ++j) {
                const int k = k0 + r0 + j;
                if (k < K) *reinterpret_cast<float4 *>(braw + (size_t)k * N + n) = v[j];
            }
        }
    }
    if (ready) {
        asm volatile("fence.proxy.async;" ::: "memory");
        __threadfence();
        __syncthreads();
        if (threadIdx.x == 0) atomicAdd(ready + kblock, 1u);
    }
}

// Copy-engine route: the opening flag exchange (the peers' panels are complete) as a one-CTA kernel in front of the copies.
__global__ void __launch_bounds__(32) pull_open_kernel(const __grid_constant__ feo_peer::XchgParams sync) { feo_peer::signal_and_wait(sync); }

typedef CUresult (*EncodeTiledFn)(CUtensorMap *, CUtensorMapDataType, cuuint32_t, void *, const cuuint64_t *, const cuuint64_t *,
                                  const cuuint32_t *, const cuuint32_t *, CUtensorMapInterleave, CUtensorMapSwizzle,
                                  CUtensorMapL2promotion, CUtensorMapFloatOOBfill);

EncodeTiledFn get_encode() {
    static EncodeTiledFn fn = nullptr;
    if (!fn) {
        void *p = nullptr;
        cudaDriverEntryPointQueryResult q;
        if (cudaGetDriverEntryPoint("cuTensorMapEncodeTiled", &p, cudaEnableDefault, &q) == cudaSuccess && q == cudaDriverEntryPointSuccess)
            fn = reinterpret_cast<EncodeTiledFn>(p);
    }
    return fn;
}

typedef CUresult (*StreamWriteFn)(CUstream, CUdeviceptr, cuuint32_t, unsigned int);
StreamWriteFn get_stream_write() {
    void *p = nullptr;
    cudaDriverEntryPointQueryResult q;
    if (cudaGetDriverEntryPoint("cuStreamWriteValue32", &p, cudaEnableDefault, &q) == cudaSuccess && q == cudaDriverEntryPointSuccess)
        return reinterpret_cast<StreamWriteFn>(p);
    return nullptr;
}

// 2-D row-major [rows, K] f32 tensor, box = BK x box_rows, 128-byte swizzle (16-byte atoms for K-major operands, 32-byte
// atoms for the raw MN-major B).
bool make_map(CUtensorMap *map, const float *base, size_t rows, size_t K, int box_rows, bool atom32 = false) {
    EncodeTiledFn enc = get_encode();
    if (!enc) return false;
    const cuuint64_t dims[2] = {(cuuint64_t)K, (cuuint64_t)rows};
    const cuuint64_t strides[1] = {(cuuint64_t)K * 4};
    const cuuint32_t box[2] = {(cuuint32_t)BK, (cuuint32_t)box_rows};
    const cuuint32_t estr[2] = {1, 1};
    return enc(map, CU_TENSOR_MAP_DATA_TYPE_FLOAT32, 2, const_cast<float *>(base), dims, strides, box, estr, CU_TENSOR_MAP_INTERLEAVE_NONE,
               atom32 ? CU_TENSOR_MAP_SWIZZLE_128B_ATOM_32B : CU_TENSOR_MAP_SWIZZLE_128B, CU_TENSOR_MAP_L2_PROMOTION_L2_256B,
               CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE) == CUDA_SUCCESS;
}

}  // namespace

bool feo_matmul_tf32x3_supported(size_t M, size_t K, size_t N) {
    // any shape works (ragged edges are zero-filled / guarded); tiny products stay on the SIMT kernel, which is
    // launch-bound there anyway and keeps the k-order of the reference
    const size_t lim = size_t(1) << 31;
    return M >= 1 && N >= 1 && K >= 1 && M < lim && N < lim && K < lim - BK && M * N * K >= (size_t(1) << 21);
}

namespace {

int matmul_tf32x3_run(feo_ctx *ctx, void *C, const void *A, const void *B, const feo_peer::BParts *parts, const feo_peer::XchgParams *sync,
                      size_t M, size_t K, size_t N) {
    if (!feo_aligned16(A) || !feo_aligned16(C) || (B && !feo_aligned16(B))) return feo_fail(ctx, FEO_ERR_UNSUPPORTED, "TF32X3 needs 16-byte aligned operands");
    const bool pair = ctx->knobs[FEO_KNOB_MATMUL_TF32_PAIR] != 1;
    // RAW_B: the CTA-pair kernel reads B as it lies in memory and splits it on the fly (TMA needs a 16-byte row pitch)
    bool raw_b = pair && ctx->knobs[FEO_KNOB_MATMUL_TF32_BRAW] != 1 && N % 4 == 0;
    if (parts)
        for (int r = 0; r < parts->world; ++r) raw_b = raw_b && feo_aligned16(parts->part[r]);
    static bool attr_set[64] = {false};  // the opt-in is per device
    const int dev = ctx->device & 63;
    if (!attr_set[dev]) {
        FEO_CUDA(ctx, cudaFuncSetAttribute(matmul_tf32x3_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, kSmemBytes));
        FEO_CUDA(ctx, cudaFuncSetAttribute(matmul_tf32x3_pair_kernel<false>, cudaFuncAttributeMaxDynamicSharedMemorySize, P_SMEM));
        FEO_CUDA(ctx, cudaFuncSetAttribute(matmul_tf32x3_pair_kernel<true>, cudaFuncAttributeMaxDynamicSharedMemorySize, P_SMEM));
        FEO_CUDA(ctx, cudaFuncSetAttribute(matmul_tf32x3_pair_kernel<true, true>, cudaFuncAttributeMaxDynamicSharedMemorySize, P_SMEM + P_PULL_SMEM));
        attr_set[dev] = true;
    }
    // workspace: A_hi | A_lo [M,Kp], then either Bt_hi | Bt_lo [N,Kp] (pre-pass) or, for RAW_B with a sharded B, the gathered
    // raw B [K,N]; Kp = K rounded up to one k-block; then one readiness counter per k-block
    const size_t Kp = (K + BK - 1) / BK * BK;
    void *ws = nullptr;
    const size_t a_elems = M * Kp, b_elems = raw_b ? (parts ? K * N : 0) : 2 * N * Kp;
    int rc = feo_workspace(ctx, (2 * a_elems + b_elems) * sizeof(float) + (Kp / BK) * sizeof(unsigned int), &ws);
    if (rc != FEO_OK) return rc;
    float *a_hi = (float *)ws, *a_lo = a_hi + a_elems, *b_ws = a_lo + a_elems;
    float *b_hi = b_ws, *b_lo = raw_b ? nullptr : b_ws + N * Kp;
    unsigned int *ready_ws = reinterpret_cast<unsigned int *>(b_ws + b_elems);
    const int num_kb = (int)(Kp / BK);
    static_assert(GT_K == BK, "one gather tile row = one k-block of the main kernel");

    // B arriving from peers, four routes (knob FEO_KNOB_GATHER_OVERLAP; measured in DESIGN.md 7.1):
    //   copy engines (0 = auto when every panel is whole k-blocks, 3 = forced): peer cudaMemcpyAsync pieces into the local raw B
    //       on two side streams, one readiness flag per piece; no SM is involved, so the pull proceeds at NVLink speed while
    //       the persistent matmul kernel owns every SM.  This rank's own panel is read in place.
    //   in-kernel pull (0 = auto otherwise, 4 = forced): one thread per CTA of the matmul kernel copies 8 KiB chunks peer memory
    //       -> shared memory -> local raw B with TMA bulk copies (PULL) -- one launch does the all-gather and the product.
    //   gather kernel on a high-priority side stream (2) / in plain stream order (1): round 1's scheme, and the only one for a
    //       pre-split B (N % 4 != 0) or the single-CTA kernel.
    unsigned int *ready = nullptr;
    unsigned int ready_need = 0;
    int kb_rot = 0;
    bool overlap = false, pull = false, own_direct = false;  // own_direct: this rank's panel is read in place (mb_lo), not copied
    int n_streams = 1;    // copy-engine route: side streams in use (their join events are waited for behind the matmul kernel)
    int ready_shift = 0;  // one readiness counter per 2^ready_shift k-blocks (copy-engine route: one per piece)
    if (parts) {
        kb_rot = (int)(((size_t)sync->rank * parts->rows_per_part) / BK);
        const int kn = ctx->knobs[FEO_KNOB_GATHER_OVERLAP];
        pull = raw_b && (kn == 4 || (kn == 0 && parts->rows_per_part % BK != 0));
        // copy engines: needs whole k-blocks per panel (the own panel is then read in place through mb_lo and never copied)
        const bool ce = raw_b && (kn == 3 || kn == 0) && parts->rows_per_part % BK == 0 && feo_side_stream(ctx) == FEO_OK;
        if (ce) {
            // Pieces of up to 16 k-blocks (16 MiB at N = 8192; a power of two that divides the panel), dealt alternately to two side
            // streams, one readiness flag per piece (cuStreamWriteValue32 where the driver takes it for this memory, else a 4-byte
            // host-to-device copy: both are front-end / copy-engine work, no SM).  A stream operation costs several us whatever
            // its size and the consumer needs the pieces in order, so the piece size is a compromise -- measured at 8 GPUs
            // (tools/peer_bench.py, whole call): 28 pieces on one stream 1051 us, 14 pieces on two streams 747 us, 7 pieces
            // (whole panels) on two streams ~900 us (two panels share the ingress, the first one needed arrives late).
            int piece_max = 16;
            n_streams = 2;
            if (const int shape = ctx->knobs[FEO_KNOB_PULL_SHAPE]; shape > 0) {  // experiments: 100 * streams + k-blocks per piece
                n_streams = shape / 100 < 1 ? 1 : (shape / 100 > 4 ? 4 : shape / 100);
                piece_max = shape % 100 > 0 ? shape % 100 : 16;
            }
            int piece_kb = 1;
            while (2 * piece_kb <= piece_max && (parts->rows_per_part / BK) % (2 * piece_kb) == 0) piece_kb *= 2;
            while ((1 << ready_shift) < piece_kb) ++ready_shift;
            static StreamWriteFn write32 = get_stream_write();  // cleared for good the first time the driver refuses (e.g. for pool memory)
            if (ctx->host_ones_len < (size_t)num_kb) {
                if (ctx->host_ones) cudaFreeHost(ctx->host_ones);
                ctx->host_ones = nullptr;
                ctx->host_ones_len = 0;
                FEO_CUDA(ctx, cudaMallocHost(&ctx->host_ones, (size_t)num_kb * sizeof(unsigned int)));
                for (int i = 0; i < num_kb; ++i) ctx->host_ones[i] = 1u;
                ctx->host_ones_len = (size_t)num_kb;
            }
            ready = ready_ws;
            ready_need = 1;
            own_direct = true;
            cudaStream_t ss[4] = {ctx->side_stream, ctx->pull_streams[0], ctx->pull_streams[1], ctx->pull_streams[2]};
            FEO_CUDA(ctx, cudaMemsetAsync(ready, 0, (size_t)num_kb * sizeof(unsigned int), ctx->stream));
            FEO_CUDA(ctx, cudaEventRecord(ctx->ev_fork, ctx->stream));
            FEO_CUDA(ctx, cudaStreamWaitEvent(ss[0], ctx->ev_fork, 0));
            pull_open_kernel<<<1, 32, 0, ss[0]>>>(*sync);
            FEO_LAUNCH_CHECK(ctx);
            FEO_CUDA(ctx, cudaEventRecord(ctx->ev_open, ss[0]));
            for (int i = 1; i < n_streams; ++i) FEO_CUDA(ctx, cudaStreamWaitEvent(ss[i], ctx->ev_open, 0));
            const int kb_per_part = parts->rows_per_part / BK;
            int piece = 0;
            for (int step = 1; step < parts->world; ++step) {  // ring order: every rank drains a different owner at any time
                const int r = (sync->rank + step) % parts->world;
                for (int kb0 = 0; kb0 < kb_per_part; kb0 += piece_kb, ++piece) {
                    const size_t row0 = (size_t)r * parts->rows_per_part + (size_t)kb0 * BK;
                    cudaStream_t st = ss[piece % n_streams];
                    FEO_CUDA(ctx, cudaMemcpyAsync(b_ws + row0 * N, parts->part[r] + (size_t)kb0 * BK * N, (size_t)piece_kb * BK * N * sizeof(float),
                                                  cudaMemcpyDeviceToDevice, st));
                    unsigned int *flag = ready + ((row0 / BK) >> ready_shift);
                    if (write32 && write32((CUstream)st, (CUdeviceptr)(uintptr_t)flag, 1u, 0) != CUDA_SUCCESS) write32 = nullptr;
                    if (!write32) FEO_CUDA(ctx, cudaMemcpyAsync(flag, ctx->host_ones, sizeof(unsigned int), cudaMemcpyHostToDevice, st));
                }
            }
            FEO_CUDA(ctx, cudaEventRecord(ctx->ev_join, ss[0]));
            for (int i = 1; i < n_streams; ++i) FEO_CUDA(ctx, cudaEventRecord(ctx->pull_joins[i - 1], ss[i]));
            overlap = true;   // join the side streams behind the matmul kernel
        } else if (pull) {
            ready = ready_ws;
            ready_need = (unsigned)(BK * feo_ceil_div(N, (size_t)P_PULL_CHUNK));
            own_direct = parts->rows_per_part % BK == 0;
            FEO_CUDA(ctx, cudaMemsetAsync(ready, 0, (size_t)num_kb * sizeof(unsigned int), ctx->stream));
        } else {
            overlap = kn != 1 && feo_side_stream(ctx) == FEO_OK;
            cudaStream_t gstream = ctx->stream;
            if (overlap) {
                ready = ready_ws;
                FEO_CUDA(ctx, cudaMemsetAsync(ready, 0, (size_t)num_kb * sizeof(unsigned int), ctx->stream));
                FEO_CUDA(ctx, cudaEventRecord(ctx->ev_fork, ctx->stream));
                FEO_CUDA(ctx, cudaStreamWaitEvent(ctx->side_stream, ctx->ev_fork, 0));
                gstream = ctx->side_stream;
            }
            if (raw_b) {
                const dim3 ggrid((unsigned)feo_ceil_div(N, (size_t)GR_COLS), (unsigned)num_kb);
                ready_need = ggrid.x;
                gather_rows_kernel<<<ggrid, 256, 0, gstream>>>(b_ws, *parts, *sync, (int)K, (int)N, kb_rot, ready);
            } else {
                bool vec = N % 4 == 0;
                for (int r = 0; r < parts->world; ++r) vec = vec && feo_aligned16(parts->part[r]);
                const dim3 ggrid((unsigned)feo_ceil_div(N, (size_t)GT_N), (unsigned)num_kb);
                ready_need = ggrid.x;
                split_transpose_gather_kernel<<<ggrid, 256, 0, gstream>>>(b_hi, b_lo, *parts, *sync, (int)K, (int)N, (int)Kp, vec, kb_rot, ready);
            }
            FEO_LAUNCH_CHECK(ctx);
            if (overlap) FEO_CUDA(ctx, cudaEventRecord(ctx->ev_join, ctx->side_stream));
        }
    }

    size_t blocks = feo_ceil_div(a_elems / 4, (size_t)256);
    if (blocks > (size_t)ctx->sm_count * 16) blocks = (size_t)ctx->sm_count * 16;
    split_rows_kernel<<<(unsigned)blocks, 256, 0, ctx->stream>>>(a_hi, a_lo, (const float *)A, (int)M, (int)K, (int)Kp, K % 4 == 0);
    FEO_LAUNCH_CHECK(ctx);
    if (!parts && !raw_b) {
        split_transpose_kernel<<<dim3((unsigned)feo_ceil_div(N, (size_t)32), (unsigned)(Kp / 32)), 256, 0, ctx->stream>>>(b_hi, b_lo, (const float *)B, (int)K, (int)N, (int)Kp);
        FEO_LAUNCH_CHECK(ctx);
    }

    CUtensorMap ma_hi, ma_lo, mb_hi, mb_lo;
    int kc = ctx->knobs[FEO_KNOB_MATMUL_TF32_KC];
    if (kc <= 0) kc = 2;  // k-blocks per TMEM accumulation chain (24 truncating accumulates), see the file header
    if (pair) {
        // CTA-pair kernel: persistent over min(tiles, SMs / 2) pairs; BN = the tile width with the fewest (waves x width)
        const int pairs_max = ctx->sm_count / 2;
        const size_t tiles_m = feo_ceil_div(M, (size_t)(2 * P_BM));
        int bn = ctx->knobs[FEO_KNOB_MATMUL_TF32_BN];
        if (bn <= 0 || bn > P_BN_MAX || bn % 32 != 0) {
            double best = 0;
            bn = P_BN_MAX;
            for (int cand = P_BN_MAX; cand >= 128; cand -= 32) {
                const size_t t = tiles_m * feo_ceil_div(N, (size_t)cand);
                // cost of a tile ~ its width plus a fixed part (narrow tiles stream A more often per flop and drain as often)
                const double cost = (double)feo_ceil_div(t, (size_t)pairs_max) * (cand + 16);
                if (best == 0 || cost < best * 0.985) { best = cost; bn = cand; }
            }
        }
        const size_t tiles_n = feo_ceil_div(N, (size_t)bn), tiles = tiles_m * tiles_n;
        if (tiles > 0x3fffffffu) return feo_fail(ctx, FEO_ERR_UNSUPPORTED, "matmul grid too large");
        bool ok = make_map(&ma_hi, a_hi, M, Kp, P_BM) && make_map(&ma_lo, a_lo, M, Kp, P_BM);
        if (raw_b) {
            // B[K,N] (or its gathered copy) as it lies: inner dimension N, boxes of 32 columns x 32 k-rows; rows >= K read as zero
            ok = ok && make_map(&mb_hi, parts ? b_ws : (const float *)B, K, N, BK, true);
            mb_lo = mb_hi;
            // PULL with whole k-blocks per panel: this rank's own panel is never copied, the kernel reads it through mb_lo
            if (own_direct)
                ok = ok && make_map(&mb_lo, parts->part[sync->rank], (size_t)parts->rows_per_part, N, BK, true);
        } else {
            ok = ok && make_map(&mb_hi, b_hi, N, Kp, bn / 2) && make_map(&mb_lo, b_lo, N, Kp, bn / 2);
        }
        if (!ok) return feo_fail(ctx, FEO_ERR_CUDA, "cuTensorMapEncodeTiled failed");
        PairArgs pa{};
        pa.C = (float *)C;
        pa.M = (int)M;
        pa.N = (int)N;
        pa.BN = bn;
        pa.num_k_blocks = num_kb;
        pa.kc = kc;
        pa.tiles_m = (int)tiles_m;
        pa.tiles_n = (int)tiles_n;
        pa.ready = ready;
        pa.ready_need = ready_need;
        pa.ready_shift = ready_shift;
        pa.kb_rot = kb_rot;
        pa.peer_timeout_ns = ctx->peer_timeout_ns;
        pa.status = ctx->status_flag;
        pa.own_kb = own_direct ? parts->rows_per_part / BK : 0;
        if (pull) {
            pa.parts = *parts;
            pa.sync = *sync;
            pa.pull_dst = b_ws;
            pa.ready_w = ready;
            pa.K = (int)K;
        }
        const unsigned pairs = (unsigned)(tiles < (size_t)pairs_max ? tiles : (size_t)pairs_max);
        if (pull)
            matmul_tf32x3_pair_kernel<true, true><<<dim3(2 * pairs), P_THREADS, P_SMEM + P_PULL_SMEM, ctx->stream>>>(ma_hi, ma_lo, mb_hi, mb_lo, pa);
        else if (raw_b)
            matmul_tf32x3_pair_kernel<true><<<dim3(2 * pairs), P_THREADS, P_SMEM, ctx->stream>>>(ma_hi, ma_lo, mb_hi, mb_lo, pa);
        else
            matmul_tf32x3_pair_kernel<false><<<dim3(2 * pairs), P_THREADS, P_SMEM, ctx->stream>>>(ma_hi, ma_lo, mb_hi, mb_lo, pa);
        FEO_LAUNCH_CHECK(ctx);
    } else {
        if (!make_map(&ma_hi, a_hi, M, Kp, BM) || !make_map(&ma_lo, a_lo, M, Kp, BM) || !make_map(&mb_hi, b_hi, N, Kp, BN) ||
            !make_map(&mb_lo, b_lo, N, Kp, BN))
            return feo_fail(ctx, FEO_ERR_CUDA, "cuTensorMapEncodeTiled failed");
        const size_t tiles = feo_ceil_div(N, (size_t)BN) * feo_ceil_div(M, (size_t)BM);
        if (tiles > 0x7fffffffu) return feo_fail(ctx, FEO_ERR_UNSUPPORTED, "matmul grid too large");
        const dim3 grid((unsigned)tiles);
        matmul_tf32x3_kernel<<<grid, kThreads, kSmemBytes, ctx->stream>>>(ma_hi, ma_lo, mb_hi, mb_lo, (float *)C, (int)M, (int)N, num_kb, kc, ready,
                                                                          ready_need, kb_rot, ctx->peer_timeout_ns, ctx->status_flag);
        FEO_LAUNCH_CHECK(ctx);
    }
    if (overlap) FEO_CUDA(ctx, cudaStreamWaitEvent(ctx->stream, ctx->ev_join, 0));  // the workspace is reused by the next call
    for (int i = 1; i < n_streams; ++i) FEO_CUDA(ctx, cudaStreamWaitEvent(ctx->stream, ctx->pull_joins[i - 1], 0));
    return FEO_OK;
}

}  // namespace

int feo_impl_matmul_tf32x3(feo_ctx *ctx, void *C, const void *A, const void *B, size_t M, size_t K, size_t N) {
    return matmul_tf32x3_run(ctx, C, A, B, nullptr, nullptr, M, K, N);
}

int feo_impl_matmul_tf32x3_parts(feo_ctx *ctx, void *C, const void *A, const feo_peer::BParts &parts, const feo_peer::XchgParams &sync,
                                 size_t M, size_t K, size_t N) {
    return matmul_tf32x3_run(ctx, C, A, nullptr, &parts, &sync, M, K, N);
}
