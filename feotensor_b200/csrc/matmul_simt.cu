// matmul_simt.cu -- K6: DynamicMatrix::matmul (matrix.rs:76-90) on the SIMT pipes, every dtype.
//
// C[M,N] = A[M,K] * B[K,N], all dense row-major.  128x128 block tile, 8x8 register tile per thread
// (256 threads), BK-deep shared-memory panels double-buffered through registers, 128-bit global and
// shared accesses on the aligned fast path and guarded scalar accesses on ragged shapes; aligned shapes
// of every dtype take the cp.async-pipelined kernel further down (f32 with packed FFMA2).
// Floats accumulate with FMA in k order per thread (reassociation-free along k, but fused: within the
// stated tolerance of the reference's separate multiply and add); integers wrap, hence bit-exact.
// Bound: the FP32 pipe (2*M*N*K flop); this is the parity workhorse and the comparison point for
// the tcgen05 3xTF32 kernel in matmul_tf32x3.cu.
#include "feo_common.cuh"

namespace {

constexpr int BM = 128, BN = 128, TM = 8, TN = 8, PAD = 4;
constexpr int kThreads = 256;

template <typename T> __device__ __forceinline__ T mac(T acc, T a, T b) {
    if constexpr (std::is_same<T, float>::value) return fmaf(a, b, acc);
    else if constexpr (std::is_same<T, double>::value) return fma(a, b, acc);
    else return w_add(acc, w_mul(a, b));
}

template <typename T, bool FAST>
__global__ void __launch_bounds__(kThreads, (sizeof(T) == 4 ? 2 : 1)) matmul_simt_kernel(T *__restrict__ C, const T *__restrict__ A,
                                                               const T *__restrict__ B, long long M, long long K, long long N) {
    constexpr int VN = VecN<T>::value;        // elements per 128-bit access
    constexpr int BK = 4 * VN;                // 16 (4-byte) or 8 (8-byte): four vectors per A-tile row
    constexpr int A_VECS = BM * BK / VN / kThreads;  // vectors of the A tile per thread (= 2)
    constexpr int B_VECS = BK * BN / VN / kThreads;  // (= 2)
    using V = AVec<T, VN>;
    using F = AVec<T, 4>;
    __shared__ __align__(32) T As[2][BK][BM + PAD];  // transposed: As[k][m]
    __shared__ __align__(32) T Bs[2][BK][BN + PAD];

    const int tid = threadIdx.x;
    const int tx = tid % 16, ty = tid / 16;
    const long long m0 = (long long)blockIdx.y * BM, n0 = (long long)blockIdx.x * BN;

    V ra[A_VECS], rb[B_VECS];
    auto load_tiles = [&](long long k0) {
#pragma unroll
        for (int i = 0; i < A_VECS; ++i) {
            const int idx = tid + i * kThreads;
            const int row = idx / (BK / VN), cv = idx % (BK / VN);
            const long long gm = m0 + row, gk = k0 + cv * VN;
            if constexpr (FAST) {
                ra[i] = ld_keep(reinterpret_cast<const V *>(A + gm * K + gk));
            } else {
#pragma unroll
                for (int j = 0; j < VN; ++j) ra[i].v[j] = (gm < M && gk + j < K) ? A[gm * K + gk + j] : (T)0;
            }
        }
#pragma unroll
        for (int i = 0; i < B_VECS; ++i) {
            const int idx = tid + i * kThreads;
            const int row = idx / (BN / VN), cv = idx % (BN / VN);
            const long long gk = k0 + row, gn = n0 + cv * VN;
            if constexpr (FAST) {
                rb[i] = ld_keep(reinterpret_cast<const V *>(B + gk * N + gn));
            } else {
#pragma unroll
                for (int j = 0; j < VN; ++j) rb[i].v[j] = (gk < K && gn + j < N) ? B[gk * N + gn + j] : (T)0;
            }
        }
    };
    auto store_tiles = [&](int buf) {
#pragma unroll
        for (int i = 0; i < A_VECS; ++i) {
            const int idx = tid + i * kThreads;
            const int row = idx / (BK / VN), cv = idx % (BK / VN);
#pragma unroll
            for (int j = 0; j < VN; ++j) As[buf][cv * VN + j][row] = ra[i].v[j];
        }
#pragma unroll
        for (int i = 0; i < B_VECS; ++i) {
            const int idx = tid + i * kThreads;
            const int row = idx / (BN / VN), cv = idx % (BN / VN);
            *reinterpret_cast<V *>(&Bs[buf][row][cv * VN]) = rb[i];
        }
    };

    T acc[TM][TN];
#pragma unroll
    for (int i = 0; i < TM; ++i)
#pragma unroll
        for (int j = 0; j < TN; ++j) acc[i][j] = (T)0;

    const long long nkt = (K + BK - 1) / BK;
    load_tiles(0);
    store_tiles(0);
    __syncthreads();
    int buf = 0;
    for (long long kt = 0; kt < nkt; ++kt) {
        if (kt + 1 < nkt) load_tiles((kt + 1) * BK);
#pragma unroll
        for (int k = 0; k < BK; ++k) {
            const F a0 = *reinterpret_cast<const F *>(&As[buf][k][ty * 4]);
            const F a1 = *reinterpret_cast<const F *>(&As[buf][k][64 + ty * 4]);
            const F b0 = *reinterpret_cast<const F *>(&Bs[buf][k][tx * 4]);
            const F b1 = *reinterpret_cast<const F *>(&Bs[buf][k][64 + tx * 4]);
            T a[TM], b[TN];
#pragma unroll
            for (int j = 0; j < 4; ++j) { a[j] = a0.v[j]; a[4 + j] = a1.v[j]; b[j] = b0.v[j]; b[4 + j] = b1.v[j]; }
#pragma unroll
            for (int i = 0; i < TM; ++i)
#pragma unroll
                for (int j = 0; j < TN; ++j) acc[i][j] = mac(acc[i][j], a[i], b[j]);
        }
        if (kt + 1 < nkt) store_tiles(buf ^ 1);
        __syncthreads();
        buf ^= 1;
    }

#pragma unroll
    for (int i = 0; i < TM; ++i) {
        const long long gm = m0 + (i < 4 ? ty * 4 + i : 64 + ty * 4 + (i - 4));
#pragma unroll
        for (int h = 0; h < 2; ++h) {
            const long long gn = n0 + (h == 0 ? tx * 4 : 64 + tx * 4);
            if constexpr (FAST) {
                F r;
#pragma unroll
                for (int j = 0; j < 4; ++j) r.v[j] = acc[i][h * 4 + j];
                *reinterpret_cast<F *>(C + gm * N + gn) = r;
            } else {
                if (gm < M) {
#pragma unroll
                    for (int j = 0; j < 4; ++j)
                        if (gn + j < N) C[gm * N + gn + j] = acc[i][h * 4 + j];
                }
            }
        }
    }
}


// ---- aligned shapes: cp.async pipeline -------------------------------------------------------------------------
// The register-double-buffered kernel above leaves the order of "global load, compute, shared store" to the compiler,
// and SASS shows the loads of the next panels sunk to just in front of their shared-memory stores (fewer live
// registers), i.e. a full L2 round trip exposed per k-panel -- ncu: long_scoreboard the top stall, FMA pipe 2/3 busy.
// Here the panels go global -> shared with cp.async (LDGSTS, no registers), STAGES panels deep, one barrier per panel.
// A keeps its row-major layout in shared memory ([m][k], k contiguous: no transposing stores); a thread reads its 8 rows
// KV k-values at a time (a warp touches two distinct rows per access: broadcast, conflict-free at any row stride) and
// B as before.  Per-thread accumulation order along k is unchanged, so results are bit-identical to the kernel above.
__device__ __forceinline__ void cp_async16(void *smem, const void *gmem) {
    const unsigned s = (unsigned)__cvta_generic_to_shared(smem);
    asm volatile("cp.async.cg.shared.global [%0], [%1], 16;" ::"r"(s), "l"(gmem) : "memory");
}
__device__ __forceinline__ void cp_async_commit() { asm volatile("cp.async.commit_group;" ::: "memory"); }
template <int N> __device__ __forceinline__ void cp_async_wait() { asm volatile("cp.async.wait_group %0;" ::"n"(N) : "memory"); }

// Panel geometry in bytes is the same for 4- and 8-byte elements: an A row is 64 bytes (four 16-byte chunks: 16 or 8
// k-values), rows are padded by one chunk, and both panels are 512 chunks = two per thread.
template <typename T, int KM = 1> struct PipeGeom {
    static constexpr int VN = 16 / (int)sizeof(T);  // elements per 16-byte chunk
    static constexpr int PBK = 4 * VN * KM;         // k-panel depth: 16 (4-byte) or 8 (8-byte), times KM
    static constexpr int PAS = PBK + VN;            // A panel row stride (elements): 80 bytes
    static constexpr int PBS = BN + VN;             // B panel row stride (elements)
    static constexpr int CPR = BN / VN;             // chunks per B row: 32 or 64
    static constexpr int stage_elems = BM * PAS + PBK * PBS;
};
template <typename T, int STAGES, int KM = 1> constexpr int pipe_smem_bytes() { return STAGES * PipeGeom<T, KM>::stage_elems * (int)sizeof(T); }

// f32 only: two FMAs per instruction on 64-bit register pairs (FFMA2, new with sm_100).  Each lane is an IEEE fused
// multiply-add, so results are bit-identical to fmaf; what changes is the issue rate and the register-file reads per FMA.
__device__ __forceinline__ unsigned long long pack2(float lo, float hi) {
    unsigned long long r;
    asm("mov.b64 %0, {%1, %2};" : "=l"(r) : "f"(lo), "f"(hi));
    return r;
}
__device__ __forceinline__ void fma2(unsigned long long &acc, unsigned long long a, unsigned long long b) {
    asm("fma.rn.f32x2 %0, %1, %2, %0;" : "+l"(acc) : "l"(a), "l"(b));
}

// WL (warp layout): false = a warp is 2 x 16 threads of the 16 x 16 thread grid and a thread owns rows 4 ty .. 4 ty + 3 (+ 64);
// true = a warp is 4 x 8 threads and a thread owns rows ty + 16 i: a warp's `a` access then touches four rows 80 bytes apart
// (banks 0-3, 20-23, 8-11, 28-31) and its `b` access 128 contiguous bytes -- one shared-memory wavefront each, where the
// 2 x 16 shape needs two for every `b` access (256 bytes): 4 instead of 6 wavefronts per 32 FFMA2.
// KM: k-panel depth multiplier (KM = 2: 32 k-values per panel for 4-byte types, half as many CTA-wide barriers and cp.async
// wait points per flop; three stages of 35 KiB then fill the shared memory of two CTAs per SM).
template <typename T, int STAGES, int KV, bool F2, int MINB = 2, bool WL = false, int KM = 1>
__global__ void __launch_bounds__(kThreads, MINB) matmul_simt_pipe_kernel(T *__restrict__ C, const T *__restrict__ A, const T *__restrict__ B,
                                                                        long long M, long long K, long long N) {
    static_assert(!F2 || std::is_same<T, float>::value, "FFMA2 is f32 only");
    using PG = PipeGeom<T, KM>;
    constexpr int VN = PG::VN, PBK = PG::PBK, PAS = PG::PAS, PBS = PG::PBS, CPR = PG::CPR;
    static_assert(PBK % KV == 0 && KV * sizeof(T) <= 16, "a is read in 16-byte (or narrower) k-vectors");
    using F = AVec<T, 4>;
    using BV = AVec<T, VN>;  // one 16-byte chunk of a B row
    using KVec = AVec<T, KV>;
    extern __shared__ __align__(16) unsigned char pipe_smem[];
    T *As = reinterpret_cast<T *>(pipe_smem);  // [STAGES][BM][PAS]
    T *Bs = As + STAGES * BM * PAS;            // [STAGES][PBK][PBS]

    const int tid = threadIdx.x;
    const int tx = WL ? 8 * ((tid >> 5) & 1) + (tid & 7) : tid % 16;
    const int ty = WL ? 4 * (tid >> 6) + ((tid >> 3) & 3) : tid / 16;
    auto row_of = [&](int i) { return WL ? ty + 16 * i : (i < 4 ? ty * 4 + i : 64 + ty * 4 + (i - 4)); };  // i-th row of this thread's tile
    const long long m0 = (long long)blockIdx.y * BM, n0 = (long long)blockIdx.x * BN;
    // 2 * KM 16-byte chunks of each panel per thread: chunk c = tid + 256 i is (row c / ACH, k-chunk c % ACH) of the A panel
    // and (k-row c / CPR, column chunk c % CPR) of the B panel
    constexpr int ACH = PBK / VN;  // chunks per A row: 4 * KM
    auto issue = [&](long long kt, int stage) {
        T *ad = As + stage * (BM * PAS), *bd = Bs + stage * (PBK * PBS);
        const T *ag = A + m0 * K + kt * PBK, *bg = B + kt * PBK * N + n0;
#pragma unroll
        for (int i = 0; i < 2 * KM; ++i) {
            const int c = tid + i * kThreads;
            cp_async16(ad + (c / ACH) * PAS + (c % ACH) * VN, ag + (long long)(c / ACH) * K + (c % ACH) * VN);
        }
#pragma unroll
        for (int i = 0; i < 2 * KM; ++i) {
            const int c = tid + i * kThreads;
            cp_async16(bd + (c / CPR) * PBS + (c % CPR) * VN, bg + (long long)(c / CPR) * N + (c % CPR) * VN);
        }
    };

    T acc[TM][TN];
    unsigned long long acc2[TM][TN / 2];  // F2: acc2[i][h] = (C[i][2h], C[i][2h+1]) as one 64-bit pair
#pragma unroll
    for (int i = 0; i < TM; ++i)
#pragma unroll
        for (int j = 0; j < TN; ++j) {
            acc[i][j] = (T)0;
            acc2[i][j / 2] = 0ull;
        }

    const long long nkt = K / PBK;
#pragma unroll
    for (int s = 0; s < STAGES - 1; ++s) {
        if (s < nkt) issue(s, s);
        cp_async_commit();
    }
    int stage = 0;
    for (long long kt = 0; kt < nkt; ++kt) {
        cp_async_wait<STAGES - 2>();  // panel kt has landed (this thread's chunks) ...
        __syncthreads();              // ... and everyone's; everyone is also done reading the stage refilled next
        {
            const long long nx = kt + STAGES - 1;
            int ns = stage + STAGES - 1;
            if (ns >= STAGES) ns -= STAGES;
            if (nx < nkt) issue(nx, ns);
            cp_async_commit();
        }
        const T *as = As + stage * (BM * PAS);
        const T *bs = Bs + stage * (PBK * PBS) + tx * 4;
#pragma unroll
        for (int kk = 0; kk < PBK; kk += KV) {
            KVec a[TM];
#pragma unroll
            for (int i = 0; i < TM; ++i) a[i] = *reinterpret_cast<const KVec *>(as + row_of(i) * PAS + kk);
#pragma unroll
            for (int k = 0; k < KV; ++k) {
                if constexpr (F2) {
                    const ulonglong2 p0 = *reinterpret_cast<const ulonglong2 *>(bs + (kk + k) * PBS);
                    const ulonglong2 p1 = *reinterpret_cast<const ulonglong2 *>(bs + (kk + k) * PBS + 64);
                    const unsigned long long b2[TN / 2] = {p0.x, p0.y, p1.x, p1.y};
#pragma unroll
                    for (int i = 0; i < TM; ++i) {
                        const unsigned long long a2 = pack2(a[i].v[k], a[i].v[k]);
#pragma unroll
                        for (int h = 0; h < TN / 2; ++h) fma2(acc2[i][h], a2, b2[h]);
                    }
                } else {
                    T b[TN];  // columns tx*4 .. +3 and 64 + tx*4 .. +3, in 16-byte chunks
#pragma unroll
                    for (int c = 0; c < TN / VN; ++c) {
                        constexpr int half = (TN / VN) / 2;
                        const BV t = *reinterpret_cast<const BV *>(bs + (kk + k) * PBS + (c < half ? c * VN : 64 + (c - half) * VN));
#pragma unroll
                        for (int j = 0; j < VN; ++j) b[c * VN + j] = t.v[j];
                    }
#pragma unroll
                    for (int i = 0; i < TM; ++i)
#pragma unroll
                        for (int j = 0; j < TN; ++j) acc[i][j] = mac(acc[i][j], a[i].v[k], b[j]);
                }
            }
        }
        if (++stage == STAGES) stage = 0;
    }
    cp_async_wait<0>();
    if constexpr (F2) {
#pragma unroll
        for (int i = 0; i < TM; ++i)
#pragma unroll
            for (int h = 0; h < TN / 2; ++h)
                asm("mov.b64 {%0, %1}, %2;" : "=f"(acc[i][2 * h]), "=f"(acc[i][2 * h + 1]) : "l"(acc2[i][h]));
    }

#pragma unroll
    for (int i = 0; i < TM; ++i) {
        const long long gm = m0 + row_of(i);
#pragma unroll
        for (int h = 0; h < 2; ++h) {
            const long long gn = n0 + (h == 0 ? tx * 4 : 64 + tx * 4);
            F r;
#pragma unroll
            for (int j = 0; j < 4; ++j) r.v[j] = acc[i][h * 4 + j];
            *reinterpret_cast<F *>(C + gm * N + gn) = r;
        }
    }
}

template <typename T, int STAGES, int KV, bool F2, int MINB = 2, bool WL = false, int KM = 1>
int launch_pipe(feo_ctx *ctx, dim3 grid, T *C, const T *A, const T *B, size_t M, size_t K, size_t N) {
    static bool attr_set[64] = {false};  // the opt-in is per device (and per instantiation)
    const int dev = ctx->device & 63;
    constexpr int smem = pipe_smem_bytes<T, STAGES, KM>();
    if (!attr_set[dev]) {
        FEO_CUDA(ctx, cudaFuncSetAttribute(matmul_simt_pipe_kernel<T, STAGES, KV, F2, MINB, WL, KM>, cudaFuncAttributeMaxDynamicSharedMemorySize, smem));
        attr_set[dev] = true;
    }
    matmul_simt_pipe_kernel<T, STAGES, KV, F2, MINB, WL, KM><<<grid, kThreads, smem, ctx->stream>>>(C, A, B, (long long)M, (long long)K, (long long)N);
    FEO_LAUNCH_CHECK(ctx);
    return FEO_OK;
}

template <typename T>
int matmul_typed(feo_ctx *ctx, T *C, const T *A, const T *B, size_t M, size_t K, size_t N) {
    constexpr int VN = VecN<T>::value;
    constexpr int BK = 4 * VN;
    const dim3 grid((unsigned)feo_ceil_div(N, (size_t)BN), (unsigned)feo_ceil_div(M, (size_t)BM));
    if (grid.y > 65535) return feo_fail(ctx, FEO_ERR_UNSUPPORTED, "matmul M too large for the SIMT grid");
    const bool fast = (M % BM == 0) && (N % BN == 0) && (K % BK == 0) && feo_aligned16(A) && feo_aligned16(B) && feo_aligned16(C) &&
                      (sizeof(T) == 4 || ((reinterpret_cast<uintptr_t>(C) & 31u) == 0));
    // knob MATMUL_SIMT: 0 = auto (cp.async pipeline; f32 with FFMA2), 1 = register double buffering,
    // 2 = pipeline with scalar FMAs, 3..6 = stage / k-vector / occupancy variants, 7..10 = the 4 x 8 warp shape,
    // 11, 12 = 32-deep k-panels (12 = auto for K % 32 == 0), 13 = round 1's default (16-deep, 4 stages, 2 x 16 warps)
    const int v = ctx->knobs[FEO_KNOB_MATMUL_SIMT];
    if constexpr (sizeof(T) == 4) {
        constexpr bool kF = std::is_same<T, float>::value;
        if (fast && v != 1) {
            switch (v) {
                case 2: return launch_pipe<T, 4, 4, false>(ctx, grid, C, A, B, M, K, N);
                case 3: return launch_pipe<T, 3, 4, kF>(ctx, grid, C, A, B, M, K, N);
                case 4: return launch_pipe<T, 4, 2, kF>(ctx, grid, C, A, B, M, K, N);
                case 5: return launch_pipe<T, 5, 4, kF>(ctx, grid, C, A, B, M, K, N);
                case 6: return launch_pipe<T, 4, 4, kF, 1>(ctx, grid, C, A, B, M, K, N);
                case 7: return launch_pipe<T, 4, 4, kF, 2, true>(ctx, grid, C, A, B, M, K, N);
                case 8: return launch_pipe<T, 4, 2, kF, 2, true>(ctx, grid, C, A, B, M, K, N);
                case 9: return launch_pipe<T, 4, 4, kF, 1, true>(ctx, grid, C, A, B, M, K, N);
                case 10: return launch_pipe<T, 3, 4, kF, 2, true>(ctx, grid, C, A, B, M, K, N);
                case 11: if (K % 32 == 0) return launch_pipe<T, 3, 4, kF, 2, false, 2>(ctx, grid, C, A, B, M, K, N); break;
                case 12: if (K % 32 == 0) return launch_pipe<T, 3, 4, kF, 2, true, 2>(ctx, grid, C, A, B, M, K, N); break;
                case 13: return launch_pipe<T, 4, 4, kF>(ctx, grid, C, A, B, M, K, N);  // round 1's default
                default:
                    // 32-deep k-panels (half the CTA-wide barriers per flop), three stages, the 4 x 8 warp shape: 51.3 -> 54.3
                    // TFLOP/s f32, 33.7 -> 34.1 Tmac/s i32 at 8192^3; K % 32 != 0 keeps the 16-deep panels
                    if (K % 32 == 0) return launch_pipe<T, 3, 4, kF, 2, true, 2>(ctx, grid, C, A, B, M, K, N);
                    return launch_pipe<T, 4, 4, kF>(ctx, grid, C, A, B, M, K, N);
            }
        }
    } else {
        // 8-byte types: 128 accumulator registers, one CTA per SM -- eight warps have nothing to hide an exposed load behind
        if (fast && v != 1) {
            switch (v) {
                case 3: return launch_pipe<T, 3, 2, false, 1>(ctx, grid, C, A, B, M, K, N);
                case 4: return launch_pipe<T, 4, 1, false, 1>(ctx, grid, C, A, B, M, K, N);
                case 11: if (K % 16 == 0) return launch_pipe<T, 4, 2, false, 1, false, 2>(ctx, grid, C, A, B, M, K, N); break;
                case 12: if (K % 32 == 0) return launch_pipe<T, 3, 2, false, 1, false, 4>(ctx, grid, C, A, B, M, K, N); break;
                case 7: return launch_pipe<T, 4, 2, false, 1, true>(ctx, grid, C, A, B, M, K, N);
                case 8: if (K % 16 == 0) return launch_pipe<T, 4, 2, false, 1, true, 2>(ctx, grid, C, A, B, M, K, N); break;
                case 13: return launch_pipe<T, 4, 2, false, 1>(ctx, grid, C, A, B, M, K, N);  // round 1's default
                default:
                    // deeper k-panels = fewer CTA-wide barriers per flop: f64 8192^3 22.6 -> 25.0 TFLOP/s with 32-deep panels
                    if (K % 32 == 0) return launch_pipe<T, 3, 2, false, 1, false, 4>(ctx, grid, C, A, B, M, K, N);
                    if (K % 16 == 0) return launch_pipe<T, 4, 2, false, 1, false, 2>(ctx, grid, C, A, B, M, K, N);
                    return launch_pipe<T, 4, 2, false, 1>(ctx, grid, C, A, B, M, K, N);
            }
        }
    }
    if (fast) matmul_simt_kernel<T, true><<<grid, kThreads, 0, ctx->stream>>>(C, A, B, (long long)M, (long long)K, (long long)N);
    else matmul_simt_kernel<T, false><<<grid, kThreads, 0, ctx->stream>>>(C, A, B, (long long)M, (long long)K, (long long)N);
    FEO_LAUNCH_CHECK(ctx);
    return FEO_OK;
}

}  // namespace

int feo_impl_matmul_simt(feo_ctx *ctx, int dtype, void *C, const void *A, const void *B, size_t M, size_t K, size_t N) {
    FEO_DISPATCH(ctx, dtype, T, return matmul_typed<T>(ctx, (T *)C, (const T *)A, (const T *)B, M, K, N));
    return FEO_OK;
}
