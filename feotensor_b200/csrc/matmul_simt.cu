// matmul_simt.cu -- K6: DynamicMatrix::matmul (matrix.rs:76-90) on the SIMT pipes, every dtype.
//
// C[M,N] = A[M,K] * B[K,N], all dense row-major.  128x128 block tile, 8x8 register tile per thread
// (256 threads), BK-deep shared-memory panels double-buffered through registers, 128-bit global and
// shared accesses on the aligned fast path and guarded scalar accesses on ragged shapes.
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

template <typename T>
int matmul_typed(feo_ctx *ctx, T *C, const T *A, const T *B, size_t M, size_t K, size_t N) {
    constexpr int VN = VecN<T>::value;
    constexpr int BK = 4 * VN;
    const dim3 grid((unsigned)feo_ceil_div(N, (size_t)BN), (unsigned)feo_ceil_div(M, (size_t)BM));
    if (grid.y > 65535) return feo_fail(ctx, FEO_ERR_UNSUPPORTED, "matmul M too large for the SIMT grid");
    const bool fast = (M % BM == 0) && (N % BN == 0) && (K % BK == 0) && feo_aligned16(A) && feo_aligned16(B) && feo_aligned16(C) &&
                      (sizeof(T) == 4 || ((reinterpret_cast<uintptr_t>(C) & 31u) == 0));
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
