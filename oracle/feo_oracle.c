/* feo_oracle.c -- CPU oracle for the feotensor hot path.
 *
 * TEST INFRASTRUCTURE, NOT PRODUCT.  Only tests/, __graft_entry__.smoke() and bench.py's
 * cpu_baseline / --impl reference legs may load this library.  The product (libfeotensor_b200.so)
 * never links, loads or calls it and has no CPU fallback.
 *
 * What it is: a plain-C restatement of the reference's algorithms (Rust crate feotensor,
 * /root/reference/src/{tensor,matrix,vector,storage,shape,iter,coordinate}.rs), following the
 * reference's loop order and accumulation order so that results are bit-identical to what the Rust
 * code produces (the Rust toolchain is not available in this image, so the reference itself cannot
 * be compiled).  Each function cites the reference lines it follows.
 *
 * Parity pin: every known-answer test the reference holds for this path (149 inline #[test]s; the
 * hot-path ones are transcribed in tests/golden/reference_kat.json) is replayed against this
 * oracle by tests/test_oracle_golden.py.
 *
 * Build: gcc -O2 -ffp-contract=off -fno-fast-math -shared -fPIC (see oracle/Makefile).
 * dtype codes: 0 f32, 1 f64, 2 i32, 3 i64, 4 u32.  op codes: 0 add, 1 sub, 2 mul, 3 div.
 * reduce kinds: 0 sum, 1 mean, 2 var, 3 max, 4 min.
 */
#include <math.h>
#include <stddef.h>
#include <stdint.h>
#include <stdlib.h>
#include <string.h>

#define ORACLE_OK 0
#define ORACLE_ERR_SHAPE 1      /* a ShapeError, or a precondition the reference panics on */
#define ORACLE_ERR_UNSUPPORTED 3
#define ORACLE_ERR_ARITH 4      /* integer /0 or MIN/-1: Rust panics */
#define ORACLE_ERR_NOMEM 5
#define ORACLE_MAX_RANK 16

#define T float
#define SFX f32
#define IS_FLOAT 1
#include "oracle_impl.inc"
#undef T
#undef SFX
#undef IS_FLOAT

#define T double
#define SFX f64
#define IS_FLOAT 1
#include "oracle_impl.inc"
#undef T
#undef SFX
#undef IS_FLOAT

#define T int32_t
#define UT uint32_t
#define SFX i32
#define IS_FLOAT 0
#define IS_SIGNED 1
#define T_MIN INT32_MIN
#include "oracle_impl.inc"
#undef T
#undef UT
#undef SFX
#undef IS_FLOAT
#undef IS_SIGNED
#undef T_MIN

#define T int64_t
#define UT uint64_t
#define SFX i64
#define IS_FLOAT 0
#define IS_SIGNED 1
#define T_MIN INT64_MIN
#include "oracle_impl.inc"
#undef T
#undef UT
#undef SFX
#undef IS_FLOAT
#undef IS_SIGNED
#undef T_MIN

#define T uint32_t
#define UT uint32_t
#define SFX u32
#define IS_FLOAT 0
#define IS_SIGNED 0
#include "oracle_impl.inc"
#undef T
#undef UT
#undef SFX
#undef IS_FLOAT
#undef IS_SIGNED

#define DISPATCH(dtype, CALL)                                        \
    switch (dtype) {                                                 \
    case 0: { CALL(f32, float); }                                    \
    case 1: { CALL(f64, double); }                                   \
    case 2: { CALL(i32, int32_t); }                                  \
    case 3: { CALL(i64, int64_t); }                                  \
    case 4: { CALL(u32, uint32_t); }                                 \
    default: return ORACLE_ERR_UNSUPPORTED;                          \
    }

size_t feo_oracle_dtype_size(int dtype) {
    switch (dtype) { case 0: return 4; case 1: return 8; case 2: return 4; case 3: return 8; case 4: return 4; }
    return 0;
}

/* shape.rs:12-17 -- any zero dimension is a ShapeError; empty dims are allowed. */
int feo_oracle_shape_check(int order, const size_t *shape) {
    for (int d = 0; d < order; ++d) if (shape[d] == 0) return ORACLE_ERR_SHAPE;
    return ORACLE_OK;
}

/* shape.rs:19-21 */
size_t feo_oracle_shape_size(int order, const size_t *shape) {
    size_t s = 1;
    for (int d = 0; d < order; ++d) s *= shape[d];
    return s;
}

/* storage.rs:18-38 -- row-major flatten.  Returns 0 ok; 1 order mismatch (:19-22);
 * 2 + i when coord[i] >= shape[i] (:24-30, "out of bounds for dimension i"). */
int feo_oracle_flatten(const size_t *coord, int corder, const size_t *shape, int order, size_t *index) {
    if (corder != order) return 1;
    for (int i = 0; i < corder; ++i) if (coord[i] >= shape[i]) return 2 + i;
    size_t idx = 0;
    for (int k = 0; k < order; ++k) {
        size_t stride = 1;
        for (int m = k + 1; m < order; ++m) stride *= shape[m];
        idx += coord[k] * stride;
    }
    *index = idx;
    return 0;
}

/* iter.rs:27-47 -- all coordinates of `shape` in row-major order, written as rows of `order`
 * entries into out (size(shape) * order entries).  Order-0 shapes yield nothing (iter.rs:28). */
size_t feo_oracle_index_iterator(int order, const size_t *shape, size_t *out) {
    if (order == 0) return 0;
    size_t cur[ORACLE_MAX_RANK] = {0};
    size_t count = 0;
    int done = 0;
    while (!done) {
        memcpy(out + count * (size_t)order, cur, (size_t)order * sizeof(size_t));
        count++;
        for (int i = order - 1; i >= 0; --i) {
            if (cur[i] + 1 < shape[i]) { cur[i] += 1; break; }
            cur[i] = 0;
            if (i == 0) done = 1;
        }
    }
    return count;
}

/* Axes the reference handles without panicking: strictly ascending, unique, in range
 * (tensor.rs:77-81,97-100 with coordinate.rs:27-33). */
int feo_oracle_axes_check(int order, int naxes, const size_t *axes) {
    for (int i = 0; i < naxes; ++i) {
        if (axes[i] >= (size_t)order) return ORACLE_ERR_SHAPE;
        if (i > 0 && axes[i] <= axes[i - 1]) return ORACLE_ERR_SHAPE;
    }
    return ORACLE_OK;
}

/* Result shape of a reduction: kept dims ascending (tensor.rs:72-80); [1] on the scalar path
 * (tensor.rs:84-87).  Returns the order written to out_shape. */
int feo_oracle_reduce_shape(int order, const size_t *shape, int naxes, const size_t *axes, size_t *out_shape) {
    int nk = 0;
    for (int d = 0; d < order; ++d) {
        int removed = 0;
        for (int i = 0; i < naxes; ++i) if (axes[i] == (size_t)d) removed = 1;
        if (!removed) out_shape[nk++] = shape[d];
    }
    if (naxes == 0 || nk == 0) { out_shape[0] = 1; return 1; }
    return nk;
}

int feo_oracle_ewise(int dtype, int op, void *out, const void *a, const void *b, size_t n) {
#define CALL(S, TT) return ewise_##S(op, (TT *)out, (const TT *)a, (const TT *)b, n)
    DISPATCH(dtype, CALL)
#undef CALL
}

int feo_oracle_ewise_faithful(int dtype, int op, void *out, const void *a, const void *b, size_t n) {
#define CALL(S, TT) return ewise_faithful_##S(op, (TT *)out, (const TT *)a, (const TT *)b, n)
    DISPATCH(dtype, CALL)
#undef CALL
}

int feo_oracle_ewise_scalar(int dtype, int op, void *out, const void *a, const void *scalar, size_t n) {
#define CALL(S, TT) return ewise_scalar_##S(op, (TT *)out, (const TT *)a, *(const TT *)scalar, n)
    DISPATCH(dtype, CALL)
#undef CALL
}

int feo_oracle_ewise_broadcast(int dtype, int op, void *out, const void *a, const void *b, int rank,
                               const size_t *shape, const size_t *sa, const size_t *sb) {
    if (rank > ORACLE_MAX_RANK) return ORACLE_ERR_UNSUPPORTED;
#define CALL(S, TT) return ewise_broadcast_##S(op, (TT *)out, (const TT *)a, (const TT *)b, rank, shape, sa, sb)
    DISPATCH(dtype, CALL)
#undef CALL
}

int feo_oracle_materialize(int dtype, void *dst, const void *src, int rank, const size_t *shape,
                           const size_t *strides) {
    if (rank > ORACLE_MAX_RANK) return ORACLE_ERR_UNSUPPORTED;
#define CALL(S, TT) materialize_##S((TT *)dst, (const TT *)src, rank, shape, strides); return ORACLE_OK
    DISPATCH(dtype, CALL)
#undef CALL
}

int feo_oracle_count_in_type(int dtype, size_t count, void *out) {
#define CALL(S, TT) *(TT *)out = count_in_t_##S(count); return ORACLE_OK
    DISPATCH(dtype, CALL)
#undef CALL
}

int feo_oracle_reduce_divisor(int dtype, int order, const size_t *shape, int naxes, const size_t *axes, void *out) {
#define CALL(S, TT) *(TT *)out = reduce_divisor_##S(order, shape, naxes, axes); return ORACLE_OK
    DISPATCH(dtype, CALL)
#undef CALL
}

int feo_oracle_reduce(int dtype, int kind, void *out, const void *x, int order, const size_t *shape,
                      int naxes, const size_t *axes) {
    if (order > ORACLE_MAX_RANK) return ORACLE_ERR_UNSUPPORTED;
    if (feo_oracle_shape_check(order, shape) != ORACLE_OK) return ORACLE_ERR_SHAPE;
    if (feo_oracle_axes_check(order, naxes, axes) != ORACLE_OK) return ORACLE_ERR_SHAPE;
#define CALL(S, TT) return reduce_##S(kind, (TT *)out, (const TT *)x, order, shape, naxes, axes)
    DISPATCH(dtype, CALL)
#undef CALL
}

int feo_oracle_reduce_faithful(int dtype, int kind, void *out, const void *x, int order,
                               const size_t *shape, int naxes, const size_t *axes,
                               int recompute_mean_per_target) {
    if (order > ORACLE_MAX_RANK) return ORACLE_ERR_UNSUPPORTED;
    if (feo_oracle_shape_check(order, shape) != ORACLE_OK) return ORACLE_ERR_SHAPE;
    if (feo_oracle_axes_check(order, naxes, axes) != ORACLE_OK) return ORACLE_ERR_SHAPE;
#define CALL(S, TT) return reduce_faithful_##S(kind, (TT *)out, (const TT *)x, order, shape, naxes, axes, recompute_mean_per_target)
    DISPATCH(dtype, CALL)
#undef CALL
}

int feo_oracle_prod(int dtype, void *out, const void *a, size_t na, const void *b, size_t nb) {
#define CALL(S, TT) prod_##S((TT *)out, (const TT *)a, na, (const TT *)b, nb); return ORACLE_OK
    DISPATCH(dtype, CALL)
#undef CALL
}

int feo_oracle_dot(int dtype, void *out, const void *a, const void *b, size_t n) {
#define CALL(S, TT) *(TT *)out = dot_##S((const TT *)a, (const TT *)b, n); return ORACLE_OK
    DISPATCH(dtype, CALL)
#undef CALL
}

int feo_oracle_matvec(int dtype, void *out, const void *A, const void *v, size_t M, size_t K) {
#define CALL(S, TT) matvec_##S((TT *)out, (const TT *)A, (const TT *)v, M, K); return ORACLE_OK
    DISPATCH(dtype, CALL)
#undef CALL
}

int feo_oracle_vecmat(int dtype, void *out, const void *v, const void *A, size_t K, size_t N) {
#define CALL(S, TT) vecmat_##S((TT *)out, (const TT *)v, (const TT *)A, K, N); return ORACLE_OK
    DISPATCH(dtype, CALL)
#undef CALL
}

int feo_oracle_matmul(int dtype, void *C, const void *A, const void *B, size_t M, size_t K, size_t N) {
#define CALL(S, TT) matmul_##S((TT *)C, (const TT *)A, (const TT *)B, M, K, N); return ORACLE_OK
    DISPATCH(dtype, CALL)
#undef CALL
}

int feo_oracle_matmul_faithful(int dtype, void *C, const void *A, const void *B, size_t M, size_t K, size_t N) {
#define CALL(S, TT) matmul_faithful_##S((TT *)C, (const TT *)A, (const TT *)B, M, K, N); return ORACLE_OK
    DISPATCH(dtype, CALL)
#undef CALL
}

int feo_oracle_matmul_entries(int dtype, void *out, const void *A, const void *B, size_t M, size_t K,
                              size_t N, const size_t *rows, const size_t *cols, size_t ns) {
#define CALL(S, TT) matmul_entries_##S((TT *)out, (const TT *)A, (const TT *)B, M, K, N, rows, cols, ns); return ORACLE_OK
    DISPATCH(dtype, CALL)
#undef CALL
}

/* Host twin of the library's synthetic-data generator (feo_fill_uniform in include/feotensor_b200.h):
 * element i = f(seed, first_index + i).  Not part of the reference; it lets tests regenerate any
 * element of a full-size device tensor without downloading it. */
static inline uint64_t oracle_mix64(uint64_t z) {
    z += 0x9E3779B97F4A7C15ULL;
    z = (z ^ (z >> 30)) * 0xBF58476D1CE4E5B9ULL;
    z = (z ^ (z >> 27)) * 0x94D049BB133111EBULL;
    return z ^ (z >> 31);
}
int feo_oracle_fill_uniform(int dtype, void *out, size_t n, uint64_t seed, uint64_t first_index, size_t stride,
                            double lo, double hi) {
    /* writes n elements, element k taken at global index first_index + k * stride */
    for (size_t k = 0; k < n; ++k) {
        const uint64_t h = oracle_mix64(seed * 0xD1342543DE82EF95ULL + (first_index + (uint64_t)k * stride));
        switch (dtype) {
        case 0: { const float l = (float)lo, span = (float)((float)hi - (float)lo);
                  const float u = (float)(h >> 40) * (1.0f / 16777216.0f);
                  const float m = span * u; ((float *)out)[k] = l + m; break; }
        case 1: { const double span = hi - lo; const double u = (double)(h >> 11) * (1.0 / 9007199254740992.0);
                  const double m = span * u; ((double *)out)[k] = lo + m; break; }
        case 2: { const uint64_t r = (uint64_t)((int64_t)hi - (int64_t)lo); ((int32_t *)out)[k] = (int32_t)((uint32_t)(int32_t)lo + (uint32_t)(h % r)); break; }
        case 3: { const uint64_t r = (uint64_t)((int64_t)hi - (int64_t)lo); ((int64_t *)out)[k] = (int64_t)((uint64_t)(int64_t)lo + (h % r)); break; }
        case 4: { const uint64_t r = (uint64_t)((int64_t)hi - (int64_t)lo); ((uint32_t *)out)[k] = (uint32_t)lo + (uint32_t)(h % r); break; }
        default: return ORACLE_ERR_UNSUPPORTED;
        }
    }
    return ORACLE_OK;
}

/* tensor.rs:325-333 -- result.data[i] = self.data[i].powf(power); Rust's powf is the platform libm's (glibc here). */
int feo_oracle_pow(int dtype, void *out, const void *a, const void *power, size_t n) {
    if (dtype == 0) { const float p = *(const float *)power; for (size_t i = 0; i < n; ++i) ((float *)out)[i] = powf(((const float *)a)[i], p); return ORACLE_OK; }
    if (dtype == 1) { const double p = *(const double *)power; for (size_t i = 0; i < n; ++i) ((double *)out)[i] = pow(((const double *)a)[i], p); return ORACLE_OK; }
    return ORACLE_ERR_UNSUPPORTED;
}
