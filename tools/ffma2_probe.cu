// ffma2_probe.cu -- what the FP32 pipe of sm_100 sustains from registers alone, in the operand forms the SIMT matmul can use.
// No memory traffic in the loop: 8 x 4 accumulator pairs per thread (the matmul's register tile), per k-step 8 `a` scalars
// and 4 `b` pairs, 32 FFMA2 (or 64 FFMA).  Reports cycles per FFMA2 per SM sub-partition (2.0 = the pipe's peak) for
//   0: fma.rn.f32x2 with `a` as the scalar-broadcast operand (what the matmul kernel compiles to)
//   1: fma.rn.f32x2 with `a` pre-packed as a pair (a, a)
//   2: plain fmaf, 64 per k-step
// at 8 and 16 resident warps per SM sub-partition... (blocks of 256 threads, 2 or 4 per SM).
//     nvcc -gencode arch=compute_100a,code=sm_100a -O3 -o build/ffma2_probe tools/ffma2_probe.cu && build/ffma2_probe
#include <cstdio>
#include <cuda_runtime.h>

__device__ __forceinline__ unsigned long long pack2(float lo, float hi) {
    unsigned long long r;
    asm("mov.b64 %0, {%1, %2};" : "=l"(r) : "f"(lo), "f"(hi));
    return r;
}
__device__ __forceinline__ void fma2(unsigned long long &acc, unsigned long long a, unsigned long long b) {
    asm("fma.rn.f32x2 %0, %1, %2, %0;" : "+l"(acc) : "l"(a), "l"(b));
}

template <int MODE>
__global__ void __launch_bounds__(256, 2) probe(float *out, const float *in, int iters) {
    float a[8];
    unsigned long long ap[8];  // MODE 1: `a` as genuine 64-bit pairs (distinct halves, so no scalar-broadcast form is possible)
    unsigned long long b2[4], acc2[8][4];
    float acc[8][8];
    for (int i = 0; i < 8; ++i) { a[i] = in[threadIdx.x + 32 * i]; ap[i] = pack2(in[threadIdx.x + 32 * i], in[threadIdx.x + 32 * i + 1]); }
    for (int h = 0; h < 4; ++h) b2[h] = pack2(in[threadIdx.x + 7 * h], in[threadIdx.x + 11 * h + 1]);
    for (int i = 0; i < 8; ++i)
        for (int h = 0; h < 4; ++h) { acc2[i][h] = 0ull; acc[i][2 * h] = 0.f; acc[i][2 * h + 1] = 0.f; }
    for (int it = 0; it < iters; ++it) {
#pragma unroll
        for (int k = 0; k < 4; ++k) {
#pragma unroll
            for (int i = 0; i < 8; ++i) {
                if constexpr (MODE == 2) {
                    float bl[8];
#pragma unroll
                    for (int h = 0; h < 4; ++h) asm("mov.b64 {%0, %1}, %2;" : "=f"(bl[2 * h]), "=f"(bl[2 * h + 1]) : "l"(b2[h]));
#pragma unroll
                    for (int j = 0; j < 8; ++j) acc[i][j] = fmaf(a[i], bl[j], acc[i][j]);
                } else {
                    const unsigned long long a2 = MODE == 1 ? ap[i] : pack2(a[i], a[i]);
#pragma unroll
                    for (int h = 0; h < 4; ++h) fma2(acc2[i][h], a2, b2[h]);
                }
            }
            // rotate the operands so that nothing is loop invariant (keeps ptxas from hoisting or folding anything)
            const float t = a[0];
#pragma unroll
            for (int i = 0; i < 7; ++i) a[i] = a[i + 1];
            a[7] = t;
            if constexpr (MODE == 1) {
                const unsigned long long u = ap[0];
#pragma unroll
                for (int i = 0; i < 7; ++i) ap[i] = ap[i + 1];
                ap[7] = u;
            }
        }
    }
    float s = 0.f;
    for (int i = 0; i < 8; ++i)
        for (int h = 0; h < 4; ++h) {
            float lo, hi;
            asm("mov.b64 {%0, %1}, %2;" : "=f"(lo), "=f"(hi) : "l"(acc2[i][h]));
            s += lo + hi + acc[i][2 * h] + acc[i][2 * h + 1];
        }
    out[blockIdx.x * 256 + threadIdx.x] = s;
}

template <int MODE> void run(const char *name, int blocks_per_sm, float *out, const float *in) {
    int sms = 0;
    cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, 0);
    int clock_khz = 0;
    cudaDeviceGetAttribute(&clock_khz, cudaDevAttrClockRate, 0);
    const int iters = 1 << 14;
    cudaEvent_t e0, e1;
    cudaEventCreate(&e0);
    cudaEventCreate(&e1);
    probe<MODE><<<sms * blocks_per_sm, 256>>>(out, in, 64);
    cudaEventRecord(e0);
    probe<MODE><<<sms * blocks_per_sm, 256>>>(out, in, iters);
    cudaEventRecord(e1);
    cudaEventSynchronize(e1);
    float ms = 0;
    cudaEventElapsedTime(&ms, e0, e1);
    // per SM sub-partition: warps x iters x 4 k-steps x 32 FFMA2 (MODE 2: 64 FFMA = 32 FFMA2-equivalents)
    const double ffma2 = (double)blocks_per_sm * 8 / 4 * iters * 4 * 32;
    const double cycles = ms * 1e-3 * clock_khz * 1e3;
    const double tflops = (double)sms * blocks_per_sm * 256 * (double)iters * 4 * 64 * 2 / (ms * 1e-3) / 1e12;
    printf("{\"probe\": \"%s\", \"blocks_per_sm\": %d, \"ms\": %.3f, \"cycles_per_ffma2_at_max_clock\": %.3f, \"TFLOP/s\": %.1f, \"err\": \"%s\"}\n", name,
           blocks_per_sm, ms, cycles / ffma2, tflops, cudaGetErrorString(cudaGetLastError()));
}

int main() {
    float *out, *in;
    cudaMalloc(&out, 148 * 4 * 256 * sizeof(float));
    cudaMalloc(&in, 4096 * sizeof(float));
    cudaMemset(in, 0, 4096 * sizeof(float));
    for (int bps = 1; bps <= 2; ++bps) {
        run<0>("ffma2_scalar_a", bps, out, in);
        run<1>("ffma2_pair_a", bps, out, in);
        run<2>("ffma", bps, out, in);
    }
    return 0;
}
