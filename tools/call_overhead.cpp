// Per-call cost of the C ABI from a compiled host program, at BASELINE config 1 sizes (f32 [64,128,256], 8 MiB): the calls
// are launch-bound there, and the Python tools measure ctypes marshalling on top of the library's own planning + launch.
//
//   g++ -std=c++17 -O2 -Iinclude tools/call_overhead.cpp -Lfeotensor_b200/lib -lfeotensor_b200
//       -Wl,-rpath,$PWD/feotensor_b200/lib -o build/call_overhead   (one line);  build/call_overhead [iters]
//
// Prints, per operation, the host time to ISSUE one call (the stream is never waited on inside the loop) and the time per
// call including the final feo_sync -- whichever is larger is what a tight loop of such calls runs at.
#include <chrono>
#include <cstdio>
#include <cstdlib>
#include <functional>

#include "feotensor_b200.h"

static void must(feo_ctx *ctx, int rc, const char *what) {
    if (rc != FEO_OK) {
        std::fprintf(stderr, "%s failed (%d): %s\n", what, rc, ctx ? feo_last_error(ctx) : "no context");
        std::exit(1);
    }
}

int main(int argc, char **argv) {
    const int iters = argc > 1 ? std::atoi(argv[1]) : 2000;
    feo_ctx *ctx = nullptr;
    must(nullptr, feo_init(0, &ctx), "feo_init");
    const size_t shape[3] = {64, 128, 256}, axes02[2] = {0, 2}, axes0[1] = {0};
    const size_t n = 64 * 128 * 256;
    void *x = nullptr, *y = nullptr, *z = nullptr, *r = nullptr;
    must(ctx, feo_malloc(ctx, n * 4, &x), "feo_malloc");
    must(ctx, feo_malloc(ctx, n * 4, &y), "feo_malloc");
    must(ctx, feo_malloc(ctx, n * 4, &z), "feo_malloc");
    must(ctx, feo_malloc(ctx, n * 4, &r), "feo_malloc");
    must(ctx, feo_fill_uniform(ctx, FEO_F32, x, n, 1, 0, -1.0, 1.0), "feo_fill_uniform");
    must(ctx, feo_fill_uniform(ctx, FEO_F32, y, n, 2, 0, 1.0, 2.0), "feo_fill_uniform");

    struct Case { const char *name; std::function<int()> call; };
    const float two = 2.0f;
    const Case cases[] = {
        {"ewise add", [&] { return feo_ewise(ctx, FEO_ADD, FEO_F32, z, x, y, n); }},
        {"ewise_scalar mul", [&] { return feo_ewise_scalar(ctx, FEO_MUL, FEO_F32, z, x, &two, n); }},
        {"reduce sum [0,2]", [&] { return feo_reduce(ctx, FEO_SUM, FEO_F32, r, x, 3, shape, 2, axes02); }},
        {"reduce var [0,2]", [&] { return feo_reduce(ctx, FEO_VAR, FEO_F32, r, x, 3, shape, 2, axes02); }},
        {"reduce max [0]", [&] { return feo_reduce(ctx, FEO_MAX, FEO_F32, r, x, 3, shape, 1, axes0); }},
        {"dot 2M", [&] { return feo_dot(ctx, FEO_F32, r, x, y, n); }},
    };
    for (const Case &c : cases) {
        for (int i = 0; i < 20; ++i) must(ctx, c.call(), c.name);
        must(ctx, feo_sync(ctx), "feo_sync");
        const auto t0 = std::chrono::steady_clock::now();
        for (int i = 0; i < iters; ++i) must(ctx, c.call(), c.name);
        const auto t1 = std::chrono::steady_clock::now();
        must(ctx, feo_sync(ctx), "feo_sync");
        const auto t2 = std::chrono::steady_clock::now();
        const double issue = std::chrono::duration<double, std::micro>(t1 - t0).count() / iters;
        const double total = std::chrono::duration<double, std::micro>(t2 - t0).count() / iters;
        std::printf("%-18s issue %6.2f us/call   with the stream drained %6.2f us/call\n", c.name, issue, total);
    }
    must(ctx, feo_free(ctx, x), "feo_free");
    must(ctx, feo_free(ctx, y), "feo_free");
    must(ctx, feo_free(ctx, z), "feo_free");
    must(ctx, feo_free(ctx, r), "feo_free");
    must(ctx, feo_shutdown(ctx), "feo_shutdown");
    return 0;
}
