import ctypes as C, os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
import feotensor_b200 as feo, feotensor_b200._lib as L
torch.cuda.set_stream(torch.cuda.Stream())
ctx = feo.Context(0); ctx.set_stream(torch.cuda.current_stream().cuda_stream)
lib, h, vp = ctx.lib, ctx.handle, C.c_void_p
def timed(fn, iters=5):
    fn(); torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(iters): fn()
    e1.record(); torch.cuda.synchronize()
    return e0.elapsed_time(e1) / iters
shape = [16384, 16384, 8]; n3 = 16384 * 16384 * 8
GiB = 1 << 30
big = torch.empty(12 * GiB, dtype=torch.uint8, device="cuda")
base = big.data_ptr()
print("base %% 2MiB = %d" % (base % (2 << 20)))
ctx.check(lib.feo_fill_uniform(h, L.F32, vp(base), n3, 30, 0, -1.0, 1.0))
for delta in (0, 256, 4096, 65536, 1 << 20, 2 << 20, 3 << 20, 16 << 20, 100 << 20, 512 << 20, (1 << 30) + (1 << 20), 2 << 30):
    outp = base + 8 * GiB + delta
    red = lambda: ctx.check(lib.feo_reduce(h, L.SUM, L.F32, vp(outp), vp(base), 3, L.sizes(shape), 1, L.sizes([2])))
    print("out at 8GiB + %11d: %.4f ms" % (delta, timed(red)))
# input offset sweep with out fixed
for delta in (0, 1 << 20, 2 << 20, 64 << 20, 1 << 30):
    inp = base + delta
    ctx.check(lib.feo_fill_uniform(h, L.F32, vp(inp), n3, 30, 0, -1.0, 1.0))
    outp = base + 10 * GiB
    red = lambda: ctx.check(lib.feo_reduce(h, L.SUM, L.F32, vp(outp), vp(inp), 3, L.sizes(shape), 1, L.sizes([2])))
    print("in at +%11d, out at 10GiB: %.4f ms" % (delta, timed(red)))
