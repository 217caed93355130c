import ctypes as C, os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
import feotensor_b200 as feo, feotensor_b200._lib as L
torch.cuda.set_stream(torch.cuda.Stream())
ctx = feo.Context(0); ctx.set_stream(torch.cuda.current_stream().cuda_stream)
lib, h, vp = ctx.lib, ctx.handle, C.c_void_p
keep_off = int(sys.argv[1])
ctx.tune(L.KNOB_BCAST_L2KEEP, keep_off)
D, R = 4096, 16
def dev(n): return torch.empty(n * 4, dtype=torch.uint8, device="cuda")
a, b, o = dev(R * D), dev(D * D), dev(R * D * D)
for t, n, s in ((a, R * D, 2), (b, D * D, 3)): ctx.check(lib.feo_fill_uniform(h, L.F32, vp(t.data_ptr()), n, s, 0, -1.0, 1.0))
def timed(fn, iters=5):
    fn(); torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(iters): fn()
    e1.record(); torch.cuda.synchronize()
    return e0.elapsed_time(e1) / iters
shape = [16384, 16384, 8]; n3 = 16384 * 16384 * 8
x, out = dev(n3), dev(16384 * 16384)
ctx.check(lib.feo_fill_uniform(h, L.F32, vp(x.data_ptr()), n3, 30, 0, -1.0, 1.0))
red = lambda: ctx.check(lib.feo_reduce(h, L.SUM, L.F32, vp(out.data_ptr()), vp(x.data_ptr()), 3, L.sizes(shape), 1, L.sizes([2])))
print("keep_off", keep_off, "reduce [2] before any broadcast: %.4f ms" % timed(red))
bc = lambda: ctx.check(lib.feo_ewise_broadcast(h, L.MUL, L.F32, vp(o.data_ptr()), vp(a.data_ptr()), vp(b.data_ptr()), 3, L.sizes([R, D, D]), L.sizes([D, 0, 1]), L.sizes([0, D, 1])))
print("broadcast slab: %.4f ms" % timed(bc, 20))
print("reduce [2] after broadcasts: %.4f ms" % timed(red))
red0 = lambda: ctx.check(lib.feo_reduce(h, L.SUM, L.F32, vp(out.data_ptr()), vp(x.data_ptr()), 3, L.sizes(shape), 1, L.sizes([0])))
print("reduce [0] after broadcasts: %.4f ms" % timed(red0))
print("reduce [2] again: %.4f ms" % timed(red))
