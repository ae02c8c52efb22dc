"""GPU box: where does host time go? step time with/without the stage profiler, fwd/bwd split."""
import os, sys, time
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
import bench
from gaussiancity_b200 import _cabi, ext

dev = torch.device("cuda:0")
wl = sys.argv[1] if len(sys.argv) > 1 else "cfg3_1M_sh3_1080p"
s = bench.make_scene(wl, dev)
inp = bench.scene_inputs(s)
grad_out = torch.randn(3, s.img_h, s.img_w, device=dev)

def step():
    R, color, radii, geom, binning, img = ext.rasterize_gaussians(*bench.fwd_args(s, inp))
    return ext.rasterize_gaussians_backward(*bench.bwd_args(s, inp, radii, grad_out, geom, R, binning, img))

def timeit(fn, n=10, warm=3):
    for _ in range(warm): fn()
    torch.cuda.synchronize(); t = time.perf_counter()
    for _ in range(n): fn()
    torch.cuda.synchronize(); return (time.perf_counter() - t) / n * 1e3

print("step (profiler off): %.3f ms" % timeit(step))
_cabi.profile_enable(True)
print("step (profiler on):  %.3f ms" % timeit(step))
st = _cabi.profile_read(); print("stages:", {k: round(v, 3) for k, v in st.items()}, "sum %.3f" % sum(st.values()))
_cabi.profile_enable(False)
print("step (profiler off): %.3f ms" % timeit(step))
def fwd(): return ext.rasterize_gaussians(*bench.fwd_args(s, inp))
print("fwd only: %.3f ms" % timeit(fwd))
# host time of the forward call itself (returns after launching blend)
torch.cuda.synchronize(); t = time.perf_counter(); out = fwd(); t1 = time.perf_counter(); torch.cuda.synchronize(); t2 = time.perf_counter()
print("fwd host return %.3f ms, +drain %.3f ms" % ((t1 - t) * 1e3, (t2 - t1) * 1e3))
R, color, radii, geom, binning, img = out
torch.cuda.synchronize(); t = time.perf_counter(); g = ext.rasterize_gaussians_backward(*bench.bwd_args(s, inp, radii, grad_out, geom, R, binning, img)); t1 = time.perf_counter(); torch.cuda.synchronize(); t2 = time.perf_counter()
print("bwd host return %.3f ms, +drain %.3f ms" % ((t1 - t) * 1e3, (t2 - t1) * 1e3))
