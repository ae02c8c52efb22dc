import os, sys, time
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
import bench
from gaussiancity_b200 import _cabi, ext
dev = torch.device("cuda:0")
wl = sys.argv[1]
s = bench.make_scene(wl, dev); inp = bench.scene_inputs(s)
grad_out = torch.randn(3, s.img_h, s.img_w, device=dev)
def step():
    t0 = time.perf_counter()
    R, color, radii, geom, binning, img = ext.rasterize_gaussians(*bench.fwd_args(s, inp))
    t1 = time.perf_counter()
    g = ext.rasterize_gaussians_backward(*bench.bwd_args(s, inp, radii, grad_out, geom, R, binning, img))
    t2 = time.perf_counter()
    return (t1 - t0) * 1e3, (t2 - t1) * 1e3
for i in range(8):
    torch.cuda.synchronize() if i == 0 else None
    t = time.perf_counter(); a, b = step(); print("iter %d host fwd %.3f bwd %.3f  total %.3f  mem %.2f GB reserved %.2f GB" % (i, a, b, (time.perf_counter() - t) * 1e3, torch.cuda.memory_allocated() / 1e9, torch.cuda.memory_reserved() / 1e9), file=sys.stderr)
torch.cuda.synchronize()
print(torch.cuda.memory_summary(abbreviated=True)[:1500], file=sys.stderr)
