"""Large-scene sanity (GPU box): 20 M Gaussians / 1080p, ours vs the reference extension --
num_rendered, radii and the image must match bit for bit; reports both timings."""
import os, sys, time
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
from gaussiancity_b200 import ext
from gaussiancity_b200.synthetic import uniform_scene
from tests import refext

dev = torch.device("cuda:0")
P = int(sys.argv[1]) if len(sys.argv) > 1 else 20_000_000
s = uniform_scene(P, 1920, 1080, sh_degree=0, seed=3, device=dev)
ref = refext.load_reference_ext()
G = torch.randn(3, 1080, 1920, device=dev)
def run(mod):
    torch.cuda.synchronize(); t = time.perf_counter()
    R, color, radii, geom, binning, img = mod.rasterize_gaussians(*refext.scene_forward_args(s))
    grads = mod.rasterize_gaussians_backward(*refext.scene_backward_args(s, radii, G, geom, R, binning, img))
    torch.cuda.synchronize(); return R, color, radii, grads, (time.perf_counter() - t) * 1e3
for _ in range(2): ro = run(ext)
print(f"ours: R={ro[0]} {ro[4]:.2f} ms  peak mem {torch.cuda.max_memory_allocated()/1e9:.1f} GB")
if ref is not None:
    for _ in range(2): rr = run(ref)
    print(f"ref : R={rr[0]} {rr[4]:.2f} ms")
    print("R equal", ro[0] == rr[0], "radii equal", torch.equal(ro[2], rr[2]), "image equal", torch.equal(ro[1], rr[1]))
    for n, a, b in zip(["m2d", "col", "op", "m3d", "cov", "sh", "sc", "rot"], ro[3], rr[3]):
        if b.numel(): print(n, f"{((a.double()-b.double()).norm()/b.double().norm().clamp_min(1e-30)).item():.2e}", end="  ")
    print()
