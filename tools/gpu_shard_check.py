"""Multi-GPU check (run under torchrun on the B200 box): the NCCL tile-sharded frame equals the
single-GPU frame bit for bit and the reduce-scattered gradients match the monolithic backward."""
import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch, torch.distributed as dist
import bench
from gaussiancity_b200 import ext, sharding

rank, world, local = int(os.environ["RANK"]), int(os.environ["WORLD_SIZE"]), int(os.environ["LOCAL_RANK"])
dev = torch.device(f"cuda:{local}"); torch.cuda.set_device(dev)
dist.init_process_group("nccl", device_id=dev)
s = bench.make_scene(sys.argv[1] if len(sys.argv) > 1 else "cfg3_1M_sh3_1080p", dev)
inp = bench.scene_inputs(s)
if rank != 0:   # only rank 0 owns the data before the broadcast
    for k in ("means3D", "opacity", "scales", "rotations", "sh"):
        inp[k] = torch.zeros_like(inp[k])
G = torch.randn(3, s.img_h, s.img_w, generator=torch.Generator().manual_seed(3)).to(dev)
eng = sharding.TileShardedRasterizer(device=dev)
cam = eng._cam(s, inp)
color, radii, state = eng.render(inp, cam, src=0)
grads, (st, cnt) = eng.backward(state, inp, cam, G, gather=True)
torch.cuda.synchronize()
# single-GPU reference on every rank (inputs are identical after the broadcast)
R, c1, r1, geom, binning, img = ext.rasterize_gaussians(*bench.fwd_args(s, inp))
g1 = ext.rasterize_gaussians_backward(*bench.bwd_args(s, inp, r1, G, geom, R, binning, img))
ok = torch.equal(color, c1) and torch.equal(radii, r1)
errs = []
for a, b in zip(grads, g1):
    if b.numel():
        errs.append(((a.double() - b.double()).norm() / b.double().norm().clamp_min(1e-30)).item())
Rl = torch.tensor([state["R"]], device=dev); dist.all_reduce(Rl)
print(f"rank {rank}/{world}: frame bit-identical={ok}  R_local={state['R']} sum={int(Rl)} single={R} "
      f"slice=({st},{cnt}) grad rel errs max={max(errs):.2e}", flush=True)
assert ok and int(Rl) == R and max(errs) < 1e-5
dist.destroy_process_group()
