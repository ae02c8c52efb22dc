"""2 ranks: where does the striped frame differ from this rank's own single-GPU frame?"""
import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch, torch.distributed as dist
from gaussiancity_b200 import ext, sharding
from gaussiancity_b200.synthetic import uniform_scene
from tests import refext
rank, world = int(os.environ["RANK"]), int(os.environ["WORLD_SIZE"])
dev = torch.device(f"cuda:{int(os.environ['LOCAL_RANK'])}"); torch.cuda.set_device(dev)
dist.init_process_group("nccl", device_id=dev)
P, W, H = 200_000, 640, 360
s = uniform_scene(P, W, H, sh_degree=3, seed=17, device=dev, bg=(0.1, 0.2, 0.3))
e = torch.Tensor([])
chk = torch.stack([s.means3D.double().sum(), s.shs.double().sum(), s.scales.double().sum()])
allchk = [torch.zeros_like(chk) for _ in range(world)]; dist.all_gather(allchk, chk)
print(f"rank {rank}: scene checksums equal across ranks: {all(torch.equal(allchk[0], c) for c in allchk)}", flush=True)
inp = dict(bg=s.bg, means3D=s.means3D, opacity=s.opacities, scales=s.scales, rotations=s.rotations, sh=s.shs, colors=e,
           view=s.view_matrix, proj=s.proj_matrix, campos=s.campos)
cam = dict(bg=inp["bg"], view=inp["view"], proj=inp["proj"], campos=inp["campos"], tanfovx=s.tanfovx, tanfovy=s.tanfovy,
           img_h=H, img_w=W, sh_degree=3)
R1, col1, radii1, g1, b1, i1 = ext.rasterize_gaussians(*refext.scene_forward_args(s))
R1b, col1b, *_ = ext.rasterize_gaussians(*refext.scene_forward_args(s))
print(f"rank {rank}: single-GPU render deterministic: {torch.equal(col1, col1b)} R {R1} {R1b}", flush=True)
allc = [torch.zeros_like(col1) for _ in range(world)]; dist.all_gather(allc, col1)
print(f"rank {rank}: single-GPU frames equal across ranks: {all(torch.equal(allc[0], c) for c in allc)}", flush=True)
for balanced in (True, False):
    eng = sharding.TileShardedRasterizer(device=dev, exchange="collective", balanced=balanced)
    for frame in range(2):
        color, radii, state = eng.render(inp, cam, assemble=False)
        bounds = eng.backend.stripe_bounds(state, inp, world).cpu().tolist()
        own = color.clone()
        full = eng.assemble_image(color.clone())
        torch.cuda.synchronize()
        r0, r1 = bounds[rank] * 16, min(H, bounds[rank + 1] * 16)
        bad_rows = torch.nonzero((full != col1).any(dim=0).any(dim=1)).flatten().tolist()
        own_bad = torch.nonzero((own[:, r0:r1] != col1[:, r0:r1]).any(dim=0).any(dim=1)).flatten().tolist()
        outside = bool((own[:, :r0] != 0).any()) or bool((own[:, r1:] != 0).any())
        print(f"rank {rank} balanced={balanced} f{frame}: bounds {bounds} R_local {state['R']} full==single {torch.equal(full, col1)} "
              f"bad rows (full) {bad_rows[:8]}..{len(bad_rows)} own-stripe bad rows {own_bad[:8]}..{len(own_bad)} nonzero outside stripe {outside} "
              f"radii ok {torch.equal(radii, radii1)}", flush=True)
dist.destroy_process_group()
