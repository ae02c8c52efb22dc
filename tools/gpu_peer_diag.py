"""2 ranks under torchrun: where does the time of the peer-exchange backward go?
Times blend (peer accumulators) / barrier / owners' geometry on (a) the IPC-exported accumulator
with clearing, (b) the same without clearing, (c) a torch-allocated copy of it."""
import ctypes, os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch, torch.distributed as dist
import bench
from gaussiancity_b200 import _cabi, ext, sharding

rank, world = int(os.environ["RANK"]), int(os.environ["WORLD_SIZE"])
dev = torch.device(f"cuda:{int(os.environ['LOCAL_RANK'])}"); torch.cuda.set_device(dev)
dist.init_process_group("nccl", device_id=dev)
s = bench.make_scene("cfg4_5M_sh3_1080p", dev); inp = bench.scene_inputs(s)
P, H, W = s.means3D.shape[0], s.img_h, s.img_w
G = torch.randn(3, H, W, device=dev)
eng = sharding.TileShardedRasterizer(device=dev, exchange="peer")
cam = eng._cam(s, inp); be = eng.backend
color, radii, state = eng.render(inp, cam, assemble=False)
pr = be.peer_setup(P, rank, world, None)
out = None
def ev(): return torch.cuda.Event(enable_timing=True)
def timed(fn, n=8):
    fn(); torch.cuda.synchronize(); dist.barrier()
    a, b = ev(), ev(); a.record()
    for _ in range(n): fn()
    b.record(); torch.cuda.synchronize(); return a.elapsed_time(b) / n
accs = [b + 256 for b in pr["bases"]]
local_t = torch.zeros(P, 12, device=dev)
def blend(): ext.rasterize_gaussians_backward_blend(cam["bg"], P, state["R"], G, state["geom"], state["binning"], state["img"], shard_rank=rank, shard_count=world, accumulators=accs)
def blend_local(): ext.rasterize_gaussians_backward_blend(cam["bg"], P, state["R"], G, state["geom"], state["binning"], state["img"], shard_rank=rank, shard_count=world, accumulators=[local_t.data_ptr()])
def barrier():
    pr["epoch"] += 1
    flags = (ctypes.c_void_p * world)(*pr["bases"])
    _cabi.check(_cabi.lib().gcr_peer_barrier(flags, rank, world, pr["epoch"], ctypes.c_void_p(torch.cuda.current_stream(dev).cuda_stream)), "barrier")
def geom(acc, clear):
    global out
    out = be.backward_geometry(state, inp, cam, acc, rank, world, out=out, clear=clear)
res = {}
res["blend_peer"] = timed(blend)
res["blend_local"] = timed(blend_local)
res["barrier"] = timed(barrier)
res["geom_ipc_clear"] = timed(lambda: geom(accs[rank], True))
res["geom_ipc_noclear"] = timed(lambda: geom(accs[rank], False))
res["geom_torch_noclear"] = timed(lambda: geom(local_t, False))
res["geom_torch_clear"] = timed(lambda: geom(local_t, True))
def full():
    blend(); barrier(); geom(accs[rank], True)
res["blend+barrier+geom"] = timed(full)
res["forward"] = timed(lambda: eng.render(inp, cam, assemble=False))
print(f"rank {rank}: " + ", ".join(f"{k}={v:.3f}" for k, v in res.items()), flush=True)
eng.close(); dist.destroy_process_group()
