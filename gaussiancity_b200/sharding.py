"""Screen-tile sharding of ONE frame across the GPUs of a node (north_star).

The hot path shards by tile rows: tile row r belongs to rank r % world (interleaved, so the
load balances without knowing the scene).  One exchange step each way, both through
torch.distributed (NCCL over NVLink 5 / NVSwitch on the B200 box; gloo in the CPU tests):

  forward : [broadcast of the Gaussian buffers from the owning rank, once per frame]
            -> every rank preprocesses and depth-sorts ALL Gaussians (cheap, replicated), but
               bins and blends only its own tile rows
            -> image rows are disjoint across ranks: all_reduce(SUM) assembles the frame
               (x + 0 is exact, so the frame is bit-identical to the single-GPU render)
  backward: every rank blends the gradients of its own tiles into a PARTIAL [P,12]
            per-Gaussian accumulator -> reduce_scatter over P -> each rank finishes the
            geometry backward for its own slice of Gaussians (optionally all_gather'ed).

The compute is delegated to a backend object so the collective plumbing can be exercised on
CPU (tests/ plug the CPU oracle in); the product backend is the CUDA library (CudaBackend).
"""
from typing import Optional

import torch
import torch.distributed as dist


def owned_tile_rows(grid_y: int, rank: int, world: int):
    return list(range(rank, grid_y, world))


def gaussian_slice(P: int, rank: int, world: int):
    """Contiguous slice of Gaussians whose geometry backward `rank` finishes, for P padded up to
    a multiple of world: (start, count, padded_P)."""
    padded = (P + world - 1) // world * world
    per = padded // world
    start = min(rank * per, P)
    count = max(0, min(P, (rank + 1) * per) - start)
    return start, count, padded


class CudaBackend:
    """The sm_100a library through gaussiancity_b200.ext (no CPU path)."""

    def forward(self, inp, cam, rank, world):
        from . import ext
        e = torch.Tensor([])
        R, color, radii, geom, binning, img = ext.rasterize_gaussians(
            cam["bg"], inp["means3D"], inp.get("colors", e), inp["opacity"], inp["scales"],
            inp["rotations"], 1.0, e, cam["view"], cam["proj"], cam["tanfovx"], cam["tanfovy"],
            cam["img_h"], cam["img_w"], inp.get("sh", e), cam["sh_degree"], cam["campos"], False,
            False, shard_rank=rank, shard_count=world)
        return color, radii, dict(R=R, geom=geom, binning=binning, img=img, radii=radii)

    def backward_blend(self, state, inp, cam, grad_out, rank, world):
        from . import ext
        P = inp["means3D"].shape[0]
        return ext.rasterize_gaussians_backward_blend(cam["bg"], P, state["R"], grad_out,
                                                      state["binning"], state["img"],
                                                      shard_rank=rank, shard_count=world)

    def backward_geometry(self, state, inp, cam, grad_acc, start, count):
        from . import ext
        e = torch.Tensor([])
        return ext.rasterize_gaussians_backward_geometry(
            inp["means3D"], state["radii"], inp["scales"], inp["rotations"], 1.0, e, cam["view"],
            cam["proj"], cam["tanfovx"], cam["tanfovy"], cam["img_h"], cam["img_w"],
            inp.get("sh", e), cam["sh_degree"], cam["campos"], state["geom"], grad_acc,
            range_start=start, range_count=count)


GRAD_NAMES = ("dL_dmeans2D", "dL_dcolors", "dL_dopacity", "dL_dmeans3D", "dL_dcov3D", "dL_dsh",
              "dL_dscales", "dL_drotations")


class TileShardedRasterizer:
    def __init__(self, device=None, group: Optional[dist.ProcessGroup] = None, backend=None):
        self.group = group
        self.rank = dist.get_rank(group) if dist.is_initialized() else 0
        self.world = dist.get_world_size(group) if dist.is_initialized() else 1
        self.device = device
        self.backend = backend if backend is not None else CudaBackend()
        self.last_num_rendered_total = None

    # -- collectives -------------------------------------------------------------------------
    def broadcast_gaussians(self, inp, src=0):
        """Broadcast the per-frame Gaussian buffers from rank `src` (in place)."""
        if self.world == 1:
            return inp
        for k in sorted(inp):
            t = inp[k]
            if isinstance(t, torch.Tensor) and t.numel() > 0:
                dist.broadcast(t, src=src, group=self.group)
        return inp

    def assemble_image(self, color_partial):
        if self.world > 1:
            dist.all_reduce(color_partial, op=dist.ReduceOp.SUM, group=self.group)
        return color_partial

    def reduce_scatter_grad_acc(self, grad_acc):
        """grad_acc [P,12] partial -> (reduced_full_view [padded,12] where only this rank's slice
        is valid, start, count)."""
        P = grad_acc.shape[0]
        start, count, padded = gaussian_slice(P, self.rank, self.world)
        if self.world == 1:
            return grad_acc, 0, P
        if padded != P:
            pad = torch.zeros((padded - P, grad_acc.shape[1]), dtype=grad_acc.dtype, device=grad_acc.device)
            grad_acc = torch.cat([grad_acc, pad], dim=0)
        per = padded // self.world
        reduced = torch.empty_like(grad_acc)
        out_slice = reduced[self.rank * per:(self.rank + 1) * per]
        dist.reduce_scatter_tensor(out_slice, grad_acc, op=dist.ReduceOp.SUM, group=self.group)
        return reduced, start, count

    # -- one frame ---------------------------------------------------------------------------
    def render(self, inp, cam, src=0, broadcast=True, assemble=True):
        if broadcast:
            self.broadcast_gaussians(inp, src)
        color, radii, state = self.backend.forward(inp, cam, self.rank, self.world)
        if assemble:
            color = self.assemble_image(color)
        return color, radii, state

    def backward(self, state, inp, cam, grad_out, gather=False):
        acc = self.backend.backward_blend(state, inp, cam, grad_out, self.rank, self.world)
        reduced, start, count = self.reduce_scatter_grad_acc(acc)
        grads = self.backend.backward_geometry(state, inp, cam, reduced, start, count)
        if gather and self.world > 1:
            grads = tuple(self._gather_rows(g, start, count) for g in grads)
        return grads, (start, count)

    def _gather_rows(self, g, start, count):
        P = g.shape[0]
        _, _, padded = gaussian_slice(P, self.rank, self.world)
        per = padded // self.world
        flat = g.reshape(P, -1)
        width = flat.shape[1]
        if width == 0:
            return g
        buf = torch.zeros((padded, width), dtype=g.dtype, device=g.device)
        buf[start:start + count] = flat[start:start + count]
        out = torch.empty_like(buf)
        dist.all_gather_into_tensor(out, buf[self.rank * per:(self.rank + 1) * per].contiguous(),
                                    group=self.group)
        return out[:P].reshape(g.shape)

    # -- bench.py helpers -------------------------------------------------------------------------
    @staticmethod
    def _cam(s, inp):
        return dict(bg=inp["bg"], view=inp["view"], proj=inp["proj"], campos=inp["campos"],
                    tanfovx=s.tanfovx, tanfovy=s.tanfovy, img_h=s.img_h, img_w=s.img_w,
                    sh_degree=s.sh_degree)

    def forward(self, s, inp, src=0):
        color, radii, state = self.render(inp, self._cam(s, inp), src=src)
        self._note_R(state)
        return color

    def forward_backward(self, s, inp, grad_out, src=0):
        cam = self._cam(s, inp)
        color, radii, state = self.render(inp, cam, src=src)
        grads, sl = self.backward(state, inp, cam, grad_out)
        self._note_R(state)
        return color, grads, sl

    # -- frame pipelining: the broadcast of frame i+1 overlaps the compute of frame i --------
    def start_prefetch(self, inp_next, src=0):
        """Issue the per-frame broadcast of `inp_next` on a side stream (NCCL's own stream is
        ordered after it, not after the compute stream). Returns a handle for wait_prefetch().
        `inp_next` must not be read or written by the compute stream until then."""
        if self.world == 1:
            return None
        if not hasattr(self, "_comm_stream"):
            self._comm_stream = torch.cuda.Stream(device=self.device)
        cur = torch.cuda.current_stream(self.device)
        self._comm_stream.wait_stream(cur)   # the buffers' previous readers (frame i-1) are done
        works = []
        with torch.cuda.stream(self._comm_stream):
            for k in sorted(inp_next):
                t = inp_next[k]
                if isinstance(t, torch.Tensor) and t.numel() > 0:
                    works.append(dist.broadcast(t, src=src, group=self.group, async_op=True))
        return works

    def wait_prefetch(self, handle):
        if handle:
            for w in handle:
                w.wait()   # makes the current (compute) stream wait for the broadcast

    def forward_backward_prefetched(self, s, inp, grad_out):
        """forward_backward on buffers whose broadcast was started with start_prefetch()."""
        cam = self._cam(s, inp)
        color, radii, state = self.render(inp, cam, broadcast=False)
        grads, sl = self.backward(state, inp, cam, grad_out)
        self._note_R(state)
        return color, grads, sl

    def _note_R(self, state):
        self.last_num_rendered_local = int(state["R"])
        self.last_num_rendered_total = self.last_num_rendered_local
