"""Screen-tile sharding of ONE frame across the GPUs of a node (north_star).

The frame's tile rows are cut into `world` CONTIGUOUS stripes, balanced by a prefix sum of the
per-row tile-instance counts (a geometry-only pre-pass every rank runs on the same inputs, so
all ranks obtain the same bounds without talking).  Nothing that is per-Gaussian heavy is
replicated: each rank

  forward : projects all Gaussians (44 B each; needed to find the ones that reach its stripe --
            the same pass counts tile instances per tile row and cuts the balanced stripes), but
            evaluates SH colours, depth-sorts, bins and blends only the ~P/world Gaussians whose
            tile rect touches its stripe;
  backward: blends the gradients of its own tiles.  The per-Gaussian sums of a Gaussian that
            straddles a stripe boundary have to meet somewhere: every Gaussian is OWNED by the
            rank whose stripe holds its centre row, and the blend kernel adds each record's sums
            straight into the owner's accumulator -- local memory for the ~85-90 % of Gaussians
            that live inside one stripe, a peer-mapped buffer over NVLink 5 / NVSwitch for the
            boundary ones (`exchange="peer"`).  One cross-GPU barrier kernel later every owner
            finishes the geometry backward for its own Gaussians.  No dense [P,12]
            reduce-scatter (240 MB at 5 M Gaussians), no collective on the data path at all.

`exchange="collective"` is the torch.distributed formulation of the same exchange (partial
[P,12] accumulators -> all_reduce -> owners finish): it runs on any backend -- the CPU tests plug
the oracle in and use gloo -- and is the fallback when peer mapping is unavailable.

Outputs stay sharded, as SURVEY 8e allows: the image stripe of each rank (optionally assembled
into a full frame on every rank), and dense [P,...] gradient tensors in which each rank has
written exactly the rows of the Gaussians it owns (`owner_mask`).

The compute is delegated to a backend object so the plumbing can be exercised on CPU; the
product backend is the CUDA library (CudaBackend).
"""
import ctypes
from typing import Optional

import torch
import torch.distributed as dist

GRAD_NAMES = ("dL_dmeans2D", "dL_dcolors", "dL_dopacity", "dL_dmeans3D", "dL_dcov3D", "dL_dsh",
              "dL_dscales", "dL_drotations")


def equal_stripes(grid_y: int, world: int):
    """Equal-height contiguous stripes (what the library uses when no bounds are given)."""
    return [grid_y * k // world for k in range(world + 1)]


def balanced_stripes(row_counts, world: int):
    """Host restatement of the device partition (csrc/preprocess.cu stripe_partition_kernel):
    bounds[k] = the row boundary whose instance prefix is nearest to k/world of the total."""
    grid_y = len(row_counts)
    total = int(sum(int(c) for c in row_counts))
    if total == 0:
        return equal_stripes(grid_y, world)
    bounds, prefix, r = [0], 0, 0
    for k in range(1, world):
        target = total * k // world
        while r < grid_y and prefix + int(row_counts[r]) <= target:
            prefix += int(row_counts[r])
            r += 1
        if r < grid_y and (prefix + int(row_counts[r]) - target) < (target - prefix):
            prefix += int(row_counts[r])
            r += 1
        bounds.append(r)
    bounds.append(grid_y)
    return bounds


class CudaBackend:
    """The sm_100a library through gaussiancity_b200.ext (no CPU path)."""

    def __init__(self, device=None):
        self.device = device
        self._peer = None        # peer-mapped accumulators (exchange="peer")
        self._part_ws = None
        self._bounds = None

    # -- stripes ---------------------------------------------------------------------------------
    def partition(self, inp, cam, world):
        from . import ext
        dev = inp["means3D"].device
        grid_y = (cam["img_h"] + 15) // 16
        if self._part_ws is None or self._part_ws.numel() < grid_y + 1:
            self._part_ws = torch.empty(grid_y + 1, dtype=torch.int32, device=dev)
        if self._bounds is None or self._bounds.numel() != world + 1:
            self._bounds = torch.empty(world + 1, dtype=torch.int32, device=dev)
        ext.stripe_partition(inp["means3D"], inp["scales"], inp["rotations"], 1.0, cam["view"], cam["proj"],
                             cam["tanfovx"], cam["tanfovy"], cam["img_h"], cam["img_w"], world,
                             self._part_ws, self._bounds)
        return self._bounds

    def forward(self, inp, cam, rank, world, bounds, balanced=False):
        from . import ext
        e = torch.Tensor([])
        R, color, radii, geom, binning, img = ext.rasterize_gaussians(
            cam["bg"], inp["means3D"], inp.get("colors", e), inp["opacity"], inp["scales"],
            inp["rotations"], 1.0, e, cam["view"], cam["proj"], cam["tanfovx"], cam["tanfovy"],
            cam["img_h"], cam["img_w"], inp.get("sh", e), cam["sh_degree"], cam["campos"], False,
            False, shard_rank=rank, shard_count=world, stripe_bounds=bounds, balanced=balanced)
        return color, radii, dict(R=R, geom=geom, binning=binning, img=img, radii=radii)

    # -- exchange="collective": partial accumulators, reduced by torch.distributed ----------------
    def backward_blend(self, state, inp, cam, grad_out, rank, world):
        from . import ext
        P = inp["means3D"].shape[0]
        return ext.rasterize_gaussians_backward_blend(
            cam["bg"], P, state["R"], grad_out, state["geom"], state["binning"], state["img"],
            shard_rank=rank, shard_count=world)

    def backward_geometry(self, state, inp, cam, grad_acc, rank, world, out=None, clear=False):
        """out: None, a tuple of 8 dense tensors, or -- striped frames -- `(rows [P,24], dL_dsh)`: the
        packed form (one 96-byte row per owned Gaussian instead of seven scattered partial-sector
        writes); the returned gradients are then views into `rows`."""
        from . import ext
        e = torch.Tensor([])
        packed = None
        if world > 1:
            packed = out if (out is not None and len(out) == 2) else True
            out = None
        return ext.rasterize_gaussians_backward_geometry(
            inp["means3D"], state["radii"], inp["scales"], inp["rotations"], 1.0, e, cam["view"],
            cam["proj"], cam["tanfovx"], cam["tanfovy"], cam["img_h"], cam["img_w"],
            inp.get("sh", e), cam["sh_degree"], cam["campos"], state["geom"], grad_acc,
            shard_rank=rank, striped=world > 1, clear_accumulator=clear, out=out, packed=packed)

    def owner_mask(self, state, inp, rank):
        from . import ext
        return ext.owner_bytes(state["geom"], inp["means3D"].shape[0]) == rank

    def stripe_bounds(self, state, inp, world):
        from . import ext
        return ext.stripe_bounds_of(state["geom"], inp["means3D"].shape[0], world)

    # -- exchange="peer": accumulators other ranks add into over NVLink ---------------------------
    def peer_setup(self, P, rank, world, group):
        """One zero-filled device block per rank: [flags 256 B | accumulator 0 | accumulator 1],
        its IPC handle published to the other ranks, theirs mapped here.  Two accumulators
        alternate by frame so that a fast rank's next frame can never add into an accumulator
        its owner has not finished reading (the single barrier per frame orders the rest)."""
        from . import _cabi
        if self._peer is not None and self._peer["P"] >= P and self._peer["world"] == world:
            return self._peer
        self.peer_teardown()
        lib = _cabi.lib()
        acc_bytes = (48 * P + 255) // 256 * 256
        total = 256 + 2 * acc_bytes
        ptr = ctypes.c_void_p()
        handle = (ctypes.c_ubyte * _cabi.PEER_HANDLE_BYTES)()
        _cabi.check(lib.gcr_peer_alloc(total, ctypes.byref(ptr), handle), "gcr_peer_alloc")
        handles = [None] * world
        dist.all_gather_object(handles, bytes(handle), group=group)
        bases = []
        for k in range(world):
            if k == rank:
                bases.append(ptr.value)
                continue
            q = ctypes.c_void_p()
            hk = (ctypes.c_ubyte * _cabi.PEER_HANDLE_BYTES).from_buffer_copy(handles[k])
            _cabi.check(lib.gcr_peer_open(hk, ctypes.byref(q)), "gcr_peer_open")
            bases.append(q.value)
        dist.barrier(group=group)   # every rank has mapped every block before anyone writes
        self._peer = dict(P=P, world=world, rank=rank, bases=bases, acc_bytes=acc_bytes, epoch=0, frame=0)
        return self._peer

    def peer_teardown(self):
        if self._peer is None:
            return
        from . import _cabi
        lib = _cabi.lib()
        torch.cuda.synchronize()
        pr = self._peer
        for k, b in enumerate(pr["bases"]):
            if k == pr["rank"]:
                lib.gcr_peer_free(ctypes.c_void_p(b))
            else:
                lib.gcr_peer_close(ctypes.c_void_p(b))
        self._peer = None

    def peer_backward(self, state, inp, cam, grad_out, rank, world, out=None, remote_scalar=False):
        """blend (sums go to the owners over peer memory) -> barrier -> owners' geometry."""
        from . import _cabi, ext
        pr = self._peer
        P = inp["means3D"].shape[0]
        parity = pr["frame"] & 1
        accs = [b + 256 + parity * pr["acc_bytes"] for b in pr["bases"]]
        ext.rasterize_gaussians_backward_blend(
            cam["bg"], P, state["R"], grad_out, state["geom"], state["binning"], state["img"],
            shard_rank=rank, shard_count=world, accumulators=accs, remote_scalar=remote_scalar)
        pr["epoch"] += 1
        flags = (ctypes.c_void_p * world)(*pr["bases"])
        dev = inp["means3D"].device
        with torch.cuda.device(dev):
            _cabi.check(_cabi.lib().gcr_peer_barrier(flags, rank, world, pr["epoch"],
                                                     ctypes.c_void_p(torch.cuda.current_stream(dev).cuda_stream)),
                        "gcr_peer_barrier")
        grads = self.backward_geometry(state, inp, cam, accs[rank], rank, world, out=out, clear=True)
        pr["frame"] += 1
        return grads


class TileShardedRasterizer:
    """One frame over the ranks of `group`.  exchange: "peer" (CUDA backend on NVLink-connected
    GPUs) or "collective" (any backend / fabric)."""

    def __init__(self, device=None, group: Optional[dist.ProcessGroup] = None, backend=None,
                 exchange: str = "collective", balanced: bool = True, remote_scalar: bool = False):
        self.group = group
        self.rank = dist.get_rank(group) if dist.is_initialized() else 0
        self.world = dist.get_world_size(group) if dist.is_initialized() else 1
        self.device = device
        self.backend = backend if backend is not None else CudaBackend(device)
        if exchange not in ("peer", "collective"):
            raise ValueError("exchange must be 'peer' or 'collective'")
        self.exchange = exchange if self.world > 1 else "collective"
        self.balanced = balanced
        self.remote_scalar = remote_scalar
        self.last_num_rendered_local = None
        self._peer_checked = False
        self.peer_error = None

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
        """Stripes are disjoint and every rank's frame is zero outside its stripe, so a SUM
        all_reduce assembles the full frame on every rank, bit-identical to the single-GPU one
        (x + 0 is exact)."""
        if self.world > 1:
            dist.all_reduce(color_partial, op=dist.ReduceOp.SUM, group=self.group)
        return color_partial

    def num_rendered_total(self):
        """Sum over ranks of the per-stripe instance counts (a collective: call on every rank)."""
        n = int(self.last_num_rendered_local or 0)
        if self.world == 1:
            return n
        dev = self.device if self.device is not None else "cpu"
        t = torch.tensor([n], dtype=torch.int64, device=dev)
        dist.all_reduce(t, op=dist.ReduceOp.SUM, group=self.group)
        return int(t.item())

    # -- one frame ---------------------------------------------------------------------------
    def render(self, inp, cam, src=0, broadcast=False, assemble=True, bounds=None):
        """bounds: explicit stripe bounds (device int32 [world + 1]); None = balanced stripes cut by
        the forward itself (self.balanced) or equal-height ones."""
        if broadcast:
            self.broadcast_gaussians(inp, src)
        balanced = self.balanced and bounds is None and self.world > 1
        color, radii, state = self.backend.forward(inp, cam, self.rank, self.world, bounds, balanced=balanced)
        self.last_num_rendered_local = int(state["R"])
        if assemble:
            color = self.assemble_image(color)
        return color, radii, state

    def backward(self, state, inp, cam, grad_out, out=None):
        """-> (grads, owner_mask): dense [P,...] gradient tensors in which this rank has written
        the rows of the Gaussians it owns (all visible ones when world == 1, where the remaining
        rows are zeros as in the reference)."""
        if self.exchange == "peer" and not self._peer_checked:
            # every rank must take the same path: agree once on whether peer mapping works everywhere
            ok = 1
            try:
                self.backend.peer_setup(inp["means3D"].shape[0], self.rank, self.world, self.group)
            except Exception as e:   # no P2P / IPC on this box: fall back to the collective formulation
                ok, self.peer_error = 0, repr(e)
            flag = torch.tensor([ok], device=inp["means3D"].device)
            dist.all_reduce(flag, op=dist.ReduceOp.MIN, group=self.group)
            if int(flag.item()) == 0:
                if ok:
                    self.backend.peer_teardown()
                self.exchange = "collective"
            self._peer_checked = True
        if self.exchange == "peer":
            self.backend.peer_setup(inp["means3D"].shape[0], self.rank, self.world, self.group)
            grads = self.backend.peer_backward(state, inp, cam, grad_out, self.rank, self.world, out=out,
                                               remote_scalar=self.remote_scalar)
        else:
            acc = self.backend.backward_blend(state, inp, cam, grad_out, self.rank, self.world)
            if self.world > 1:
                dist.all_reduce(acc, op=dist.ReduceOp.SUM, group=self.group)
            grads = self.backend.backward_geometry(state, inp, cam, acc, self.rank, self.world, out=out)
        return grads, (lambda: self.backend.owner_mask(state, inp, self.rank))

    def gather_gradients(self, grads, owner_mask):
        """Full gradients on every rank: rows a rank does not own are zeroed, then summed."""
        if self.world == 1:
            return grads
        m = owner_mask() if callable(owner_mask) else owner_mask
        out = []
        for g in grads:
            if g.numel() == 0:
                out.append(g)
                continue
            z = torch.where(m.reshape((-1,) + (1,) * (g.dim() - 1)), g, torch.zeros((), dtype=g.dtype, device=g.device))
            dist.all_reduce(z, op=dist.ReduceOp.SUM, group=self.group)
            out.append(z)
        return tuple(out)

    # -- bench.py helpers -------------------------------------------------------------------------
    @staticmethod
    def _cam(s, inp):
        return dict(bg=inp["bg"], view=inp["view"], proj=inp["proj"], campos=inp["campos"],
                    tanfovx=s.tanfovx, tanfovy=s.tanfovy, img_h=s.img_h, img_w=s.img_w,
                    sh_degree=s.sh_degree)

    def forward_backward(self, s, inp, grad_out, assemble=False, out=None):
        cam = self._cam(s, inp)
        color, radii, state = self.render(inp, cam, assemble=assemble)
        grads, mask = self.backward(state, inp, cam, grad_out, out=out)
        return color, grads, state

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

    def close(self):
        if hasattr(self.backend, "peer_teardown"):
            self.backend.peer_teardown()
