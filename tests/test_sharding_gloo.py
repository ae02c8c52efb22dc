"""CPU, world_size 2 and 3, gloo: the collective plumbing of gaussiancity_b200.sharding
(broadcast -> per-rank tile rows -> image all_reduce; partial [P,12] accumulators ->
reduce_scatter -> per-slice geometry backward -> optional all_gather), with the CPU oracle
plugged in as the compute backend.  The sharded result must equal the unsharded oracle."""
import os
import socket

import numpy as np
import pytest
import torch
import torch.distributed as dist
import torch.multiprocessing as mp

from gaussiancity_b200 import sharding
from gaussiancity_b200.synthetic import uniform_scene
from oracle import oracle


def _rects(r):
    """Tile rects of the oracle's visible Gaussians (auxiliary.h:36-46): (x0, x1, y0, y1) arrays."""
    gx, gy = (r.W + 15) // 16, (r.H + 15) // 16
    px, py = r.means2D[:, 0].astype(np.float32), r.means2D[:, 1].astype(np.float32)
    rad = r.radii.astype(np.float32)
    clip = lambda v, hi: np.minimum(hi, np.maximum(0, np.trunc(v).astype(np.int64)))
    x0, x1 = clip((px - rad) / 16, gx), clip((px + rad + 15) / 16, gx)
    y0, y1 = clip((py - rad) / 16, gy), clip((py + rad + 15) / 16, gy)
    vis = r.radii > 0
    return x0, x1, y0, y1, vis


class OracleBackend:
    """Emulates one rank's share with the CPU oracle: the full render masked to the rank's stripe
    of tile rows; the backward blend restricted to those rows by zeroing dL/dpix elsewhere (the
    blend gradient is linear in dL/dpix, so this equals blending only the rank's tiles); owners =
    the stripe holding each Gaussian's centre tile row (csrc/preprocess.cu gcr_owner_rank)."""

    def __init__(self, s):
        self.s = s
        self._r = None

    def _full(self, inp, cam):
        if self._r is None:
            self._r = oracle.forward(inp["means3D"].numpy(), inp["opacity"].numpy(), inp["scales"].numpy(),
                                     inp["rotations"].numpy(), cam["view"].numpy(), cam["proj"].numpy(),
                                     cam["campos"].numpy(), cam["img_w"], cam["img_h"], cam["tanfovx"],
                                     cam["tanfovy"], cam["bg"].numpy(), shs=inp["sh"].numpy(),
                                     sh_degree=cam["sh_degree"], precision="f32")
        return self._r

    def partition(self, inp, cam, world):
        r = self._full(inp, cam)
        x0, x1, y0, y1, vis = _rects(r)
        rows = np.zeros((r.H + 15) // 16, np.int64)
        for g in np.nonzero(vis)[0]:
            rows[y0[g]:y1[g]] += x1[g] - x0[g]
        assert rows.sum() == r.num_rendered
        return torch.tensor(sharding.balanced_stripes(rows.tolist(), world), dtype=torch.int32)

    def forward(self, inp, cam, rank, world, bounds, balanced=False):
        r = self._full(inp, cam)
        gy = (r.H + 15) // 16
        if bounds is None and balanced:
            bounds = self.partition(inp, cam, world)
        b = bounds.tolist() if bounds is not None else sharding.equal_stripes(gy, world)
        mask = np.zeros(r.H, dtype=bool)
        mask[b[rank] * 16:min(r.H, b[rank + 1] * 16)] = True
        col = r.color.copy()
        col[:, ~mask, :] = 0
        x0, x1, y0, y1, vis = _rects(r)
        centre = np.clip(np.floor(r.means2D[:, 1].astype(np.float32) / 16).astype(np.int64), y0, np.maximum(y0, y1 - 1))
        owner = np.full(r.P, 255, np.uint8)
        owner[vis] = (np.searchsorted(np.asarray(b[1:-1]), centre[vis], side="right")).astype(np.uint8)
        R_local = int(((x1 - x0) * np.maximum(0, np.minimum(y1, b[rank + 1]) - np.maximum(y0, b[rank])))[vis].sum())
        return torch.from_numpy(col), torch.from_numpy(r.radii), dict(r=r, mask=mask, R=R_local, owner=owner)

    def backward_blend(self, state, inp, cam, grad_out, rank, world):
        g = grad_out.numpy().copy()
        g[:, ~state["mask"], :] = 0
        b = oracle.backward_blend(state["r"], g)
        P = b["dL_dmean2D"].shape[0]
        acc = np.zeros((P, 12), np.float32)
        acc[:, 0:2], acc[:, 2:5], acc[:, 5:6], acc[:, 6:9] = b["dL_dmean2D"], b["dL_dconic"], b["dL_dopacity"], b["dL_dcolor"]
        return torch.from_numpy(acc)

    def backward_geometry(self, state, inp, cam, grad_acc, rank, world, out=None, clear=False):
        a = grad_acc.numpy().copy()
        mine = state["owner"] == rank
        a[~mine] = 0   # only the owner differentiates a Gaussian
        blend = dict(dL_dmean2D=a[:, 0:2], dL_dconic=a[:, 2:5], dL_dopacity=a[:, 5:6], dL_dcolor=a[:, 6:9])
        g = oracle.backward_geometry(state["r"], blend)
        m2 = np.concatenate([a[:, 0:2], np.zeros((a.shape[0], 1), np.float32)], axis=1)
        outs = (m2, a[:, 6:9], a[:, 5:6], g["dL_dmean3D"], g["dL_dcov3D"], g["dL_dsh"], g["dL_dscale"], g["dL_drot"])
        res = []
        for x in outs:   # rows this rank does not own are "left untouched": poison them
            x = np.ascontiguousarray(x, dtype=np.float32).copy()
            x[~mine] = np.nan
            res.append(torch.from_numpy(x))
        return tuple(res)

    def owner_mask(self, state, inp, rank):
        return torch.from_numpy(state["owner"] == rank)


def _free_port():
    with socket.socket() as sk:
        sk.bind(("127.0.0.1", 0))
        return sk.getsockname()[1]


def _worker(rank, world, port, P, W, H, outdir, balanced):
    os.environ.update(MASTER_ADDR="127.0.0.1", MASTER_PORT=str(port), RANK=str(rank), WORLD_SIZE=str(world))
    dist.init_process_group("gloo", rank=rank, world_size=world)
    try:
        s = uniform_scene(P, W, H, sh_degree=1, seed=31)
        inp = dict(means3D=s.means3D.clone(), opacity=s.opacities.clone(), scales=s.scales.clone(),
                   rotations=s.rotations.clone(), sh=s.shs.clone())
        if rank != 0:   # only the owning rank has real data before the broadcast
            for t in inp.values():
                t.zero_()
        cam = dict(bg=s.bg, view=s.view_matrix, proj=s.proj_matrix, campos=s.campos, tanfovx=s.tanfovx,
                   tanfovy=s.tanfovy, img_h=H, img_w=W, sh_degree=1)
        eng = sharding.TileShardedRasterizer(backend=OracleBackend(s), exchange="collective", balanced=balanced)
        assert (eng.rank, eng.world) == (rank, world)
        color, radii, state = eng.render(inp, cam, src=0, broadcast=True)
        G = torch.randn(3, H, W, generator=torch.Generator().manual_seed(4))
        grads, owner_mask = eng.backward(state, inp, cam, G)
        own = owner_mask().numpy()
        full = eng.gather_gradients(grads, owner_mask)
        total_R = eng.num_rendered_total()
        np.savez(os.path.join(outdir, f"rank{rank}.npz"), color=color.numpy(), own=own, total_R=total_R,
                 **{n: g.numpy() for n, g in zip(sharding.GRAD_NAMES, full)})
    finally:
        dist.destroy_process_group()


# H not a multiple of 16; 3 ranks over 5 tile rows; fewer Gaussians than ranks (stripes without a
# single Gaussian, ranks that own nothing); equal-height and balanced stripes
@pytest.mark.parametrize("P,W,H,world,balanced", [(301, 96, 80, 2, True), (301, 96, 80, 3, False),
                                                 (301, 96, 80, 3, True), (2, 48, 48, 3, True)],
                         ids=["w2_balanced", "w3_equal", "w3_balanced", "w3_tiny"])
def test_sharded_frame_equals_unsharded(tmp_path, P, W, H, world, balanced):
    mp.spawn(_worker, args=(world, _free_port(), P, W, H, str(tmp_path), balanced), nprocs=world, join=True)
    s = uniform_scene(P, W, H, sh_degree=1, seed=31)
    r = oracle.forward_scene(s, "f32")
    G = torch.randn(3, H, W, generator=torch.Generator().manual_seed(4)).numpy()
    g = oracle.backward(r, G)
    ref = dict(dL_dmeans2D=np.concatenate([g["dL_dmean2D"], np.zeros((P, 1), np.float32)], 1),
               dL_dcolors=g["dL_dcolor"], dL_dopacity=g["dL_dopacity"], dL_dmeans3D=g["dL_dmean3D"],
               dL_dcov3D=g["dL_dcov3D"], dL_dsh=g["dL_dsh"], dL_dscales=g["dL_dscale"], dL_drotations=g["dL_drot"])
    outs = [np.load(tmp_path / f"rank{k}.npz") for k in range(world)]
    # every visible Gaussian has exactly one owner; culled ones have none
    owners = sum(o["own"].astype(np.int64) for o in outs)
    assert np.array_equal(owners, (r.radii > 0).astype(np.int64))
    for o in outs:
        assert int(o["total_R"]) == r.num_rendered        # the stripes' instance counts add up
        # disjoint stripes + x+0 exact => the assembled frame is bit-identical on every rank
        assert np.array_equal(o["color"], r.color)
        for n, v in ref.items():
            den = np.linalg.norm(v)
            assert np.linalg.norm(o[n] - v) <= 2e-5 * (den if den > 0 else 1), n


def test_partition_helpers():
    assert sharding.equal_stripes(68, 8) == [0, 8, 17, 25, 34, 42, 51, 59, 68]
    assert sharding.equal_stripes(2, 3) == [0, 0, 1, 2]                 # more ranks than rows: empty stripes
    rows = [10] * 68
    b = sharding.balanced_stripes(rows, 8)
    assert b[0] == 0 and b[-1] == 68 and all(x <= y for x, y in zip(b, b[1:]))
    assert max(y - x for x, y in zip(b, b[1:])) - min(y - x for x, y in zip(b, b[1:])) <= 1
    # a frame whose lower half holds 9x the work of the upper half: stripes follow the work
    rows = [1] * 34 + [9] * 34
    b = sharding.balanced_stripes(rows, 4)
    work = [sum(rows[x:y]) for x, y in zip(b, b[1:])]
    assert max(work) - min(work) <= 2 * max(rows) and b[1] > 34
    assert sharding.balanced_stripes([0] * 10, 2) == [0, 5, 10]         # nothing visible: equal stripes
    assert sharding.balanced_stripes([5, 0, 0, 0], 3)[-1] == 4
