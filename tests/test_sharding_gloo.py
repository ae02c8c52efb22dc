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


class OracleBackend:
    """Emulates one rank's share with the CPU oracle: the full render masked to the rank's tile
    rows; the backward blend restricted to those rows by zeroing dL/dpix elsewhere (the blend
    gradient is linear in dL/dpix, so this equals blending only the rank's tiles)."""

    def __init__(self, s):
        self.s = s

    def _row_mask(self, H, rank, world):
        m = np.zeros(H, dtype=bool)
        for r in sharding.owned_tile_rows((H + 15) // 16, rank, world):
            m[r * 16:(r + 1) * 16] = True
        return m

    def forward(self, inp, cam, rank, world):
        s = self.s
        r = oracle.forward(inp["means3D"].numpy(), inp["opacity"].numpy(), inp["scales"].numpy(),
                           inp["rotations"].numpy(), cam["view"].numpy(), cam["proj"].numpy(),
                           cam["campos"].numpy(), cam["img_w"], cam["img_h"], cam["tanfovx"],
                           cam["tanfovy"], cam["bg"].numpy(), shs=inp["sh"].numpy(),
                           sh_degree=cam["sh_degree"], precision="f32")
        mask = self._row_mask(cam["img_h"], rank, world)
        col = r.color.copy()
        col[:, ~mask, :] = 0
        return torch.from_numpy(col), torch.from_numpy(r.radii), dict(r=r, mask=mask, R=r.num_rendered)

    def backward_blend(self, state, inp, cam, grad_out, rank, world):
        g = grad_out.numpy().copy()
        g[:, ~state["mask"], :] = 0
        b = oracle.backward_blend(state["r"], g)
        P = b["dL_dmean2D"].shape[0]
        acc = np.zeros((P, 12), np.float32)
        acc[:, 0:2], acc[:, 2:5], acc[:, 5:6], acc[:, 6:9] = b["dL_dmean2D"], b["dL_dconic"], b["dL_dopacity"], b["dL_dcolor"]
        return torch.from_numpy(acc)

    def backward_geometry(self, state, inp, cam, grad_acc, start, count):
        a = grad_acc.numpy()[:inp["means3D"].shape[0]].copy()
        valid = np.zeros(a.shape[0], bool)
        valid[start:start + count] = True
        a[~valid] = 0   # rows outside the slice hold uninitialised data after reduce_scatter
        blend = dict(dL_dmean2D=a[:, 0:2], dL_dconic=a[:, 2:5], dL_dopacity=a[:, 5:6], dL_dcolor=a[:, 6:9])
        g = oracle.backward_geometry(state["r"], blend)
        m2 = np.concatenate([a[:, 0:2], np.zeros((a.shape[0], 1), np.float32)], axis=1)
        out = (m2, a[:, 6:9], a[:, 5:6], g["dL_dmean3D"], g["dL_dcov3D"], g["dL_dsh"], g["dL_dscale"], g["dL_drot"])
        return tuple(torch.from_numpy(np.ascontiguousarray(x, dtype=np.float32)) for x in out)


def _free_port():
    with socket.socket() as sk:
        sk.bind(("127.0.0.1", 0))
        return sk.getsockname()[1]


def _worker(rank, world, port, P, W, H, outdir):
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
        eng = sharding.TileShardedRasterizer(backend=OracleBackend(s))
        assert (eng.rank, eng.world) == (rank, world)
        color, radii, state = eng.render(inp, cam, src=0)
        G = torch.randn(3, H, W, generator=torch.Generator().manual_seed(4))
        grads, (start, count) = eng.backward(state, inp, cam, G, gather=True)
        np.savez(os.path.join(outdir, f"rank{rank}.npz"), color=color.numpy(), start=start, count=count,
                 **{n: g.numpy() for n, g in zip(sharding.GRAD_NAMES, grads)})
    finally:
        dist.destroy_process_group()


# P not divisible by the world size, H not a multiple of 16; 3 ranks over 5 tile rows; fewer
# Gaussians than ranks (the last rank's geometry slice is empty)
@pytest.mark.parametrize("P,W,H,world", [(301, 96, 80, 2), (301, 96, 80, 3), (2, 48, 48, 3)],
                         ids=["w2", "w3", "w3_tiny"])
def test_sharded_frame_equals_unsharded(tmp_path, P, W, H, world):
    mp.spawn(_worker, args=(world, _free_port(), P, W, H, str(tmp_path)), nprocs=world, join=True)
    s = uniform_scene(P, W, H, sh_degree=1, seed=31)
    r = oracle.forward_scene(s, "f32")
    G = torch.randn(3, H, W, generator=torch.Generator().manual_seed(4)).numpy()
    g = oracle.backward(r, G)
    ref = dict(dL_dmeans2D=np.concatenate([g["dL_dmean2D"], np.zeros((P, 1), np.float32)], 1),
               dL_dcolors=g["dL_dcolor"], dL_dopacity=g["dL_dopacity"], dL_dmeans3D=g["dL_dmean3D"],
               dL_dcov3D=g["dL_dcov3D"], dL_dsh=g["dL_dsh"], dL_dscales=g["dL_dscale"], dL_drotations=g["dL_drot"])
    outs = [np.load(tmp_path / f"rank{k}.npz") for k in range(world)]
    slices = sorted((int(o["start"]), int(o["count"])) for o in outs)
    assert slices[0][0] == 0 and sum(c for _, c in slices) == P
    for (s0, c0), (s1, _) in zip(slices, slices[1:]):
        assert s0 + c0 == s1
    for o in outs:
        # disjoint rows + x+0 exact => the assembled frame is bit-identical on every rank
        assert np.array_equal(o["color"], r.color)
        for n, v in ref.items():
            den = np.linalg.norm(v)
            assert np.linalg.norm(o[n] - v) <= 2e-5 * (den if den > 0 else 1), n


def test_partition_helpers():
    assert sharding.owned_tile_rows(68, 3, 8) == list(range(3, 68, 8))
    cover = sorted(r for k in range(8) for r in sharding.owned_tile_rows(68, k, 8))
    assert cover == list(range(68))
    for P, world in [(10, 4), (5_000_000, 8), (3, 8), (0, 2)]:
        tot, prev_end = 0, 0
        for k in range(world):
            st, cnt, padded = sharding.gaussian_slice(P, k, world)
            assert padded % world == 0 and padded >= P and st == min(prev_end, P)
            prev_end = st + cnt
            tot += cnt
        assert tot == P
