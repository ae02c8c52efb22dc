"""CPU-side design study (no GPU; lives under tests/ because it drives the oracle, which is test infrastructure): how many survivor iterations would the blend kernels spend per warp
for different pixel-group shapes?  Uses the C oracle's forward state for a bench workload, samples
tiles, and replays the kernels' control flow in numpy:

  * a warp owns an 8x4 pixel block and walks its tile's sorted list 32 records at a time until all
    its pixels are saturated (here: up to the block's furthest last-contributor, rounded up to 32);
  * every record is tested against a group's rectangle with the same conservative quadratic-minimum
    test as gcr_subrect_touch (blend_common.cuh); survivors are evaluated by every lane of the group;
  * with G independent groups per warp the per-chunk trip count is max over groups of the survivors.

Prints, per shape: mean evaluated iterations per warp, lane utilisation (contributing
(pixel, record) pairs / (32 x iterations)) and the cull-test count, from which the expected issue
slots per warp follow (EVAL ~ 50 and TEST ~ 25 SASS instructions, profiles/r01_ncu_full_v3).

Usage: python tests/analysis/sim_lane_util.py [--workload cfg3_1M_sh3_1080p] [--tiles 200]
"""
import argparse
import os
import sys
import time

import numpy as np

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__)))))


def rect_touch(mx, my, A, B, C, twoL, rx0, rx1, ry0, ry1):
    dx0, dx1 = rx0 - mx, rx1 - mx
    dy0, dy1 = ry0 - my, ry1 - my
    ex = np.minimum(np.maximum(0.0, dx0), dx1)
    ey = np.minimum(np.maximum(0.0, dy0), dy1)
    with np.errstate(all="ignore"):
        yv = np.minimum(np.maximum(-B * ex / C, dy0), dy1)
        qv = A * ex * ex + 2 * B * ex * yv + C * yv * yv
        xh = np.minimum(np.maximum(-B * ey / A, dx0), dx1)
        qh = A * xh * xh + 2 * B * xh * ey + C * ey * ey
    qmin = np.where(ex != 0, np.where(ey != 0, np.minimum(qv, qh), qv), np.where(ey != 0, qh, 0.0))
    ax = np.maximum(np.abs(dx0), np.abs(dx1))
    ay = np.maximum(np.abs(dy0), np.abs(dy1))
    mag = A * ax * ax + 2 * np.abs(B) * ax * ay + C * ay * ay
    pd = (A > 0) & (C > 0) & (A * C - B * B > 0)
    return ~(pd & (qmin > twoL + 1e-5 * mag + 1e-3))


SHAPES = {          # name: list of (x0, y0, w, h) groups inside the warp's 8x4 block
    "8x4 (today)": [(0, 0, 8, 4)],
    "2 x 8x2": [(0, 0, 8, 2), (0, 2, 8, 2)],
    "2 x 4x4": [(0, 0, 4, 4), (4, 0, 4, 4)],
    "4 x 4x2": [(0, 0, 4, 2), (4, 0, 4, 2), (0, 2, 4, 2), (4, 2, 4, 2)],
}


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--workload", default="cfg3_1M_sh3_1080p")
    ap.add_argument("--tiles", type=int, default=200)
    ap.add_argument("--seed", type=int, default=0)
    args = ap.parse_args()

    import bench
    from gaussiancity_b200.synthetic import uniform_scene
    from oracle import oracle

    P, W, H, deg, use_sh = bench.WORKLOADS[args.workload]
    t0 = time.time()
    s = uniform_scene(P, W, H, sh_degree=0, seed=0, device="cpu", use_sh=False)  # colours irrelevant here
    r = oracle.forward_scene(s)
    print(f"oracle forward {P} Gaussians {W}x{H}: {time.time() - t0:.1f} s, R = {r.num_rendered}")
    gx, gy = (W + 15) // 16, (H + 15) // 16
    rng = np.random.default_rng(args.seed)
    tiles = rng.choice(gx * gy, size=min(args.tiles, gx * gy), replace=False)

    m2d = r.means2D.astype(np.float64)
    co = r.conic_opacity.astype(np.float64)
    stats = {k: dict(iters=0, tests=0) for k in SHAPES}
    pairs = 0
    warps = 0
    for t in tiles:
        lo, hi = int(r.ranges[t, 0]), int(r.ranges[t, 1])
        if hi <= lo:
            continue
        ids = r.point_list[lo:hi].astype(np.int64)
        mx, my = m2d[ids, 0], m2d[ids, 1]
        A, B, C, o = co[ids, 0], co[ids, 1], co[ids, 2], co[ids, 3]
        twoL = 2 * np.log(255 * o)
        tx, ty = (t % gx) * 16, (t // gx) * 16
        for w in range(8):
            bx, by = tx + (w & 1) * 8, ty + (w >> 1) * 4
            ys, xs = np.meshgrid(np.arange(by, by + 4), np.arange(bx, bx + 8), indexing="ij")
            ok = (xs < W) & (ys < H)
            if not ok.any():
                continue
            ncon = np.where(ok, r.n_contrib[np.minimum(ys, H - 1), np.minimum(xs, W - 1)], 0)
            L = int(ncon.max())
            if L == 0:
                L = min(32, hi - lo)          # one chunk is always looked at
            L = min(((L + 31) // 32) * 32, hi - lo)
            warps += 1
            sl = slice(0, L)
            # contributing (pixel, record) pairs
            dx = mx[sl, None, None] - xs[None]
            dy = my[sl, None, None] - ys[None]
            power = -0.5 * (A[sl, None, None] * dx * dx + C[sl, None, None] * dy * dy) - B[sl, None, None] * dx * dy
            alpha = np.minimum(0.99, o[sl, None, None] * np.exp(np.minimum(power, 0)))
            pos = np.arange(1, L + 1)[:, None, None]
            contrib = (power <= 0) & (alpha >= 1 / 255) & (pos <= ncon[None]) & ok[None]
            pairs += int(contrib.sum())
            for name, groups in SHAPES.items():
                touch = []
                for (gx0, gy0, gw, gh) in groups:
                    touch.append(rect_touch(mx[sl], my[sl], A[sl], B[sl], C[sl], twoL[sl],
                                            bx + gx0, bx + gx0 + gw - 1, by + gy0, by + gy0 + gh - 1))
                touch = np.stack(touch)                              # [G, L]
                pad = (-L) % 32
                if pad:
                    touch = np.pad(touch, ((0, 0), (0, pad)))
                per_chunk = touch.reshape(len(groups), -1, 32).sum(axis=2)   # [G, chunks]
                stats[name]["iters"] += int(per_chunk.max(axis=0).sum())
                stats[name]["tests"] += len(groups) * per_chunk.shape[1]
    print(f"{warps} warps sampled from {len(tiles)} tiles; contributing pairs / warp = {pairs / warps:.1f}")
    EVAL, TEST = 50, 25
    base = None
    for name, st in stats.items():
        it, te = st["iters"] / warps, st["tests"] / warps
        cost = it * EVAL + te * TEST
        base = base or cost
        print(f"{name:>12}: iterations/warp {it:8.1f}  lane utilisation {pairs / (32 * st['iters']):.3f}  "
              f"cull tests/lane/warp {te:6.1f}  est. issue slots/warp {cost:9.0f}  ({cost / base:.3f} of today)")


if __name__ == "__main__":
    main()
