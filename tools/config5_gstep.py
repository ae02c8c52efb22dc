#!/usr/bin/env python
"""BASELINE config 5: the reference's UNMODIFIED models/generator.py G-step (forward, rasterize,
L1, backward, Adam) on synthetic points + random z, 960x540 render cropped to 640x448, one
replica per GPU under torchrun DDP -- with the reference rasterizer extension vs ours.

  python tools/config5_gstep.py --arm reference|ours|ours_wrapper [--steps 20] [--points 16384]
  python -m torch.distributed.run --nproc-per-node 8 ... tools/config5_gstep.py --arm ours

Arms (VERDICT r1 task 6):
  reference     the reference's own extensions/diff_gaussian_rasterization/__init__.py on top of
                the unmodified reference CUDA extension (oracle/_ref).
  ours          THE SAME reference __init__.py, unmodified, on top of our native module
                `diff_gaussian_rasterization_ext` (gaussiancity_b200/compat, Seam A): the drop-in
                proof -- nothing of the caller changes.
  ours_wrapper  gaussiancity_b200.GaussianRasterizerWrapper(fast_camera=True) (Seam B + the
                host-side camera path of SURVEY 8f-1) handed to the reference's helpers.
  ours_fused    the same wrapper + gaussiancity_b200.adapter in place of the two helpers (SURVEY 8f-2:
                structure-of-arrays attributes, no ones / identity tensors, crop inside the rasterizer).

The reference's Python is imported from $GCR_REFERENCE_ROOT (default: baseline/_ref/GaussianCity,
staged by oracle/build_ref.py; /root/reference where it exists).  Dependencies of files on the
import path that this configuration never executes (spconv / torch_scatter / addict / flash_attn
for PTv3, plyfile for PLY dumps) are stubbed at import time; README.md:152-161 BLDG settings with
PTV3.ENABLED = False.  Prints one JSON line (rank 0): it/s per rank, loss curve.

SURVEY 8f-4 (hash-grid encoder): `--pos-emd HASH_GRID` switches the generator's positional encoder
to the reference's GridEncoder (config.py:121-123: 16 levels x 8 channels); `--encoder global` adds
the reference's GlobalEncoder in front (config.py:118-119: ENCODER_OUT_DIM = 5, the "REST" generator),
so the grid encoder sees 5 coordinates that require a gradient.  Which `grid_encoder_ext` the
reference's unmodified extensions/grid_encoder/__init__.py imports follows the arm: the reference
build (oracle/_ref) for `reference`, ours (gaussiancity_b200/compat) otherwise; `--grid-ext` overrides.
"""
import argparse
import json
import os
import sys
import time
import types

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def stage_imports(arm, grid_ext=None):
    ref_root = os.environ.get("GCR_REFERENCE_ROOT")
    if not ref_root:
        ref_root = "/root/reference" if os.path.isdir("/root/reference/models") else \
            os.path.join(ROOT, "baseline", "_ref", "GaussianCity")
    if not os.path.isdir(os.path.join(ref_root, "models")):
        raise SystemExit(f"reference Python not found under {ref_root} (run oracle/build_ref.py)")
    for name in ("plyfile", "addict", "torch_scatter", "spconv", "spconv.pytorch"):
        try:
            __import__(name)
        except Exception:
            m = types.ModuleType(name)
            if name == "addict":
                m.Dict = dict
            if name == "spconv":
                m.pytorch = types.ModuleType("spconv.pytorch")
                sys.modules["spconv.pytorch"] = m.pytorch
            if name == "spconv.pytorch":
                m.SparseModule = object
            sys.modules[name] = m
    try:
        import flash_attn  # noqa: F401
    except Exception:
        sys.modules["flash_attn"] = types.ModuleType("flash_attn")
    sp = sys.modules.get("spconv.pytorch")
    if sp is not None and not hasattr(sp, "SparseModule"):
        import torch
        sp.SparseModule = torch.nn.Module
    # native modules: the reference's grid_encoder_ext (as is) + one of the two rasterizer modules
    native = os.path.join(ROOT, "oracle", "_ref") if arm == "reference" else \
        os.path.join(ROOT, "gaussiancity_b200", "compat")
    sys.path.insert(0, native)                                  # diff_gaussian_rasterization_ext
    sys.path.insert(0, ref_root)
    sys.path.insert(1, ROOT)
    # grid_encoder_ext: loaded by path and registered, so the choice does not depend on sys.path order
    from tests import refext
    which = grid_ext or ("reference" if arm == "reference" else "ours")
    mod = refext.load_reference_grid_ext() if which == "reference" else refext.load_our_grid_ext()
    if mod is None:
        raise SystemExit("grid_encoder_ext (%s) is not built" % which)
    sys.modules["grid_encoder_ext"] = mod
    return ref_root


class Cfg(dict):
    __getattr__ = dict.__getitem__


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--arm", choices=["reference", "ours", "ours_wrapper", "ours_fused"], required=True)
    ap.add_argument("--steps", type=int, default=20)
    ap.add_argument("--warmup", type=int, default=5)
    ap.add_argument("--points", type=int, default=16384)
    ap.add_argument("--profile", action="store_true", help="host + device time of the rasterizer calls inside the step")
    ap.add_argument("--pos-emd", choices=["SIN_COS", "HASH_GRID"], default="SIN_COS")
    ap.add_argument("--encoder", choices=["none", "global"], default="none")
    ap.add_argument("--grid-ext", choices=["reference", "ours"], default=None)
    args = ap.parse_args()
    import torch
    import torch.distributed as dist
    rank, world = int(os.environ.get("RANK", "0")), int(os.environ.get("WORLD_SIZE", "1"))
    dev = torch.device(f"cuda:{int(os.environ.get('LOCAL_RANK', '0'))}")
    torch.cuda.set_device(dev)
    if world > 1:
        dist.init_process_group("nccl", device_id=dev)
    ref_root = stage_imports(args.arm, args.grid_ext)
    import diff_gaussian_rasterization_ext as native_ext
    import models.generator
    import utils.helpers
    import extensions.diff_gaussian_rasterization as dgr
    import numpy as np

    # README.md:152-161 (building generator) with PTv3 off; the remaining keys are config.py defaults
    use_global = args.encoder == "global"
    cfg = Cfg(ENCODER="GLOBAL" if use_global else None, ENCODER_OUT_DIM=5 if use_global else 3,
              GLOBAL_ENCODER_N_BLOCKS=6, POS_EMD=args.pos_emd,
              HASH_GRID_N_LEVELS=16, HASH_GRID_LEVEL_DIM=8, SIN_COS_FREQ_BENDS=10, Z_DIM=256,
              MLP_HIDDEN_DIM=512, MLP_N_SHARED_LAYERS=1, ATTR_FACTORS={"rgb": 2}, ATTR_N_LAYERS={"rgb": 1},
              PTV3=Cfg(ENABLED=False))
    n_classes, proj_size, scale_factor = 8, 2048, 0.5
    torch.manual_seed(1234 + rank)
    G = models.generator.Generator(cfg, n_classes, proj_size).to(dev)
    if world > 1:
        G = torch.nn.parallel.DistributedDataParallel(G, device_ids=[dev.index])
    opt = torch.optim.Adam(G.parameters(), lr=1e-4, betas=(0.0, 0.999))

    # synthetic batch: lattice points of a ground plane + box buildings seen from a GoogleEarth-style
    # orbit pose (gaussiancity_b200.synthetic.city_points), [1, N, 9] = xyz | scale | instance | rel_xyz | batch
    from gaussiancity_b200.synthetic import CITY_K, CITY_SENSOR, city_points
    pts14, cam_pos, cam_quat = city_points(args.points, seed=7 + rank, device=dev)
    N = pts14.shape[0]
    abs_xyz = pts14[None, :, 0:3].contiguous()
    g = torch.Generator(device="cpu").manual_seed(99 + rank)
    instances = torch.full((1, N, 1), 10.0, device=dev)
    classes = torch.randint(0, n_classes, (1, N, 1), generator=g).to(dev)
    rel_xyz = (torch.rand(1, N, 3, generator=g) * 2 - 1).to(dev)
    scales = utils.helpers.get_point_scales(torch.ones(1, N, 1, device=dev) * scale_factor, classes, [0, 1])
    onehots = utils.helpers.get_one_hot(classes, n_classes)
    bch_idx = torch.zeros(1, N, dtype=torch.long, device=dev)
    proj_uv = utils.helpers.get_projection_uv(abs_xyz, None, proj_size)
    rgb = (torch.rand(1, 3, 448, 640, generator=g) * 2 - 1).to(dev)
    msk = torch.ones(1, 1, 448, 640, device=dev)
    crp = [dict(x=160, y=46, w=640, h=448)]
    # GlobalEncoder inputs (height field + segmentation map); it mean-pools, so a 256x256 synthetic map
    # stands in for the 2048x2048 projection without changing what reaches the grid encoder
    proj_hf = torch.rand(1, 1, 256, 256, generator=g).to(dev) if use_global else None
    proj_seg = torch.nn.functional.one_hot(torch.randint(0, n_classes, (1, 256, 256), generator=g), n_classes) \
        .permute(0, 3, 1, 2).float().to(dev) if use_global else None
    cam_pos_b, cam_quat_b = [np.asarray(cam_pos, dtype=np.float32)], [np.asarray(cam_quat, dtype=np.float32)]

    ours = None
    if args.arm in ("ours_wrapper", "ours_fused"):
        import gaussiancity_b200 as ours
        from gaussiancity_b200 import adapter
        gr = ours.GaussianRasterizerWrapper(CITY_K, CITY_SENSOR, device=dev, fast_camera=True)
    else:
        gr = dgr.GaussianRasterizerWrapper(K=CITY_K, sensor_size=CITY_SENSOR, flip_ud=False, device=dev)
    l1 = torch.nn.L1Loss()
    losses = []
    prof = {"fwd_host_ms": 0.0, "bwd_host_ms": 0.0, "calls": 0, "events": []}
    if args.profile:
        targets = [native_ext]
        if args.arm in ("ours_wrapper", "ours_fused"):
            targets = [ours.dgr_ext]
        for mod in targets:
            for name, key in (("rasterize_gaussians", "fwd"), ("rasterize_gaussians_backward", "bwd")):
                fn = getattr(mod, name)

                def wrapped(*a, _fn=fn, _key=key):
                    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
                    e0.record()
                    t = time.perf_counter()
                    out = _fn(*a)
                    prof[_key + "_host_ms"] += (time.perf_counter() - t) * 1e3
                    e1.record()
                    prof["events"].append((_key, e0, e1))
                    prof["calls"] += 1
                    return out
                setattr(mod, name, wrapped)

    def step(i):
        torch.manual_seed(5000 + i)      # utils.helpers.get_z draws from the global generator
        z = utils.helpers.get_z(instances, cfg.Z_DIM)
        pt_attrs = G(proj_uv, rel_xyz, bch_idx, onehots, z, proj_hf, proj_seg)
        if args.arm == "ours_fused":   # no [B,N,14] tensor, crop folded into the rasterizer
            fake = adapter.get_gaussian_rasterization_fused(abs_xyz, scales, pt_attrs, gr, cam_pos_b, cam_quat_b, crp)
        else:
            gs_pts = utils.helpers.get_gaussian_points(abs_xyz.clone(), scales.clone(), pt_attrs)
            fake = utils.helpers.get_gaussian_rasterization(gs_pts, gr, cam_pos_b, cam_quat_b, crp)
        loss = l1(fake * msk, rgb * msk)
        G.zero_grad()
        loss.backward()
        opt.step()
        return loss

    for i in range(args.warmup):
        losses.append(float(step(i).detach()))
    torch.cuda.synchronize()
    if world > 1:
        dist.barrier()
    t0 = time.perf_counter()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for i in range(args.warmup, args.warmup + args.steps):
        losses.append(step(i).detach())
    e1.record()
    torch.cuda.synchronize()
    wall = time.perf_counter() - t0
    losses = [float(x) for x in losses]
    ms = e0.elapsed_time(e1) / args.steps
    raster = None
    if args.profile:
        n = max(1, prof["calls"] // 2)
        dev_ms = {"fwd": 0.0, "bwd": 0.0}
        for key, a, b in prof["events"]:
            dev_ms[key] += a.elapsed_time(b)
        raster = {"calls_fwd": n, "fwd_host_ms_per_call": prof["fwd_host_ms"] / n, "bwd_host_ms_per_call": prof["bwd_host_ms"] / n,
                  "fwd_stream_ms_per_call": dev_ms["fwd"] / n, "bwd_stream_ms_per_call": dev_ms["bwd"] / n,
                  "note": "host = wall time inside the native call; stream = CUDA-event interval around it on the "
                          "launching stream (includes waiting for earlier generator kernels only if the call synchronises)"}
    if world > 1:
        t = torch.tensor([ms], device=dev, dtype=torch.float64)
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
        ms = float(t.item())
    if rank == 0:
        print(json.dumps({"workload": "config5_generator_gstep", "arm": args.arm, "n_gpus": world,
                          "points": int(N), "render": "960x540", "crop": "640x448", "steps": args.steps,
                          "ms_per_step": ms, "it_per_s_per_rank": 1000.0 / ms, "it_per_s_total": world * 1000.0 / ms,
                          "wall_ms_per_step": wall * 1e3 / args.steps, "losses": losses,
                          "rasterizer": raster,
                          "rasterizer_module": native_ext.__file__.replace(ROOT + "/", ""),
                          "dgr_python": dgr.__file__.replace(ref_root, "<reference>"),
                          "grid_encoder_module": sys.modules["grid_encoder_ext"].__file__.replace(ROOT + "/", ""),
                          "generator": f"reference models/generator.py, ENCODER={cfg.ENCODER} POS_EMD={cfg.POS_EMD} "
                                       "Z_DIM=256 PTV3 off"}))
    if world > 1:
        dist.destroy_process_group()


if __name__ == "__main__":
    main()
