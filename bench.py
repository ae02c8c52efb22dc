#!/usr/bin/env python
"""bench.py -- headline benchmark of the B200-native Gaussian rasterizer hot path.

Metric (BASELINE.json): Msplats/s = P / t(fwd+bwd) / 1e6 on synthetic Gaussian clouds; also the
rendered MPix/s of the forward alone and the roofline fraction of the dominant kernel.
Default workload (N=1): BASELINE config 4 -- 5 M Gaussians, SH degree 3, 1920x1080, fwd+bwd on
ONE GPU (it fits: ~4 GB).  `--workload` selects the smaller parity configs.

  python bench.py [--gpus N] [--steps K] [--warmup W] [--impl ours|reference] [--workload NAME]

One JSON line on stdout (rank 0).  Timing: W untimed warm-up steps, then exactly K steps
bracketed by barrier + torch.cuda.synchronize(), CUDA events on the launching stream, max over
ranks.  `value` is measured with inputs resident in HBM; `e2e` is the same metric through the
public API with HOST (pinned) buffers, host->device copies of every input and a device->host
read of the rendered image and loss inside the timed region.
`--impl reference` runs the UNMODIFIED reference CUDA extension (oracle/_ref, built from
/root/reference by oracle/build_ref.py) on the same workload through the same harness; if it is
not available the CPU oracle port is timed instead on a bounded sample.
Inputs (1.2 GB at the default workload) are far larger than the 126 MB L2, so no explicit L2
flush is needed between timed iterations (stated in config.l2).
"""
import argparse
import json
import os
import statistics
import subprocess
import sys
import threading
import time

ROOT = os.path.dirname(os.path.abspath(__file__))
if ROOT not in sys.path:
    sys.path.insert(0, ROOT)

import torch  # noqa: E402
import torch.distributed as dist  # noqa: E402

WORKLOADS = {
    # name: (P, W, H, sh_degree, use_sh)
    "cfg4_5M_sh3_1080p": (5_000_000, 1920, 1080, 3, True),
    "cfg3_1M_sh3_1080p": (1_000_000, 1920, 1080, 3, True),
    "cfg2_100k_sh0_512": (100_000, 512, 512, 0, True),
    "city_5M_precomp_1080p": (5_000_000, 1920, 1080, 0, False),
    "tiny": (20_000, 256, 256, 1, True),
    # GaussianCity's own call pattern (BASELINE config 5 scale): lattice points, K/sensor camera
    "cfg5_city_16k_540p": (16_384, 960, 540, 0, False),
    "city_500k_540p": (500_000, 960, 540, 0, False),
}
DEFAULT_WORKLOAD = "cfg4_5M_sh3_1080p"


# ------------------------------------------------------------------------------------------
class ClockSampler:
    """nvidia-smi clocks / throttle reasons sampled DURING the timed region."""
    Q = ("index,clocks.sm,clocks.max.sm,power.draw,clocks_event_reasons.hw_slowdown,"
         "clocks_event_reasons.hw_thermal_slowdown,clocks_event_reasons.sw_thermal_slowdown,"
         "clocks_event_reasons.sw_power_cap")

    def __init__(self, gpu_index=0):
        self.gpu_index = gpu_index
        self.samples = []
        self.proc = None

    def start(self):
        try:
            self.proc = subprocess.Popen(
                ["nvidia-smi", f"--query-gpu={self.Q}", "--format=csv,noheader,nounits", "-lms", "100",
                 "-i", str(self.gpu_index)], stdout=subprocess.PIPE, stderr=subprocess.DEVNULL, text=True)
            self.t = threading.Thread(target=self._read, daemon=True)
            self.t.start()
        except Exception:
            self.proc = None

    def _read(self):
        for line in self.proc.stdout:
            self.samples.append(line.strip())

    def stop(self):
        if self.proc is None:
            return {"sm_mhz": None, "sm_max_mhz": None, "reasons": ["nvidia-smi unavailable"]}
        time.sleep(0.15)
        self.proc.terminate()
        try:
            self.proc.wait(timeout=2)
        except Exception:
            self.proc.kill()
        sm, mx, reasons = [], [], set()
        for s in self.samples:
            f = [x.strip() for x in s.split(",")]
            if len(f) < 8:
                continue
            try:
                sm.append(float(f[1])); mx.append(float(f[2]))
            except ValueError:
                continue
            for name, v in zip(("hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown", "sw_power_cap"), f[4:8]):
                if v.lower().startswith("active"):
                    reasons.add(name)
        return {"sm_mhz": statistics.median(sm) if sm else None,
                "sm_max_mhz": max(mx) if mx else None, "reasons": sorted(reasons),
                "samples": len(sm)}


def load_peaks():
    p = os.path.join(ROOT, "MEASURED_PEAKS.json")
    if os.path.exists(p):
        with open(p) as fh:
            d = json.load(fh)
        return float(d["hbm_gbs"]), "measured (MEASURED_PEAKS.json hbm_gbs)"
    return 6650.0, "fallback (B200_PROFILING.md: 6.65 TB/s)"


def tile_sort_passes(W, H):
    tiles = ((W + 15) // 16) * ((H + 15) // 16)
    bits = max(1, (tiles - 1).bit_length())
    return (bits + 7) // 8


def kernels_per_step(W, H, backward=True):
    """Kernels of OUR library launched by one forward(+backward) (memsets / copies excluded):
    preprocess, depth sort (1 histogram + 4 onesweep passes), scan+emit, tile sort passes,
    tile ranges, blend_fwd [+ blend_bwd, geometry_bwd]."""
    fwd = 1 + (1 + 4) + 1 + tile_sort_passes(W, H) + 1 + 1
    return fwd + (2 if backward else 0)


# ------------------------------------------------------------------------------------------
def make_scene(name, device):
    from gaussiancity_b200.synthetic import city_scene, uniform_scene
    P, W, H, deg, use_sh = WORKLOADS[name]
    if "city" in name and "precomp" not in name:
        return city_scene(P, seed=0, device=device)
    return uniform_scene(P, W, H, sh_degree=deg, seed=0, device=device, use_sh=use_sh)


def scene_inputs(s):
    e = torch.Tensor([])
    return dict(bg=s.bg, means3D=s.means3D, colors=s.colors_precomp if s.colors_precomp is not None else e,
                opacity=s.opacities, scales=s.scales, rotations=s.rotations,
                sh=s.shs if s.shs is not None else e, view=s.view_matrix, proj=s.proj_matrix,
                campos=s.campos)


def fwd_args(s, inp):
    e = torch.Tensor([])
    return (inp["bg"], inp["means3D"], inp["colors"], inp["opacity"], inp["scales"], inp["rotations"],
            1.0, e, inp["view"], inp["proj"], s.tanfovx, s.tanfovy, s.img_h, s.img_w, inp["sh"],
            s.sh_degree, inp["campos"], False, False)


def bwd_args(s, inp, radii, grad_out, geom, R, binning, img):
    e = torch.Tensor([])
    return (inp["bg"], inp["means3D"], radii, inp["colors"], inp["scales"], inp["rotations"], 1.0, e,
            inp["view"], inp["proj"], s.tanfovx, s.tanfovy, grad_out, inp["sh"], s.sh_degree,
            inp["campos"], geom, R, binning, img, False)


def time_steps(step, steps, warmup, barrier):
    for _ in range(warmup):
        step()
    torch.cuda.synchronize()
    barrier()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(steps):
        step()
    e1.record()
    torch.cuda.synchronize()
    barrier()
    return e0.elapsed_time(e1) / steps


def max_over_ranks(x, device, world):
    if world == 1:
        return x
    t = torch.tensor([x], dtype=torch.float64, device=device)
    dist.all_reduce(t, op=dist.ReduceOp.MAX)
    return float(t.item())


# ------------------------------------------------------------------------------------------
def _relerr(a, b):
    den = b.double().norm().item()
    return (a.double() - b.double()).norm().item() / (den if den > 0 else 1.0)


def run_sharded_arm(args, s, inp, grad_out, rank, world, device, barrier):
    """N > 1: ONE frame split into `world` contiguous, balanced tile-row stripes
    (gaussiancity_b200/sharding.py).  Before anything is timed, the sharded frame and gradients are
    checked once against this rank's own single-GPU render of the same scene."""
    from gaussiancity_b200 import _cabi, ext, sharding
    P, W, H = s.means3D.shape[0], s.img_w, s.img_h
    eng = sharding.TileShardedRasterizer(device=device, exchange=args.exchange, balanced=not args.equal_stripes,
                                         remote_scalar=args.remote_scalar)
    cam = eng._cam(s, inp)
    M = inp["sh"].shape[1] if inp["sh"].numel() else 0
    opts = dict(dtype=torch.float32, device=device)
    # persistent gradient storage: packed 96-byte rows + dL_dsh (sharding.CudaBackend.backward_geometry)
    out = (torch.empty((P, 24), **opts), torch.empty((P, M, 3), **opts))

    # ---- correctness gate (untimed): bit-identical frame, gradients <= 1e-4 --------------------
    check = {}
    if not args.no_check:
        R1, col1, radii1, g1, b1, i1 = ext.rasterize_gaussians(*fwd_args(s, inp))
        grads1 = ext.rasterize_gaussians_backward(*bwd_args(s, inp, radii1, grad_out, g1, R1, b1, i1))
        color, radii, state = eng.render(inp, cam, assemble=True)
        grads, owner_mask = eng.backward(state, inp, cam, grad_out)
        full = eng.gather_gradients(grads, owner_mask)
        Rtot = eng.num_rendered_total()
        torch.cuda.synchronize()
        frame_ok = bool(torch.equal(color, col1)) and bool(torch.equal(radii, radii1)) and Rtot == R1
        errs = {n: _relerr(a, b) for n, a, b in zip(sharding.GRAD_NAMES, full, grads1) if b.numel()}
        ok = frame_ok and all(e <= 1e-4 for e in errs.values())
        flag = torch.tensor([1 if ok else 0], device=device)
        dist.all_reduce(flag, op=dist.ReduceOp.MIN)
        check = {"frame_bit_identical": frame_ok, "num_rendered_total": Rtot, "max_grad_relerr": max(errs.values()),
                 "all_ranks_ok": bool(flag.item())}
        if not bool(flag.item()):
            raise SystemExit(f"[bench] rank {rank}: sharded frame != single-GPU frame: {check} {errs}")
        del grads1, full, g1, b1, i1, col1
    else:
        eng.render(inp, cam, assemble=False)
    R_total = eng.num_rendered_total()
    R_local = eng.last_num_rendered_local

    big = ("means3D", "opacity", "scales", "rotations", "sh", "colors")
    if args.shard_mode == "broadcast":
        # rank 0 owns the Gaussians and NCCL-broadcasts all buffers every frame, prefetched one frame
        # ahead on a side stream into the other of two buffer sets
        sets = [inp, {k: (v.clone() if k in big and v.numel() else v) for k, v in inp.items()}]
        st = {"i": 0, "h": None}

        def step():
            i = st["i"]
            if st["h"] is None and i == 0:
                st["h"] = eng.start_prefetch({k: sets[0][k] for k in big}, src=0)
            eng.wait_prefetch(st["h"])
            nxt = eng.start_prefetch({k: sets[(i + 1) % 2][k] for k in big}, src=0)
            o = eng.forward_backward(s, sets[i % 2], grad_out, assemble=args.assemble, out=out)
            st["i"], st["h"] = i + 1, nxt
            return o
    else:
        def step():
            return eng.forward_backward(s, inp, grad_out, assemble=args.assemble, out=out)

    for _ in range(args.warmup):
        step()
    torch.cuda.synchronize()
    _cabi.profile_enable(True)
    ms = max_over_ranks(time_steps(step, args.steps, 0, barrier), device, world)
    stages = _cabi.profile_read()
    _cabi.profile_enable(False)
    fwd_only = lambda: eng.render(inp, cam, assemble=args.assemble)
    ms_fwd = max_over_ranks(time_steps(fwd_only, max(2, args.steps // 2), 2, barrier), device, world)
    # the other assembly mode, for the record
    alt = lambda: eng.forward_backward(s, inp, grad_out, assemble=not args.assemble, out=out)
    ms_alt = max_over_ranks(time_steps(alt, max(2, args.steps // 2), 2, barrier), device, world)
    # slowest rank's stage times (max over ranks per stage): what bounds the frame
    if stages:
        names = sorted(stages)
        t = torch.tensor([stages[n] for n in names], dtype=torch.float64, device=device)
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
        stages = {n: float(v) for n, v in zip(names, t.tolist())}
    rl = torch.tensor([float(R_local)], dtype=torch.float64, device=device)
    rmax = rl.clone()
    dist.all_reduce(rmax, op=dist.ReduceOp.MAX)

    # ---- e2e at N GPUs: every rank uploads 1/N of each input over ITS OWN PCIe link, the slices are
    # all-gathered over NVLink, the frame is rendered / differentiated in stripes, every rank reads
    # its own image stripe and its part of the loss back ----------------------------------------------
    e2e = None
    if not args.no_e2e:
        big_keys = [k for k in big if inp[k].numel()]
        per = (P + world - 1) // world
        lo, hi = min(P, rank * per), min(P, (rank + 1) * per)
        host = {k: inp[k][lo:hi].cpu().pin_memory() for k in big_keys}
        slab = {k: torch.zeros((per,) + tuple(inp[k].shape[1:]), **opts) for k in big_keys}
        gathered = {k: torch.empty((per * world,) + tuple(inp[k].shape[1:]), **opts) for k in big_keys}
        _c, _r, st0 = eng.render(inp, cam, assemble=False)
        bounds_h = eng.backend.stripe_bounds(st0, inp, world).cpu().tolist()
        del _c, _r, st0
        r0, r1 = bounds_h[rank] * 16, min(H, bounds_h[rank + 1] * 16)
        host_img = torch.empty(3, max(1, r1 - r0), W, dtype=torch.float32).pin_memory()
        host_loss = torch.empty(1, dtype=torch.float32).pin_memory()

        def step_e2e():
            for k in big_keys:
                slab[k][:hi - lo].copy_(host[k], non_blocking=True)
                dist.all_gather_into_tensor(gathered[k], slab[k])
            frame = dict(inp)
            for k in big_keys:
                frame[k] = gathered[k][:P]
            color, grads, state = eng.forward_backward(s, frame, grad_out, assemble=False, out=out)
            if r1 > r0:
                host_img[:, :r1 - r0].copy_(color[:, r0:r1], non_blocking=True)
                host_loss.copy_((color[:, r0:r1] * grad_out[:, r0:r1]).sum().reshape(1), non_blocking=True)
        ms_e = max_over_ranks(time_steps(step_e2e, max(2, args.steps // 2), 2, barrier), device, world)
        h2d = sum(host[k].numel() * 4 for k in big_keys)
        e2e = {"value": P / (ms_e * 1e-3) / 1e6, "unit": "Msplats/s", "ms_per_step": ms_e,
               "h2d_bytes_per_step": int(h2d) * world, "d2h_bytes_per_step": int(3 * H * W * 4 + 4 * world),
               "note": f"every rank uploads 1/{world} of each input from pinned host memory over its own PCIe "
                       f"link, all_gather over NVLink, striped fwd+bwd, every rank reads its image stripe + "
                       f"partial loss back (bytes are whole-job totals)"}
    eng.close()
    mode = ("Gaussians resident on every rank" if args.shard_mode == "replicated" else
            "rank 0 NCCL-broadcasts all Gaussian buffers every frame (prefetched one frame ahead)")
    exch = ("per-Gaussian gradient sums added straight into the owner rank's accumulator over NVLink peer "
            "memory inside the blend kernel + one cross-GPU barrier kernel" if eng.exchange == "peer" else
            "partial [P,12] accumulators all_reduced by NCCL" +
            (f" (peer mapping unavailable: {eng.peer_error})" if args.exchange == "peer" else ""))
    extra = {"parallelism": f"tile-row stripes x{world} ({'equal-height' if args.equal_stripes else 'balanced by per-row instance counts'}), "
                            f"{mode}; {exch}; frame {'assembled on every rank (all_reduce)' if args.assemble else 'left in stripes'}; "
                            f"gradients left with their owners",
             "shard_mode": args.shard_mode, "exchange": eng.exchange, "assemble": bool(args.assemble),
             "ms_other_assembly_mode": ms_alt, "num_rendered_local_max": int(rmax.item()),
             "stripe_bounds": bounds_h if not args.no_e2e else None, "check": check}
    return dict(ms=ms, ms_fwd=ms_fwd, R=R_total, V=None, stages=stages, P=P, W=W, H=H, s=s, inp=inp,
                grad_out=grad_out, e2e=e2e, extra=extra)


# ------------------------------------------------------------------------------------------
def run_gpu_arm(args, impl, rank, world, device):
    s = make_scene(args.workload, device)
    P, W, H = s.means3D.shape[0], s.img_w, s.img_h
    inp = scene_inputs(s)
    g = torch.Generator().manual_seed(123)
    grad_out = torch.randn(3, H, W, generator=g).to(device)
    barrier = (lambda: dist.barrier()) if world > 1 else (lambda: None)
    extra = {}

    if impl == "ours":
        from gaussiancity_b200 import _cabi, ext
        _cabi.lib()  # fail loudly if the CUDA library is missing
        if world > 1:
            return run_sharded_arm(args, s, inp, grad_out, rank, world, device, barrier)
        import gaussiancity_b200 as pkg
        mod = pkg.dgr_ext                      # the binding the public operator surface uses
        extra["host_binding"] = pkg.HOST_BINDING
    else:
        from tests import refext
        mod = refext.load_reference_ext()
        if mod is None:
            return None

    state = {}

    def step():
        R, color, radii, geom, binning, img = mod.rasterize_gaussians(*fwd_args(s, inp))
        grads = mod.rasterize_gaussians_backward(*bwd_args(s, inp, radii, grad_out, geom, R, binning, img))
        state["R"], state["radii"], state["grads"], state["color"] = R, radii, grads, color

    def fwd_only():
        state["fwd"] = mod.rasterize_gaussians(*fwd_args(s, inp))

    # warm up first (lazy module loading, allocator), then profile exactly the timed steps
    for _ in range(args.warmup):
        step()
    torch.cuda.synchronize()
    if impl == "ours":
        _cabi.profile_enable(True)
    ms = time_steps(step, args.steps, 0, barrier)
    stages = None
    if impl == "ours":
        stages = _cabi.profile_read()
        _cabi.profile_enable(False)
    ms_fwd = time_steps(fwd_only, max(2, args.steps // 2), 3, barrier)
    R = int(state["R"])
    V = int((state["radii"] > 0).sum().item())
    return dict(ms=ms, ms_fwd=ms_fwd, R=R, V=V, stages=stages, P=P, W=W, H=H, s=s, inp=inp,
                grad_out=grad_out, extra=extra)


def run_e2e(args, impl, res, device):
    """Same metric through the public API with HOST buffers: per step, H2D of every input from
    pinned memory, forward + backward, D2H of the rendered image and the loss."""
    s, P, W, H = res["s"], res["P"], res["W"], res["H"]
    host = {k: (v.cpu().pin_memory() if v.numel() else v) for k, v in res["inp"].items()}
    dev = {k: (torch.empty_like(v, device=device) if v.numel() else v) for k, v in host.items()}
    grad_out = res["grad_out"]
    host_img = torch.empty(3, H, W, dtype=torch.float32).pin_memory()
    host_loss = torch.empty(1, dtype=torch.float32).pin_memory()
    h2d = sum(v.numel() * v.element_size() for v in host.values() if v.numel())
    d2h = host_img.numel() * 4 + 4

    if impl == "ours":
        from gaussiancity_b200 import GaussianRasterizationSettings, GaussianRasterizer

        def step():
            for k, v in host.items():
                if v.numel():
                    dev[k].copy_(v, non_blocking=True)
            settings = GaussianRasterizationSettings(
                img_h=H, img_w=W, tanfovx=s.tanfovx, tanfovy=s.tanfovy, bg=dev["bg"], scale_modifier=1.0,
                view_matrix=dev["view"], proj_matrix=dev["proj"], sh_degree=s.sh_degree,
                campos=dev["campos"], prefiltered=False, debug=False)
            leaves = {k: dev[k].requires_grad_(True) for k in ("means3D", "opacity", "scales", "rotations")}
            has_sh = dev["sh"].numel() > 0
            col_leaf = (dev["sh"] if has_sh else dev["colors"]).requires_grad_(True)
            means2D = torch.zeros_like(dev["means3D"], requires_grad=True)
            color, _ = GaussianRasterizer(settings)(
                leaves["means3D"], means2D, leaves["opacity"], shs=col_leaf if has_sh else None,
                colors_precomp=None if has_sh else col_leaf, scales=leaves["scales"],
                rotations=leaves["rotations"])
            loss = (color * grad_out).sum()
            loss.backward()
            host_img.copy_(color.detach(), non_blocking=True)
            host_loss.copy_(loss.detach().reshape(1), non_blocking=True)
            for t in list(leaves.values()) + [col_leaf]:
                t.grad = None
                t.requires_grad_(False)
    else:
        from tests import refext
        mod = refext.load_reference_ext()

        def step():
            for k, v in host.items():
                if v.numel():
                    dev[k].copy_(v, non_blocking=True)
            R, color, radii, geom, binning, img = mod.rasterize_gaussians(*fwd_args(s, dev))
            loss = (color * grad_out).sum()
            mod.rasterize_gaussians_backward(*bwd_args(s, dev, radii, grad_out, geom, R, binning, img))
            host_img.copy_(color, non_blocking=True)
            host_loss.copy_(loss.reshape(1), non_blocking=True)

    ms = time_steps(step, max(2, args.steps // 2), 2, lambda: None)
    return {"value": P / (ms * 1e-3) / 1e6, "unit": "Msplats/s", "ms_per_step": ms,
            "h2d_bytes_per_step": int(h2d), "d2h_bytes_per_step": int(d2h)}


def run_cpu_baseline(workload, budget_s=20.0):
    """CPU oracle port (oracle/gs_oracle.c, OpenMP in the blend stages) on a bounded sample of
    the same workload: `n` Gaussians of the same distribution at full resolution, fwd + bwd."""
    import numpy as np
    from gaussiancity_b200.synthetic import uniform_scene
    from oracle import oracle as cpu_oracle
    P, W, H, deg, use_sh = WORKLOADS[workload]
    n = min(P, 1_000_000)
    s = uniform_scene(n, W, H, sh_degree=deg, seed=0, device="cpu", use_sh=use_sh)
    G = np.random.default_rng(0).standard_normal((3, H, W)).astype(np.float32)
    cores = os.cpu_count() or 1
    os.environ.setdefault("OMP_NUM_THREADS", str(cores))
    t0 = time.time()
    reps = 0
    while True:
        r = cpu_oracle.forward_scene(s, "f32")
        cpu_oracle.backward(r, G)
        reps += 1
        if time.time() - t0 > budget_s * 0.5 or reps >= 6:
            break
    dt = (time.time() - t0) / reps
    # BASELINE config 1 as worded in the north star: naive PyTorch per-pixel blend on the host CPU
    # (1 k Gaussians -> 128x128, forward only). Reported next to the oracle port; never a target.
    naive = None
    try:
        from oracle import torch_naive
        torch.set_num_threads(cores)
        n1 = 100   # bounded sample of config 1 (the python loop over Gaussians costs ~50 ms each on a 128-core host)
        s1 = uniform_scene(n1, 128, 128, sh_degree=3, seed=0, device="cpu")
        t1 = time.time()
        torch_naive.render_naive(s1)
        d1 = time.time() - t1
        naive = {"workload": f"config 1 sample: {n1} of 1k Gaussians, SH 3, 128x128, forward only", "seconds": d1,
                 "Msplats_per_s_forward": n1 / d1 / 1e6, "MPix_per_s": 128 * 128 / d1 / 1e6}
    except Exception as e:  # the reported baseline must never take the bench line down
        naive = {"error": repr(e)[:200]}
    return {"value": n / dt / 1e6, "unit": "Msplats/s", "cores": cores, "kind": "port",
            "torch_naive_config1": naive,
            "sample": f"{n} Gaussians from the distribution of {workload} at {W}x{H}, fwd+bwd, {reps} rep(s), "
                      f"{dt:.2f} s each (oracle/gs_oracle.c fp32; preprocess+sort serial, blend OpenMP)"}


# ------------------------------------------------------------------------------------------
def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=10)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--impl", choices=["ours", "reference"], default="ours")
    ap.add_argument("--workload", choices=sorted(WORKLOADS) + ["grid_encoder"], default=DEFAULT_WORKLOAD)
    ap.add_argument("--shard-mode", choices=["broadcast", "replicated"], default="replicated",
                    help="N>1: 'replicated' (default; consistent with `value` = inputs resident in HBM) = "
                         "every rank holds the Gaussians; 'broadcast' = rank 0 owns them and NCCL-broadcasts all "
                         "buffers every frame (1.18 GB at the default workload: NVLink-bound, see DESIGN.md 5)")
    ap.add_argument("--exchange", choices=["peer", "collective"], default="peer",
                    help="N>1 gradient exchange: 'peer' = blend kernels add into the owner rank's accumulator over "
                         "NVLink peer memory (default); 'collective' = NCCL all_reduce of partial accumulators")
    ap.add_argument("--assemble", action="store_true", help="N>1: assemble the full frame on every rank inside the timed step")
    ap.add_argument("--equal-stripes", action="store_true", help="N>1: equal-height stripes instead of balanced ones")
    ap.add_argument("--remote-scalar", action="store_true", help="N>1 peer exchange: scalar atomics for remote accumulators")
    ap.add_argument("--no-check", action="store_true", help="N>1: skip the untimed bit-identity gate")
    ap.add_argument("--no-e2e", action="store_true")
    ap.add_argument("--no-cpu-baseline", action="store_true")
    args = ap.parse_args()
    args.warmup = max(args.warmup, 3)
    if args.workload == "grid_encoder":
        # the generator's hash-grid positional encoder, ours vs the reference extension
        # (one GPU; bench_grid_encoder.py holds the leg and documents the line it prints)
        if int(os.environ.get("RANK", "0")) == 0:
            import bench_grid_encoder
            bench_grid_encoder.main(["--impl", args.impl, "--steps", str(args.steps), "--warmup", str(args.warmup)] +
                                    (["--no-cpu-baseline"] if args.no_cpu_baseline else []))
        return

    rank = int(os.environ.get("RANK", "0"))
    world = int(os.environ.get("WORLD_SIZE", "1"))
    local_rank = int(os.environ.get("LOCAL_RANK", "0"))
    if not torch.cuda.is_available():
        print(json.dumps({"error": "no CUDA device: this benchmark has no CPU path"}))
        sys.exit(1)
    device = torch.device(f"cuda:{local_rank}")
    torch.cuda.set_device(device)

    if args.impl == "reference" and world > 1:
        # reference arm: rank 0 alone runs and prints; the other ranks exit without work
        if rank != 0:
            return
        world_eff = 1
    else:
        world_eff = world
        if world > 1:
            dist.init_process_group(backend="nccl", device_id=device)

    P, W, H, deg, use_sh = WORKLOADS[args.workload]
    sampler = ClockSampler(local_rank)
    if rank == 0:
        sampler.start()
    res = run_gpu_arm(args, args.impl, rank, world_eff, device)
    clocks = sampler.stop() if rank == 0 else None

    line = {"metric": "Msplats/s fwd+bwd", "unit": "Msplats/s", "n_gpus": args.gpus, "steps": args.steps,
            "warmup": args.warmup, "higher_is_better": True, "dtype": "f32", "data": "synthetic",
            "vs_baseline": None}
    config = {"workload": args.workload, "gaussians": P, "image": f"{W}x{H}", "sh_degree": deg,
              "colour_path": "sh" if use_sh else "colors_precomp", "pass": "forward+backward",
              "seed": 0, "timing": "mean over the K timed steps, one CUDA-event pair around them",
              "l2": "inputs (44+12M B/Gaussian = %.2f GB) exceed the 126 MB L2; no explicit flush"
                    % ((44 + (12 * (deg + 1) ** 2 if use_sh else 12)) * P / 1e9)}

    if res is None:
        # reference extension unavailable -> CPU oracle port on a bounded sample
        if rank == 0:
            cb = run_cpu_baseline(args.workload)
            line.update({"impl": "reference", "value": cb["value"], "ms_per_step": None,
                         "scaling": "strong", "config": config, "cpu_baseline": cb,
                         "e2e": {"value": cb["value"], "unit": cb["unit"], "h2d_bytes_per_step": 0,
                                 "d2h_bytes_per_step": 0},
                         "note": "oracle/_ref not built: timed the CPU oracle port instead"})
            print(json.dumps(line))
        return

    ms, ms_fwd, R, V = res["ms"], res["ms_fwd"], res["R"], res["V"]
    if rank == 0:
        value = P / (ms * 1e-3) / 1e6
        line.update({"value": value, "ms_per_step": ms, "scaling": "strong" if world_eff > 1 else "weak",
                     "mpix_per_s_forward": W * H / (ms_fwd * 1e-3) / 1e6, "ms_forward": ms_fwd,
                     "num_rendered": R, "visible": V,
                     "mean_tile_list": R / (((W + 15) // 16) * ((H + 15) // 16))})
        if args.impl == "reference":
            line["impl"] = "reference"
            line["gpu_launches"] = 0
            line["impl_note"] = "unmodified DGR CUDA extension (oracle/_ref) on 1 B200, native entry points"
            line["cpu_baseline"] = {"value": value, "unit": "Msplats/s", "cores": 0, "kind": "reference",
                                    "sample": "full workload on the GPU: the reference has no CPU "
                                              "implementation of this path (SURVEY.md 8c)"}
        else:
            # per rank; a striped frame adds the stripe-select kernel and the cross-GPU barrier kernel
            line["gpu_launches"] = (kernels_per_step(W, H) + (2 if world_eff > 1 else 0)) * args.steps
            line["parallelism"] = res["extra"].pop("parallelism", "single GPU")
            if res["extra"]:
                line["sharding"] = res["extra"]
        # roofline of the dominant kernel (per-stage CUDA events over the timed region)
        peak, peak_src = load_peaks()
        stages = res["stages"]
        if stages:
            T = ((W + 15) // 16) * ((H + 15) // 16)
            Npix = W * H
            Rk = R
            if world_eff > 1:
                # per launch on the slowest rank: its own stripe (instances: the largest stripe's count;
                # tiles / pixels: 1/N of the frame; visible Gaussians of a stripe are not read back)
                Rk = (line.get("sharding") or {}).get("num_rendered_local_max") or R // world_eff
                T, Npix = T // world_eff, Npix // world_eff
            alg = {"blend_bwd": 8 * T + 40 * Rk + 20 * Npix + 44 * (V or 0),
                   "blend_fwd": 8 * T + 40 * Rk + 20 * Npix}
            dom = max(("blend_bwd", "blend_fwd"), key=lambda k: stages.get(k, 0.0))
            dur = stages[dom]
            achieved = alg[dom] / (dur * 1e-3) / 1e9
            traffic, issue = None, None
            tpath = os.path.join(ROOT, "profiles", "r02_traffic.json")
            if os.path.exists(tpath):
                with open(tpath) as fh:
                    tj = json.load(fh)
                if tj.get("workload") == args.workload and world_eff == 1:   # per-launch DRAM bytes from the committed
                    kj = tj["kernels"].get(dom, {})       # ncu --set full capture
                    traffic = kj.get("dram_bytes")
                    winst = kj.get("warp_instructions")
                    sm_mhz = (clocks or {}).get("sm_mhz")
                    if winst and sm_mhz:
                        # what actually binds the blend kernels: warp instructions issued (from the
                        # same ncu capture) / live kernel time, against 148 SMs x 4 schedulers x
                        # 1 instruction/clk at the SM clock sampled during this run
                        ipeak = 148 * 4 * sm_mhz * 1e6
                        iach = winst / (dur * 1e-3)
                        issue = {"warp_instructions": winst, "achieved_ginst_s": iach / 1e9,
                                 "peak_ginst_s": ipeak / 1e9, "frac": iach / ipeak}
            line["roofline"] = {"kernel": dom, "bound": "hbm", "achieved": achieved, "peak": peak,
                                "unit": "GB/s", "frac": achieved / peak, "traffic": traffic,
                                "peak_source": peak_src, "kernel_ms": dur,
                                "algorithmic_bytes": alg[dom], "issue": issue,
                                "note": "blend kernels are fp32-issue/L2-reduction bound (~160 FLOP "
                                        "per algorithmic byte); see DESIGN.md and profiles/"}
            line["stage_ms"] = stages
        line["clocks"] = clocks
        line["config"] = config

    if world_eff == 1 and not args.no_e2e:
        e2e = run_e2e(args, args.impl, res, device)
        if rank == 0:
            line["e2e"] = e2e
    elif world_eff > 1 and rank == 0 and res.get("e2e"):
        line["e2e"] = res["e2e"]
    if rank == 0 and args.impl == "ours" and world_eff == 1 and not args.no_cpu_baseline:
        line["cpu_baseline"] = run_cpu_baseline(args.workload)
    if rank == 0:
        print(json.dumps(line))
    if world_eff > 1:
        dist.destroy_process_group()


if __name__ == "__main__":
    main()
