#!/usr/bin/env python
"""Hash-grid encoder: ours vs the unmodified reference extension on one B200.

  python bench.py --workload grid_encoder [--impl ours|reference] [--steps K] [--warmup W]     (the bench leg)
  python bench_grid_encoder.py [--impl ours|reference|both] [--points 16384] [--steps 50] [--warmup 5]
                               [--dims 5] [--channels 8] [--levels 16] [--no-cpu-baseline] [--once]

This file is bench.py's `--workload grid_encoder` leg (bench.py dispatches here); like bench.py it may
time the CPU oracle as `cpu_baseline` and load the reference build from oracle/_ref as the reference arm.

Workload = the generator's positional encoder at config.py:121-123 defaults: D = ENCODER_OUT_DIM = 5
coordinates, 16 levels x 8 channels, 2^19-row fp32 tables (268 MB of embeddings -- larger than the
126 MB L2, so no flush is needed between iterations), B points per step, input gradient requested
(the coordinates come from the global encoder).  A step = forward (outputs + dy_dx) + backward
(zero-filled embedding gradient, scatter, input gradient), timed with CUDA events on the launching
stream.  Prints one JSON line per arm: Mpoints/s, per-kernel ms, roofline of the dominant kernel
against the measured HBM peak (MEASURED_PEAKS.json), cpu_baseline = the C oracle on a bounded sample.
`--once` runs each entry point once (for ncu).  Inputs identical for both arms; before timing, the
two arms are compared (outputs bit-exact, gradients <= 1e-5) and the result is part of the line.
"""
import argparse
import json
import math
import os
import sys
import time

ROOT = os.path.dirname(os.path.abspath(__file__))
if ROOT not in sys.path:
    sys.path.insert(0, ROOT)


def hbm_peak():
    p = os.path.join(ROOT, "MEASURED_PEAKS.json")
    if os.path.exists(p):
        return float(json.load(open(p))["hbm_gbs"]), "measured (MEASURED_PEAKS.json hbm_gbs)"
    return 6650.0, "fallback (B200_PROFILING.md: 6.65 TB/s)"


def run_e2e(name, torch, ge, refext, x_host, emb, D, C, L, steps, warmup, dev):
    """The same step through the public module surface with HOST inputs: pinned [B, D] coordinates in
    [-1, 1] copied H2D every step, GridEncoder.forward (maps to [0, 1], encodes), a scalar loss, backward
    to the table and the inputs, the loss read back to pinned host memory (the encoding itself feeds the
    generator's MLP on the device and is never copied out by a caller)."""
    if name == "reference":
        py = refext.load_reference_grid_python(refext.load_reference_grid_ext(), "ref_grid_py_bench")
        enc = py.GridEncoder(in_channels=D, n_levels=L, lvl_channels=C, desired_resolution=2048).to(dev)
    else:
        enc = ge.GridEncoder(in_channels=D, n_levels=L, lvl_channels=C, desired_resolution=2048,
                             fused_backward=(name == "ours_fused")).to(dev)
    with torch.no_grad():
        enc.embeddings.copy_(emb)
    B = x_host.shape[0]
    loss_host = torch.empty(1).pin_memory()
    w = torch.randn(B, L * C, device=dev)

    def step():
        x = x_host.to(dev, non_blocking=True).requires_grad_(True)
        y = enc(x)
        loss = (y * w).sum()
        enc.embeddings.grad = None
        loss.backward()
        loss_host.copy_(loss.detach().reshape(1), non_blocking=True)
    for _ in range(warmup):
        step()
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(steps):
        step()
    e1.record()
    torch.cuda.synchronize()
    ms = e0.elapsed_time(e1) / steps
    return {"value": B / ms / 1e3, "unit": "Mpoints/s", "ms_per_step": ms, "h2d_bytes_per_step": int(x_host.numel() * 4),
            "d2h_bytes_per_step": 4,
            "note": "GridEncoder module (public API), autograd fwd+bwd incl. the torch ops around the kernels"}


def main(argv=None):
    ap = argparse.ArgumentParser()
    ap.add_argument("--impl", default="both", choices=["ours", "reference", "both"])
    ap.add_argument("--points", type=int, default=16384)
    ap.add_argument("--dims", type=int, default=5)
    ap.add_argument("--channels", type=int, default=8)
    ap.add_argument("--levels", type=int, default=16)
    ap.add_argument("--steps", type=int, default=50)
    ap.add_argument("--warmup", type=int, default=5)
    ap.add_argument("--no-cpu-baseline", action="store_true")
    ap.add_argument("--once", action="store_true")
    ap.add_argument("--fused", action="store_true",
                    help="also time ours with backward_fused (no dy_dx tensor: forward without the derivative, one "
                         "backward launch that re-reads the corner rows)")
    args = ap.parse_args(argv)
    import torch
    from tests import refext
    from gaussiancity_b200 import grid_encoder as ge
    if not torch.cuda.is_available():
        print(json.dumps({"error": "no CUDA device: this benchmark has no CPU path"}))
        sys.exit(1)
    dev = torch.device("cuda:0")
    B, D, C, L, H = args.points, args.dims, args.channels, args.levels, 16
    pls = 2 ** (math.log2(2048 / H) / (L - 1))
    S = math.log2(pls)
    offsets = torch.from_numpy(ge.level_offsets(D, L)).to(dev)
    g = torch.Generator().manual_seed(0)
    emb = (torch.rand(int(offsets[-1]), C, generator=g) * 2 - 1).to(dev)
    # GaussianCity's inputs: the leading D-3 coordinates are one global feature shared by all points
    # (models/generator.py:81-83), the last three the point's relative position
    x = torch.rand(B, D, generator=g)
    if D > 3:
        x[:, :D - 3] = torch.rand(D - 3, generator=g)
    x = x.to(dev)
    grad = torch.randn(L, B, C, generator=g).to(dev)
    arms = {}
    if args.impl in ("ours", "both"):
        arms["ours"] = ge.grid_encoder_ext
    if args.impl in ("reference", "both"):
        ref = refext.load_reference_grid_ext()
        assert ref is not None, "oracle/_ref/grid_encoder_ext*.so missing"
        arms["reference"] = ref

    outputs = torch.empty(L, B, C, device=dev)
    dy_dx = torch.empty(B, L * D * C, device=dev)
    gemb = torch.empty_like(emb)
    gin = torch.empty(B, D, device=dev)

    class Fused:          # ours through gcr_grid_encode_backward_fused: same results, no [B, L, D, C] tensor
        pass
    if args.fused and "ours" in arms:
        arms["ours_fused"] = Fused

    def fwd(ext):
        if ext is Fused:
            ge.grid_encoder_ext.forward(x, emb, offsets, outputs, B, D, C, L, S, H, False, dy_dx, 0, False)
        else:
            ext.forward(x, emb, offsets, outputs, B, D, C, L, S, H, True, dy_dx, 0, False)

    def bwd(ext):
        gemb.zero_()
        gin.zero_()
        if ext is Fused:
            ge.grid_encoder_ext.backward_fused(grad, x, emb, offsets, gemb, B, D, C, L, S, H, gin, 0, False)
        else:
            ext.backward(grad, x, emb, offsets, gemb, B, D, C, L, S, H, True, dy_dx, gin, 0, False)

    results = {}
    for name, ext in arms.items():
        fwd(ext)
        bwd(ext)
        torch.cuda.synchronize()
        results[name] = (outputs.clone(), dy_dx.clone(), gemb.clone(), gin.clone())
    parity = None
    rel = lambda a, b: float((a.double() - b.double()).norm() / b.double().norm())
    fused_parity = None
    if "ours_fused" in results:
        o, f = results["ours"], results["ours_fused"]
        fused_parity = {"outputs_bit_exact": bool(torch.equal(o[0], f[0])), "grad_inputs_relerr": rel(f[3], o[3]),
                        "grad_embeddings_relerr": rel(f[2], o[2])}
    if "ours" in results and "reference" in results:
        o, r = results["ours"], results["reference"]
        parity = {"outputs_bit_exact": bool(torch.equal(o[0], r[0])), "dy_dx_bit_exact": bool(torch.equal(o[1], r[1])),
                  "grad_inputs_bit_exact": bool(torch.equal(o[3], r[3])), "grad_embeddings_relerr": rel(o[2], r[2])}
    if args.once:
        print(json.dumps({"once": True, "parity": parity}))
        return

    peak, peak_src = hbm_peak()
    ev = lambda: torch.cuda.Event(enable_timing=True)
    from bench import ClockSampler
    x_host = (x.cpu() * 2 - 1).pin_memory()
    for name, ext in arms.items():
        sampler = ClockSampler(0)
        sampler.start()
        for _ in range(args.warmup):
            fwd(ext)
            bwd(ext)
        torch.cuda.synchronize()
        t = {"fwd": 0.0, "bwd": 0.0}
        e0, e1, e2 = ev(), ev(), ev()
        marks = []
        for _ in range(args.steps):
            a, b, c = ev(), ev(), ev()
            a.record()
            fwd(ext)
            b.record()
            bwd(ext)
            c.record()
            marks.append((a, b, c))
        torch.cuda.synchronize()
        for a, b, c in marks:
            t["fwd"] += a.elapsed_time(b)
            t["bwd"] += b.elapsed_time(c)
        ms_f, ms_b = t["fwd"] / args.steps, t["bwd"] / args.steps
        # one timed region for the headline (no per-step events in between)
        e0.record()
        for _ in range(args.steps):
            fwd(ext)
            bwd(ext)
        e1.record()
        torch.cuda.synchronize()
        ms = e0.elapsed_time(e1) / args.steps
        clocks = sampler.stop()
        e2e = run_e2e(name, torch, ge, refext, x_host, emb, D, C, L, args.steps, args.warmup, dev)
        # algorithmic bytes: forward gathers 2^D rows of 4C bytes per (point, level),
        # writes 4C outputs + 4DC derivative; the backward zero-fills the gradient table, reads 4C upstream
        # gradient and reduces into 2^D rows (read-modify-write in L2: counted once, as a write), then
        # re-reads dy_dx and grad for the input gradient
        rows = B * L * (1 << D) * 4 * C
        fwd_bytes = 4 * D * B + rows + B * L * 4 * C * (1 + D)
        bwd_bytes = emb.numel() * 4 + B * L * 4 * C + rows + B * L * 4 * C * (1 + D) + 4 * D * B
        line = {
            "metric": "Mpoints/s fwd+bwd (hash-grid encoder)", "value": B / ms / 1e3, "unit": "Mpoints/s",
            "impl": name, "n_gpus": 1, "steps": args.steps, "warmup": args.warmup, "ms_per_step": ms,
            "ms_forward": ms_f, "ms_backward": ms_b, "higher_is_better": True, "dtype": "f32", "data": "synthetic",
            "config": {"workload": f"grid_encoder_d{D}_c{C}_l{L}_t19_b{B}", "points": B, "dims": D, "channels": C,
                       "levels": L, "table_rows": int(offsets[-1]), "calc_grad_inputs": True,
                       "l2": "268 MB table > 126 MB L2, no flush"},
            "roofline": {"bound": "hbm", "kernel": "forward gather (grid_fwd_kernel / kernel_grid)",
                         "achieved": fwd_bytes / (ms_f * 1e-3) / 1e9, "peak": peak, "peak_source": peak_src,
                         "unit": "GB/s", "frac": fwd_bytes / (ms_f * 1e-3) / 1e9 / peak, "traffic": None,
                         "algorithmic_bytes": fwd_bytes,
                         "backward": {"algorithmic_bytes": bwd_bytes, "achieved": bwd_bytes / (ms_b * 1e-3) / 1e9,
                                      "frac": bwd_bytes / (ms_b * 1e-3) / 1e9 / peak}},
            "parity_vs_reference": parity, "vs_baseline": None, "scaling": "weak", "clocks": clocks, "e2e": e2e,
            # forward + scatter + input-gradient kernels of this library per step (the two memsets are torch's)
            "gpu_launches": (0 if name == "reference" else (2 if name == "ours_fused" else 3)) * args.steps,
        }
        if name == "ours_fused":
            line["parity_vs_two_pass"] = fused_parity
        if name == "ours" and not args.no_cpu_baseline:
            from oracle import grid_oracle as go
            n = min(B, 2048)
            xs, es, os_ = x[:n].cpu().numpy(), emb.cpu().numpy(), offsets.cpu().numpy()
            t0 = time.perf_counter()
            _, dd = go.forward(xs, es, os_, pls, H, True, 0, False)
            go.backward(grad[:, :n].contiguous().cpu().numpy(), xs, es.shape[0], os_, pls, H, dd, 0, False)
            dt = time.perf_counter() - t0
            line["cpu_baseline"] = {"value": n / dt / 1e6, "unit": "Mpoints/s", "cores": 1, "kind": "port",
                                    "sample": f"{n} points of the same workload, C oracle, scalar ({dt:.2f} s incl. "
                                              "the 268 MB gradient-table allocation)"}
        print(json.dumps(line))


if __name__ == "__main__":
    main()
