#!/usr/bin/env python
"""Turn the files a final gpurun call merged into gpurun_out/ (tools/gpu_final_1gpu.sh,
tools/gpu_final_ngpu.sh) into the tracked evidence under profiles/: one bench line per workload and
arm (profiles/bench/r02_*.json), the scaling table, the config-5 table, the sanitizer summary."""
import glob
import json
import os
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
G = os.path.join(ROOT, "gpurun_out")
P = os.path.join(ROOT, "profiles")
os.makedirs(os.path.join(P, "bench"), exist_ok=True)


def last_json(path):
    try:
        lines = [l for l in open(path).read().strip().splitlines() if l.startswith("{")]
        return json.loads(lines[-1])
    except Exception:
        return None


def fmt(x, n=3):
    return "-" if x is None else f"{x:.{n}f}"


# ---- single-GPU bench lines -------------------------------------------------------------------
rows = []
for f in sorted(glob.glob(os.path.join(G, "final_bench_*.json"))):
    d = last_json(f)
    if d is None:
        continue
    name = os.path.basename(f)[len("final_bench_"):-5]
    json.dump(d, open(os.path.join(P, "bench", f"r02_{name}.json"), "w"))
    impl, wl = name.split("_", 1)
    rows.append((wl, impl, d))
if rows:
    out = ["# Round 2: one bench line per BASELINE config and arm (B200, 1 GPU)", "",
           "`bash tools/gpu_final_1gpu.sh`; full JSON lines: `profiles/bench/r02_<impl>_<workload>.json`.",
           "fwd+bwd device time (CUDA events, 30-50 steps after 5 warm-ups), forward alone, and e2e (pinned host",
           "inputs copied H2D every step, image + loss read back).", "",
           "| workload | arm | fwd+bwd ms | Msplats/s | fwd ms | e2e ms | num_rendered | SM MHz (median / max) | throttle |",
           "|---|---|---|---|---|---|---|---|---|"]
    by = {}
    for wl, impl, d in rows:
        by.setdefault(wl, {})[impl] = d
        ck = d.get("clocks") or {}
        out.append(f"| {wl} | {impl} | {fmt(d['ms_per_step'], 4)} | {fmt(d['value'], 1)} | {fmt(d.get('ms_forward'), 4)} | "
                   f"{fmt((d.get('e2e') or {}).get('ms_per_step'), 3)} | {d.get('num_rendered')} | {ck.get('sm_mhz')} / {ck.get('sm_max_mhz')} | {ck.get('reasons')} |")
    out += ["", "| workload | ours / reference (fwd+bwd) | forward | e2e |", "|---|---|---|---|"]
    for wl, v in by.items():
        if "ours" in v and "reference" in v:
            o, r = v["ours"], v["reference"]
            e = (r.get("e2e") or {}).get("ms_per_step"), (o.get("e2e") or {}).get("ms_per_step")
            out.append(f"| {wl} | {r['ms_per_step'] / o['ms_per_step']:.2f}x | {r['ms_forward'] / o['ms_forward']:.2f}x | "
                       f"{(e[0] / e[1]):.2f}x |" if all(e) else f"| {wl} | {r['ms_per_step'] / o['ms_per_step']:.2f}x | {r['ms_forward'] / o['ms_forward']:.2f}x | - |")
    o = by.get("cfg4_5M_sh3_1080p", {}).get("ours")
    if o and o.get("stage_ms"):
        out += ["", "Stage times of the headline workload (ours, ms): " + ", ".join(f"{k} {v:.3f}" for k, v in o["stage_ms"].items())]
        if o.get("roofline"):
            r = o["roofline"]
            out += ["", f"Roofline line of that run: kernel `{r['kernel']}`, {r['achieved']:.0f} GB/s algorithmic = {r['frac']:.3f} of the "
                    f"measured {r['peak']:.0f} GB/s ({r['peak_source']}); DRAM traffic per launch {r.get('traffic')}; issue: {json.dumps(r.get('issue'))}"]
        if o.get("cpu_baseline"):
            out += ["", "cpu_baseline: " + json.dumps(o["cpu_baseline"])]
    open(os.path.join(P, "r02_bench_lines.md"), "w").write("\n".join(out) + "\n")

# ---- scaling -----------------------------------------------------------------------------------
srows = []
for f in sorted(glob.glob(os.path.join(G, "final_scale_n*_*.json"))):
    d = last_json(f)
    if d is None:
        continue
    name = os.path.basename(f)[len("final_scale_"):-5]
    json.dump(d, open(os.path.join(P, "bench", f"r02_scale_{name}.json"), "w"))
    srows.append((int(d["n_gpus"]), name, d))
if srows:
    one = last_json(os.path.join(G, "final_bench_ours_cfg4_5M_sh3_1080p.json"))
    base = one["ms_per_step"] if one else None
    out = ["# Round 2: one 5 M-Gaussian / SH 3 / 1080p frame over N GPUs (B200, NVLink 5 / NVSwitch)", "",
           "`bash tools/gpu_final_ngpu.sh N` (torchrun, one rank per GPU; `bench.py --gpus N`). Every line was preceded by",
           "the bench's own gate: striped frame bit-identical to this rank's single-GPU frame, gradients <= 1e-4",
           "(`sharding.check` in the JSON). Times: max over ranks, CUDA events, 20 steps after 5 warm-ups.", "",
           f"N = 1 reference point: {fmt(base, 4)} ms.", "",
           "| N | mode | fwd+bwd ms | speed-up vs N=1 | Msplats/s | fwd ms | e2e ms | other assembly mode ms | stripe bounds | max grad rel err |",
           "|---|---|---|---|---|---|---|---|---|---|"]
    for n, name, d in sorted(srows, key=lambda t: (t[0], t[1])):
        sh = d.get("sharding") or {}
        out.append(f"| {n} | {name.split('_', 1)[1]} | {fmt(d['ms_per_step'], 4)} | {fmt(base / d['ms_per_step'], 2) if base else '-'}x | {fmt(d['value'], 0)} | "
                   f"{fmt(d.get('ms_forward'), 4)} | {fmt((d.get('e2e') or {}).get('ms_per_step'), 2)} | {fmt(sh.get('ms_other_assembly_mode'), 4)} | "
                   f"{sh.get('stripe_bounds')} | {(sh.get('check') or {}).get('max_grad_relerr')} |")
    out += ["", "Stage times (ms, max over ranks) of the `peer` lines:"]
    for n, name, d in sorted(srows, key=lambda t: (t[0], t[1])):
        if name.endswith("_peer") and d.get("stage_ms"):
            out.append(f"* N = {n}: " + ", ".join(f"{k} {v:.3f}" for k, v in d["stage_ms"].items()) + f" (sum {sum(d['stage_ms'].values()):.3f})")
    out += ["", "Modes: `peer` = default (Gaussians resident on every rank, balanced stripes, gradient sums added into the owner's",
            "accumulator over NVLink peer memory inside the blend kernel, frame left in stripes); `peer_assemble` = the same + the",
            "frame all-reduced onto every rank inside the timed step; `collective` = partial accumulators all-reduced by NCCL instead",
            "of peer memory; `broadcast` = rank 0 owns the Gaussians and NCCL-broadcasts all buffers every frame (1.18 GB, prefetched",
            "one frame ahead)."]
    open(os.path.join(P, "r02_scaling.md"), "w").write("\n".join(out) + "\n")

# ---- config 5 --------------------------------------------------------------------------------
crows = []
for f in sorted(glob.glob(os.path.join(G, "final_cfg5_*.json"))):
    d = last_json(f)
    if d is None:
        continue
    name = os.path.basename(f)[len("final_cfg5_"):-5]
    dd = dict(d)
    dd["losses"] = dd["losses"][:5] + ["..."] + dd["losses"][-5:]
    json.dump(dd, open(os.path.join(P, "bench", f"r02_cfg5_{name}.json"), "w"))
    crows.append((d["n_gpus"], d["arm"], d))
if crows:
    out = ["# Round 2: BASELINE config 5 -- the reference's own generator training step", "",
           "`tools/config5_gstep.py`: the UNMODIFIED reference `models/generator.py` (README.md:152-161 settings, PTv3 off) +",
           "`utils/helpers.py` + Adam, 16 384 synthetic lattice points, 960x540 render cropped to 640x448, L1 loss; one replica",
           "per GPU under torchrun DDP. Arms: `reference` = reference DGR Python over the reference extension; `ours` = the SAME",
           "reference DGR Python over our native module (Seam A, nothing of the caller changes); `ours_wrapper` = our wrapper",
           "(host-side camera path) handed to the reference helpers; `ours_fused` = our wrapper + `gaussiancity_b200.adapter`",
           "instead of the two helpers (no [B,N,14] tensor, crop inside the rasterizer).", "",
           "| GPUs | arm | ms / step (max over ranks) | it/s per rank | vs reference | last losses |", "|---|---|---|---|---|---|"]
    ref = {n: d["ms_per_step"] for n, a, d in crows if a == "reference"}
    for n, a, d in sorted(crows, key=lambda t: (t[0], ["reference", "ours", "ours_wrapper", "ours_fused"].index(t[1]))):
        out.append(f"| {n} | {a} | {d['ms_per_step']:.3f} | {d['it_per_s_per_rank']:.1f} | {ref[n] / d['ms_per_step']:.2f}x | "
                   f"{[round(x, 5) for x in d['losses'][-3:]]} |" if n in ref else f"| {n} | {a} | {d['ms_per_step']:.3f} | {d['it_per_s_per_rank']:.1f} | - | - |")
    l0 = {a: d["losses"] for n, a, d in crows if n == 1}
    if "reference" in l0:
        out += ["", "Loss curves over the 220 steps of the single-GPU runs (same seeds): max |loss - loss_reference| = " +
                ", ".join(f"{a}: {max(abs(x - y) for x, y in zip(l0[a], l0['reference'])):.2e}" for a in l0 if a != "reference")]
    open(os.path.join(P, "r02_config5.md"), "w").write("\n".join(out) + "\n")

# ---- sanitizer -------------------------------------------------------------------------------
san = []
for tool in ("memcheck", "racecheck", "synccheck"):
    f = os.path.join(G, f"final_{tool}.log")
    if os.path.exists(f):
        san.append(f"## {tool}\n```\n" + open(f).read().strip() + "\n```")
if san:
    open(os.path.join(P, "r02_sanitizer.md"), "w").write(
        "# Round 2: compute-sanitizer on tools/gpu_sanitize.py\n\nThree small scenes (SH 3 / colors_precomp / SH 1; forward, backward, "
        "two-stripe forward + both backward halves, pixel-window forward + backward with NULL opacity / rotation).\n\n" + "\n\n".join(san) + "\n")
print("written:", sorted(os.listdir(P)))
