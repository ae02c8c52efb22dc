"""Golden vectors of the hash-grid encoder from the UNMODIFIED reference extension
(oracle/_ref/grid_encoder_ext*.so, built by oracle/build_ref.build_grid_encoder) -- run on the B200
box via gpurun:  python tests/golden/make_golden_grid.py gpurun_out/golden_grid
The .npz files are then committed under tests/golden/grid/ and pin the CPU oracle
(tests/test_grid_encoder_cpu.py) without /root/reference or a GPU."""
import math
import os
import sys

import numpy as np
import torch

ROOT = os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, ROOT)
from tests import refext  # noqa: E402

# name: D, C, L, base_resolution, desired_resolution, log2_hashmap_size, gridtype, align_corners, B, oob points
CASES = {
    "hash_d5_c8": dict(D=5, C=8, L=4, H=16, desired=2048, log2T=12, gridtype="hash", align=False, B=96, oob=3),
    "hash_d3_c2_dense_levels": dict(D=3, C=2, L=6, H=4, desired=128, log2T=10, gridtype="hash", align=False, B=128, oob=4),
    "tiled_d2_c4_align": dict(D=2, C=4, L=5, H=8, desired=256, log2T=9, gridtype="tiled", align=True, B=80, oob=0),
    "hash_d4_c1": dict(D=4, C=1, L=3, H=16, desired=100, log2T=11, gridtype="hash", align=False, B=64, oob=2),
    "pow2_scale_d3_c8": dict(D=3, C=8, L=4, H=16, desired=128, log2T=14, gridtype="hash", align=False, B=64, oob=0),
}


def make_inputs(c, seed):
    g = torch.Generator().manual_seed(seed)
    x = torch.rand(c["B"], c["D"], generator=g)
    x[0] = 0.0            # the closed ends of [0, 1] are inside
    x[1] = 1.0
    for i in range(c["oob"]):     # any coordinate outside [0, 1] -> zeros, no gradient
        x[2 + i, i % c["D"]] = -0.25 if i % 2 == 0 else 1.5
    return x


def main(outdir):
    os.makedirs(outdir, exist_ok=True)
    ext = refext.load_reference_grid_ext()
    assert ext is not None, "oracle/_ref/grid_encoder_ext is not built"
    py = refext.load_reference_grid_python(ext, "ref_grid_encoder_py")
    assert py is not None, "reference Python not staged (oracle/build_ref.py)"
    dev = torch.device("cuda:0")
    for seed, (name, c) in enumerate(CASES.items()):
        torch.manual_seed(100 + seed)
        enc = py.GridEncoder(in_channels=c["D"], n_levels=c["L"], lvl_channels=c["C"],
                             desired_resolution=c["desired"], base_resolution=c["H"],
                             log2_hashmap_size=c["log2T"], gridtype=c["gridtype"],
                             align_corners=c["align"]).to(dev)
        g = torch.Generator().manual_seed(200 + seed)
        with torch.no_grad():
            enc.embeddings.copy_((torch.rand(enc.embeddings.shape, generator=g) * 2 - 1).to(dev))
        x = make_inputs(c, 300 + seed).to(dev)
        B, D, C, L = c["B"], c["D"], c["C"], c["L"]
        S = math.log2(enc.per_level_scale)
        outputs = torch.empty(L, B, C, device=dev)
        dy_dx = torch.empty(B, L * D * C, device=dev)
        ext.forward(x, enc.embeddings.data, enc.offsets, outputs, B, D, C, L, S, c["H"], True, dy_dx,
                    enc.gridtype_id, c["align"])
        grad = torch.randn(L, B, C, generator=g).to(dev)
        ge = torch.zeros_like(enc.embeddings.data)
        gi = torch.zeros(B, D, device=dev)
        ext.backward(grad, x, enc.embeddings.data, enc.offsets, ge, B, D, C, L, S, c["H"], True, dy_dx, gi,
                     enc.gridtype_id, c["align"])
        # the module-level surface too: [-1, 1] inputs through GridEncoder.forward
        xm = (x * 2 - 1).clone().requires_grad_(True)
        ym = enc(xm)
        torch.cuda.synchronize()
        # the device's own exp2f(level * S) (same CUDA libm routine the kernels call: MUFU.EX2 based), so the
        # CPU oracle can be handed the exact level scale instead of recovering it (test_grid_encoder_cpu.py)
        lv = torch.arange(L, device=dev, dtype=torch.float32) * torch.tensor(S, dtype=torch.float32, device=dev)
        e = torch.exp2(lv).cpu().numpy().astype(np.float64)
        level_scales = (e * float(c["H"]) - 1.0).astype(np.float32)
        np.savez_compressed(
            os.path.join(outdir, name + ".npz"),
            inputs=x.cpu().numpy(), embeddings=enc.embeddings.data.cpu().numpy(),
            offsets=enc.offsets.cpu().numpy(), per_level_scale=np.float64(enc.per_level_scale),
            base_resolution=np.int32(c["H"]), desired_resolution=np.int32(c["desired"]),
            log2_hashmap_size=np.int32(c["log2T"]), gridtype=np.int32(enc.gridtype_id),
            align_corners=np.int32(1 if c["align"] else 0), outputs=outputs.cpu().numpy(),
            dy_dx=dy_dx.cpu().numpy(), grad=grad.cpu().numpy(), grad_embeddings=ge.cpu().numpy(),
            grad_inputs=gi.cpu().numpy(), module_outputs=ym.detach().cpu().numpy(),
            level_scales=level_scales)
        print(name, "outputs", tuple(outputs.shape), "table rows", int(enc.offsets[-1]))


if __name__ == "__main__":
    main(sys.argv[1] if len(sys.argv) > 1 else "gpurun_out/golden_grid")
