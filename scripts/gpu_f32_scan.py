"""GPU-box scan: gradient error of the tensor-core path on fp32 (NOT 16-bit-representable) inputs against the oracle, by
temperature -- the measurement behind crossclr_choose_path's rule for fp32 inputs.

    python scripts/gpu_f32_scan.py
"""
import os
import sys

import numpy as np
import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import crossmodal_contrastive_learning_b200 as M  # noqa: E402
from oracle import crossclr_oracle as O  # noqa: E402


def main():
    for B, D, al in ((2048, 512, 2.0), (1024, 128, 1.0), (4096, 512, 0.0)):
        g = torch.Generator().manual_seed(B + D)
        v = torch.randn(B, D, generator=g)
        t = v + al * torch.randn(B, D, generator=g) if al else torch.randn(B, D, generator=g)
        for tau in (0.07, 0.03, 0.02, 0.015, 0.01, 0.0075):
            rl, rdv, rdt = O.loss_and_grads(v.numpy(), t.numpy(), tau, 0.8)
            out = []
            for path in ("tc", "auto", "simt"):
                vd, td = v.cuda().requires_grad_(), t.cuda().requires_grad_()
                loss = M.CrossCLR_onlyIntraModality(tau, 0.8, path=path)(vd, td)
                loss.backward()
                dv, dt = vd.grad.double().cpu().numpy(), td.grad.double().cpu().numpy()
                ev = np.linalg.norm(dv - rdv) / np.linalg.norm(rdv)
                et = np.linalg.norm(dt - rdt) / np.linalg.norm(rdt)
                mv = np.abs(dv - rdv).max() / np.abs(rdv).max()
                out.append(f"{path}: loss {abs(loss.item() - rl) / abs(rl):.1e} dv {ev:.2e} dt {et:.2e} max {mv:.2e}")
            print(f"B={B} D={D} aligned={al} tau={tau}: " + " | ".join(out), flush=True)


if __name__ == "__main__":
    main()
