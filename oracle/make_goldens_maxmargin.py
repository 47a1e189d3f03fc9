"""Generate tests/golden/maxmargin_*.npz from the UNMODIFIED reference (test infrastructure only).

    python oracle/make_goldens_maxmargin.py          (build container: /root/reference exists)

`MaxMargin_coot` cannot be constructed as shipped (trainer/loss.py:24 calls super(ContrastiveLoss_coot, ...), a name
that does not exist), so the unbound `forward` (trainer/loss.py:29-41) is called with a stand-in `self` carrying the
three attributes it reads (`margin`, `sim`, `use_cuda=False`).  Every arithmetic op is the reference's own.
"""
import os
import sys
import types

import numpy as np
import torch

REF = os.environ.get("CROSSCLR_REFERENCE", "/root/reference")
OUT = os.path.join(os.path.dirname(os.path.abspath(__file__)), "..", "tests", "golden", "maxmargin")


def run_ref(im, s, margin):
    sys.path.insert(0, REF)
    import trainer.loss as ref
    me = types.SimpleNamespace(margin=margin, sim=ref.cosine_sim, use_cuda=False)
    a = torch.tensor(im, dtype=torch.float32, requires_grad=True)
    b = torch.tensor(s, dtype=torch.float32, requires_grad=True)
    loss = ref.MaxMargin_coot.forward(me, a, b)
    loss.backward()
    return float(loss), a.grad.double().numpy(), b.grad.double().numpy()


def main():
    for name, seed, B, D, margin, scale in [("maxmargin_b8_d4", 0, 8, 4, 0.1, 1.0), ("maxmargin_b96_d40", 1, 96, 40, 0.2, 0.3),
                                            ("maxmargin_b128_d64", 2, 128, 64, 0.1, 0.2), ("maxmargin_b1_d8", 3, 1, 8, 0.1, 1.0)]:
        torch.manual_seed(seed)
        im = (scale * torch.randn(B, D)).to(torch.bfloat16).float()
        s = (im + scale * torch.randn(B, D)).to(torch.bfloat16).float()       # bf16-representable, partly aligned
        loss, dim_, ds = run_ref(im.numpy(), s.numpy(), margin)
        np.savez_compressed(os.path.join(OUT, name + ".npz"), im=im.numpy(), s=s.numpy(), margin=margin, loss=loss,
                            dim=dim_, ds=ds)
        print(name, loss, np.linalg.norm(dim_), np.linalg.norm(ds))


if __name__ == "__main__":
    main()
