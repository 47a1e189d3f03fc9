"""Small problems through every path (tensor-core plain / ragged / split, exact, small temperature), for
`compute-sanitizer --tool memcheck python scripts/gpu_sanitize.py`."""
import os
import sys

import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import crossmodal_contrastive_learning_b200 as M  # noqa: E402

cases = [(256, 128, torch.bfloat16, "tc", 0.03), (130, 70, torch.bfloat16, "tc", 0.03), (1, 16, torch.float32, "tc", 0.03),
         (1100, 200, torch.bfloat16, "auto", 0.03), (1024, 128, torch.float32, "auto", 0.03), (1050, 130, torch.float32, "auto", 0.03),
         (1024, 640, torch.float32, "split", 0.03), (100, 40, torch.float32, "simt", 0.03), (128, 64, torch.bfloat16, "auto", 0.005),
         (2048, 1024, torch.bfloat16, "auto", 0.03), (2100, 1000, torch.float16, "auto", 0.03)]
for B, D, dt, path, tau in ([] if "--maxmargin-only" in sys.argv else cases):
    v = torch.randn(B, D, device="cuda").to(dt).requires_grad_()
    t = torch.randn(B, D, device="cuda").to(dt).requires_grad_()
    loss = M.CrossCLR_onlyIntraModality(tau, 0.8, path=path)(v, t)
    loss.backward()
    torch.cuda.synchronize()
    print(B, D, dt, path, tau, float(loss), float(v.grad.float().norm()), flush=True)
# MaxMargin_coot / retrieval ranks on the tensor cores: rows read in place (resident and streamed row block), ragged edges, fp32
# rows staged as fp16 hi + lo pairs; `--maxmargin-only` skips the CrossCLR cases above when time on the box is short
for B, D, dt in [(384, 128, torch.bfloat16), (300, 72, torch.float16), (520, 640, torch.bfloat16), (333, 77, torch.float32),
                 (512, 576, torch.float32)]:
    a = (torch.randn(B, D, device="cuda") / D ** 0.5).to(dt).requires_grad_()
    b = (0.15 * a.detach().float() + torch.randn(B, D, device="cuda") / D ** 0.5).to(dt).requires_grad_()
    loss = M.MaxMargin_coot(True, 0.1)(a, b)
    loss.backward()
    ra, rb = M.retrieval_ranks(a.detach(), b.detach())
    torch.cuda.synchronize()
    print("maxmargin", B, D, dt, float(loss), float(a.grad.float().norm()), int(ra.sum()), int(rb.sum()), flush=True)
print("SANITIZE RUN OK")
