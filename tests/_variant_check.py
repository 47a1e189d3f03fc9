"""Child process of test_gpu_variants.py: one fwd+bwd of the criterion under the kernel-variant environment the parent
set, compared with the CPU oracle (checker only).  Prints 'OK <loss rel> <dv rel> <dt rel>' or raises."""
import os
import sys

import numpy as np
import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import crossmodal_contrastive_learning_b200 as M  # noqa: E402
from oracle import crossclr_oracle as O  # noqa: E402

B, D = int(sys.argv[1]), int(sys.argv[2])
g = torch.Generator().manual_seed(B + D)
v = torch.randn(B, D, generator=g).to(torch.bfloat16).float()
t = (v + 2.0 * torch.randn(B, D, generator=g)).to(torch.bfloat16).float()
rloss, rdv, rdt = O.loss_and_grads(v.numpy(), t.numpy(), 0.03, 0.8)
vd, td = v.cuda().requires_grad_(), t.cuda().requires_grad_()
loss = M.CrossCLR_onlyIntraModality(0.03, 0.8, path="tc").cuda()(vd, td)
loss.backward()
torch.cuda.synchronize()
rel = lambda a, b: float(np.linalg.norm(a - b) / np.linalg.norm(b))
lr = abs(loss.item() - rloss) / abs(rloss)
dvr, dtr = rel(vd.grad.double().cpu().numpy(), rdv), rel(td.grad.double().cpu().numpy(), rdt)
assert lr <= 1e-3 and dvr <= 1e-3 and dtr <= 1e-3, (lr, dvr, dtr)
print("OK", lr, dvr, dtr)
