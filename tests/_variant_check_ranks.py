"""Child process of test_gpu_variants.py: the launches of every rank of a `world`-rank job on one GPU (C ABI, row bands) under
the kernel-variant environment the parent set, compared with the CPU oracle on the concatenated batch (checker only).
Prints 'OK <loss rel> <dv rel> <dt rel>' or raises."""
import os
import sys

import numpy as np
import torch

HERE = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, os.path.dirname(HERE))
sys.path.insert(0, HERE)
import test_gpu_parity as T  # noqa: E402
from oracle import crossclr_oracle as O  # noqa: E402

B, D, world = int(sys.argv[1]), int(sys.argv[2]), int(sys.argv[3])
v, t = T._seeded(B, D, 11 + B + D + world, aligned=2.0)
rloss, rdv, rdt = O.loss_and_grads(v, t, 0.03, 0.8)
loss, dv, dt = T._run_ranks_on_one_gpu(torch.from_numpy(v).to("cuda", torch.bfloat16), torch.from_numpy(t).to("cuda", torch.bfloat16),
                                       world, 0.03, 0.8, "tc")
rel = lambda a, b: float(np.linalg.norm(a - b) / np.linalg.norm(b))
lr, dvr, dtr = abs(loss - rloss) / abs(rloss), rel(dv, rdv), rel(dt, rdt)
assert lr <= 2e-4 and dvr <= 2e-4 and dtr <= 2e-4, (lr, dvr, dtr)
print("OK", lr, dvr, dtr)
