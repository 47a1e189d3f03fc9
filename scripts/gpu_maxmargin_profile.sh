#!/bin/bash
# ncu --set full capture of the MaxMargin tensor-core kernels (forward, then the two directions of the backward) at
# B = 16384, D = 512, bf16 (scripts/README.md).  Output: gpurun_out/prof_maxmargin.ncu-rep
set -u
mkdir -p gpurun_out
cat > /tmp/mm_drive.py <<'PY'
import torch
import crossmodal_contrastive_learning_b200 as M
B, D = 16384, 512
g = torch.Generator().manual_seed(B)
im = (torch.randn(B, D, generator=g) / D ** 0.5).to(torch.bfloat16)
s = (0.15 * im.float() + torch.randn(B, D, generator=g) / D ** 0.5).to(torch.bfloat16)
a, b = im.cuda().requires_grad_(), s.cuda().requires_grad_()
crit = M.MaxMargin_coot(True, 0.1)
for _ in range(4):
    a.grad = b.grad = None
    crit(a, b).backward()
torch.cuda.synchronize()
PY
PYTHONPATH=$PWD timeout 400 ncu --set full --clock-control none --import-source on -k regex:'mm_tc_kernel' -s 6 -c 3 \
    -f -o gpurun_out/prof_maxmargin python /tmp/mm_drive.py > gpurun_out/ncu_maxmargin.log 2>&1
echo "full capture rc=$?"
tail -3 gpurun_out/ncu_maxmargin.log
