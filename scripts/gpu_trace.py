import os, sys, torch
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import crossmodal_contrastive_learning_b200 as M
crit = M.CrossCLR_onlyIntraModality().cuda()
v = torch.randn(4096, 512, device="cuda", dtype=torch.bfloat16); t = torch.randn_like(v)
for _ in range(3):
    vv = v.detach().requires_grad_(); tt = t.detach().requires_grad_()
    crit(vv, tt).backward()
torch.cuda.synchronize()
