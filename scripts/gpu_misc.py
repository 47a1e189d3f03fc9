"""Misc GPU-box probes: pinned H2D bandwidth, launch overhead of the criterion step."""
import os, sys, time, torch
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
x = torch.empty(4096, 512, dtype=torch.bfloat16).pin_memory()
d = torch.empty_like(x, device="cuda")
for n in range(3): d.copy_(x, non_blocking=True)
torch.cuda.synchronize()
a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
a.record()
for _ in range(20): d.copy_(x, non_blocking=True)
b.record(); torch.cuda.synchronize()
print("H2D 4 MiB pinned: %.3f ms each -> %.1f GB/s" % (a.elapsed_time(b) / 20, 4.194304e-3 / (a.elapsed_time(b) / 20 * 1e-3)))
big = torch.empty(256 << 20, dtype=torch.uint8).pin_memory(); dbig = torch.empty_like(big, device="cuda")
dbig.copy_(big, non_blocking=True); torch.cuda.synchronize()
a.record(); dbig.copy_(big, non_blocking=True); b.record(); torch.cuda.synchronize()
print("H2D 256 MiB pinned: %.1f GB/s" % (0.268435456 / (a.elapsed_time(b) * 1e-3)))
import crossmodal_contrastive_learning_b200 as M
crit = M.CrossCLR_onlyIntraModality().cuda()
v = torch.randn(4096, 512, device="cuda", dtype=torch.bfloat16); t = torch.randn_like(v)
def step():
    vv = v.detach().requires_grad_(); tt = t.detach().requires_grad_()
    l = crit(vv, tt); l.backward(); return l
for _ in range(5): step()
torch.cuda.synchronize()
t0 = time.perf_counter()
for _ in range(50): step()
t1 = time.perf_counter(); torch.cuda.synchronize(); t2 = time.perf_counter()
print("host issue time per step %.1f us; wall per step %.1f us" % ((t1 - t0) / 50 * 1e6, (t2 - t0) / 50 * 1e6))
