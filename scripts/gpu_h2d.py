"""H2D copy behaviour: eager vs captured in a graph, alone vs beside kernels."""
import os, sys, time
import torch
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import crossmodal_contrastive_learning_b200 as M

B, D = 4096, 512
dev = torch.device("cuda")
hv = torch.randn(B, D).to(torch.bfloat16).pin_memory()
ht = torch.randn(B, D).to(torch.bfloat16).pin_memory()
dv_ = torch.empty(B, D, dtype=torch.bfloat16, device=dev)
dt_ = torch.empty(B, D, dtype=torch.bfloat16, device=dev)

def timeit(fn, n=20):
    fn(); torch.cuda.synchronize()
    a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    a.record()
    for _ in range(n): fn()
    b.record(); torch.cuda.synchronize()
    return a.elapsed_time(b) / n * 1e3

def eager_copy():
    dv_.copy_(hv, non_blocking=True); dt_.copy_(ht, non_blocking=True)
print("eager H2D 8 MiB: %.1f us" % timeit(eager_copy))
g = torch.cuda.CUDAGraph()
with torch.cuda.graph(g):
    dv_.copy_(hv, non_blocking=True); dt_.copy_(ht, non_blocking=True)
print("graph H2D 8 MiB: %.1f us" % timeit(g.replay))
crit = M.CrossCLR_onlyIntraModality(0.03, 0.8).cuda()
res = M.HostFedCrossCLR(crit, B, D, feed="device")
print("graph compute only: %.1f us" % timeit(res.step))
pipe = M.HostFedCrossCLR(crit, B, D, feed="host")
pipe.prime()
print("graph compute || H2D (HostFed): %.1f us" % timeit(pipe.step))
# separate streams, eager copies beside graph compute
cs = torch.cuda.Stream()
def overlapped():
    cs.wait_stream(torch.cuda.current_stream())
    with torch.cuda.stream(cs):
        dv_.copy_(hv, non_blocking=True); dt_.copy_(ht, non_blocking=True)
    res.step()
    torch.cuda.current_stream().wait_stream(cs)
print("graph compute || eager H2D on a copy stream: %.1f us" % timeit(overlapped))
