"""Forward kernel alone (crossclr_fwd: stats memset + kernel), graph-replayed: python scripts/gpu_fwd_time.py B D [world]"""
import os, sys, ctypes
import torch
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import crossmodal_contrastive_learning_b200 as M
from crossmodal_contrastive_learning_b200 import _native as N, loss as L
B, D = int(sys.argv[1]), int(sys.argv[2])
world = int(sys.argv[3]) if len(sys.argv) > 3 else 1
ops = L._ops(); lib = M.load_native()
Bl = B // world
g = torch.Generator().manual_seed(0)
v = torch.randn(B, D, generator=g).to(torch.bfloat16).cuda(); t = torch.randn(B, D, generator=g).to(torch.bfloat16).cuda()
probs = [N.Problem(2 * world, Bl, D, 2 * r * Bl, 2 * Bl, 0.03, 0.8) for r in range(world)]
code, fdt, pitch = ops.plan(probs[0], v.dtype, False)
feat = torch.empty((2 * world, Bl, pitch), dtype=fdt, device="cuda"); rn = torch.empty((world, 2 * Bl), device="cuda")
stats = torch.empty((2 * B, 2), device="cuda")
for r in range(world):
    ops.pack2(v[r * Bl:(r + 1) * Bl], t[r * Bl:(r + 1) * Bl], feat[2 * r:2 * r + 2], rn[r])
def fwd(): ops.fwd(probs[0], code, feat, stats)
s = torch.cuda.Stream()
with torch.cuda.stream(s):
    for _ in range(3): fwd()
torch.cuda.synchronize()
gr = torch.cuda.CUDAGraph()
with torch.cuda.graph(gr): fwd()
flush = torch.empty(256 << 20, dtype=torch.uint8, device="cuda")
ts = []
for _ in range(20):
    flush.zero_(); a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    a.record(); gr.replay(); b.record(); torch.cuda.synchronize(); ts.append(a.elapsed_time(b) * 1e3)
ts.sort()
print(f"fwd B={B} D={D} world={world} env={ {k: v for k, v in os.environ.items() if k.startswith('CROSSCLR_')} }: median {ts[10]:.1f} us min {ts[0]:.1f} us")
