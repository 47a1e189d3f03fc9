"""Kernel-only time of the backward for a shape, graph-replayed (no host gaps): python scripts/gpu_flow_time.py B D [world] [reps]"""
import os, sys, ctypes
import torch
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import crossmodal_contrastive_learning_b200 as M
from crossmodal_contrastive_learning_b200 import _native as N, loss as L

B, D = int(sys.argv[1]), int(sys.argv[2])
world = int(sys.argv[3]) if len(sys.argv) > 3 else 1
reps = int(sys.argv[4]) if len(sys.argv) > 4 else 20
ops = L._ops()
Bl = B // world
g = torch.Generator().manual_seed(0)
v = torch.randn(B, D, generator=g).to(torch.bfloat16).cuda()
t = torch.randn(B, D, generator=g).to(torch.bfloat16).cuda()
probs = [N.Problem(2 * world, Bl, D, 2 * r * Bl, 2 * Bl, 0.03, 0.8) for r in range(world)]
code, fdt, pitch = ops.plan(probs[0], v.dtype, False)
feat = torch.empty((2 * world, Bl, pitch), dtype=fdt, device="cuda")
rn = torch.empty((world, 2 * Bl), dtype=torch.float32, device="cuda")
stats = torch.empty((2 * B, 2), dtype=torch.float32, device="cuda"); coef = torch.empty_like(stats)
scal = torch.empty(4, dtype=torch.float32, device="cuda"); loss = torch.empty((), dtype=torch.float64, device="cuda")
go = torch.ones((), dtype=torch.float64, device="cuda")
dv = torch.empty((Bl, D), dtype=v.dtype, device="cuda"); dt = torch.empty_like(dv)
for r in range(world):
    ops.pack2(v[r * Bl:(r + 1) * Bl], t[r * Bl:(r + 1) * Bl], feat[2 * r:2 * r + 2], rn[r])
for r in range(world):
    ops.fwd(probs[r], code, feat, stats)
ops.finalize(probs[0], code, stats, coef, loss, scal)
lib = M.load_native()
ws_bytes = int(lib.crossclr_workspace_bytes(ctypes.byref(probs[0]), code))
ws = torch.empty(ws_bytes, dtype=torch.uint8, device="cuda")
def bwd():
    N.check(lib.crossclr_bwd(ctypes.byref(probs[0]), code, L._ptr(feat), L._ptr(rn[0]), L._ptr(coef), L._ptr(scal), L._ptr(go), 1.0,
                             L._ptr(dv), dv.stride(0), L._ptr(dt), dt.stride(0), L._DTYPE_CODE[dv.dtype], L._ptr(ws), ws_bytes, L._stream()), "bwd")
s = torch.cuda.Stream()
with torch.cuda.stream(s):
    for _ in range(3): bwd()
torch.cuda.synchronize()
if os.environ.get("CROSSCLR_FLOW_TRACE"):
    sys.exit(0)
gr = torch.cuda.CUDAGraph()
with torch.cuda.graph(gr):
    bwd()
flush = torch.empty(256 << 20, dtype=torch.uint8, device="cuda")
ts = []
for _ in range(reps):
    flush.zero_()
    a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    a.record(); gr.replay(); b.record(); torch.cuda.synchronize()
    ts.append(a.elapsed_time(b) * 1e3)
ts.sort()
name = lib.crossclr_bwd_kernel_name(ctypes.byref(probs[0]), code).decode()
alg = 8.0 * B * B * D / world
print(f"B={B} D={D} world={world} {name} env={ {k: v for k, v in os.environ.items() if k.startswith('CROSSCLR_')} }: bwd(+memsets+grad_finish) median {ts[len(ts)//2]:.1f} us min {ts[0]:.1f} us  ({alg / (ts[len(ts)//2] * 1e-6) / 1e12:.0f} TF/s alg incl. grad_finish)")
