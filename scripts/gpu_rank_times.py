"""Per-rank kernel times of a `world`-rank job emulated on one GPU (each rank's launches through the C ABI, library timing
hooks): python scripts/gpu_rank_times.py B D world..."""
import os, sys
import torch
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT); sys.path.insert(0, os.path.join(ROOT, "scripts"))
import gpu_flow as G

B, D = int(sys.argv[1]), int(sys.argv[2])
for world in [int(a) for a in sys.argv[3:]]:
    g = torch.Generator().manual_seed(0)
    v = torch.randn(B, D, generator=g).to(torch.bfloat16).cuda()
    t = torch.randn(B, D, generator=g).to(torch.bfloat16).cuda()
    loss, dv, dt, kt, name = G.run_ranks(v, t, world, reps=3)
    print(f"B={B} D={D} world={world} {name}: " + "  ".join(f"{k} {ms * 1e3:.1f} us x{n}" for k, (ms, n) in kt.items()), flush=True)
