"""MaxMargin_coot (trainer/loss.py:17-41): the CPU checker against vectors generated from the
unmodified reference, and the CUDA kernels against both."""
import glob
import os
import sys

import numpy as np
import pytest
import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
GOLDEN = sorted(glob.glob(os.path.join(ROOT, "tests", "golden", "maxmargin", "maxmargin_*.npz")))


@pytest.mark.parametrize("path", GOLDEN, ids=[os.path.basename(p)[:-4] for p in GOLDEN])
def test_oracle_matches_reference_goldens(path):
    from oracle.maxmargin_oracle import maxmargin_loss_and_grads
    g = np.load(path)
    loss, dim_, ds = maxmargin_loss_and_grads(g["im"], g["s"], float(g["margin"]))
    assert abs(loss - float(g["loss"])) <= 1e-6 * abs(float(g["loss"])) + 1e-9      # the reference computes in fp32
    assert np.abs(dim_ - g["dim"]).max() <= 1e-6 * max(np.abs(g["dim"]).max(), 1e-12) + 1e-10
    assert np.abs(ds - g["ds"]).max() <= 1e-6 * max(np.abs(g["ds"]).max(), 1e-12) + 1e-10


def test_oracle_closed_forms():
    from oracle.maxmargin_oracle import maxmargin_loss_and_grads
    # im = s = I_n: scores = I, d = 1, every off-diagonal hinge is max(0, m + 0 - 1) = 0 for m < 1
    loss, a, b = maxmargin_loss_and_grads(np.eye(4), np.eye(4), 0.1)
    assert loss == 0.0 and not a.any() and not b.any()
    # all rows equal: scores_ij = d for all i, j -> every hinge = m: loss = 2 m n (n - 1) / n^2
    x = np.tile(np.array([[3.0, 4.0]]), (5, 1))
    loss, _, _ = maxmargin_loss_and_grads(x, x, 0.25)
    assert abs(loss - 2 * 0.25 * 5 * 4 / 25) < 1e-12


def test_module_surface_on_cpu():
    import crossmodal_contrastive_learning_b200 as M
    from trainer.loss import MaxMargin_coot
    assert MaxMargin_coot is M.MaxMargin_coot
    crit = MaxMargin_coot(use_cuda=True, margin=0.2)       # constructs (the reference's ctor raises NameError)
    assert crit.margin == 0.2 and crit.use_cuda is True and crit.sim is M.cosine_sim
    assert list(crit.state_dict().keys()) == []
    with pytest.raises(RuntimeError):
        crit(torch.randn(4, 8), torch.randn(4, 8))          # no CPU path


@pytest.mark.gpu
@pytest.mark.parametrize("path", GOLDEN, ids=[os.path.basename(p)[:-4] for p in GOLDEN])
def test_gpu_matches_reference_goldens(path):
    import crossmodal_contrastive_learning_b200 as M
    g = np.load(path)
    im = torch.from_numpy(g["im"]).cuda().requires_grad_()
    s = torch.from_numpy(g["s"]).cuda().requires_grad_()
    loss = M.MaxMargin_coot(True, float(g["margin"]))(im, s)
    assert loss.dim() == 0 and loss.dtype == torch.float32
    loss.backward()
    assert abs(loss.item() - float(g["loss"])) <= 1e-5 * abs(float(g["loss"])) + 1e-9
    for got, ref in ((im.grad, g["dim"]), (s.grad, g["ds"])):
        got = got.double().cpu().numpy()
        assert np.linalg.norm(got - ref) <= 1e-4 * np.linalg.norm(ref) + 1e-12


@pytest.mark.gpu
@pytest.mark.parametrize("B,D,dtype,scale", [(1000, 200, torch.float32, 0.1), (512, 512, torch.bfloat16, 0.05),
                                              (96, 77, torch.float16, 0.5), (333, 77, torch.float32, 0.2),
                                              (2048, 1024, torch.float32, 0.03)])   # fp16 gradients ~c/B^2: keep B small, they go subnormal
def test_gpu_against_oracle(B, D, dtype, scale):
    import crossmodal_contrastive_learning_b200 as M
    from oracle.maxmargin_oracle import maxmargin_loss_and_grads
    g = torch.Generator().manual_seed(B + D)
    im = (scale * torch.randn(B, D, generator=g)).to(torch.bfloat16).float()
    s = (im + scale * torch.randn(B, D, generator=g)).to(torch.bfloat16).float()
    rloss, rdim, rds = maxmargin_loss_and_grads(im.numpy(), s.numpy(), 0.1, grad_out=0.5)
    a = im.to(dtype).cuda().requires_grad_()
    b = s.to(dtype).cuda().requires_grad_()
    loss = M.MaxMargin_coot(True, 0.1)(a, b)
    assert loss.dtype == dtype
    (loss * 0.5).backward()
    # fp32 accumulation of exact bf16-representable products; a hinge within rounding of 0 may flip one 1/B^2 entry
    tol_l = 1e-5 if dtype == torch.float32 else 1e-2
    assert abs(loss.item() - rloss) <= tol_l * abs(rloss) + 1e-8
    tol_g = 1e-3 if dtype == torch.float32 else 1e-2
    for got, ref in ((a.grad, rdim), (b.grad, rds)):
        got = got.double().cpu().numpy()
        assert np.linalg.norm(got - ref) <= tol_g * np.linalg.norm(ref) + 1e-12
