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


# ---- tensor-core path (csrc/maxmargin_tc.cu) and the retrieval ranks it also serves ----
def _paired(B, D, seed, align=0.15):
    """Unit-scale rows with a partial alignment, bf16-representable: a healthy share of the hinges is active at margin 0.1."""
    g = torch.Generator().manual_seed(seed)
    im = (torch.randn(B, D, generator=g) / D ** 0.5).to(torch.bfloat16).float()
    s = (align * im + torch.randn(B, D, generator=g) / D ** 0.5).to(torch.bfloat16).float()
    return im, s


def _mm_raw(a, b, margin, grad_out=None):
    """MaxMargin forward + backward through the C ABI with fp32 gradient outputs -> (loss, d_im, d_s, kernel name)."""
    from crossmodal_contrastive_learning_b200 import _native as N
    from crossmodal_contrastive_learning_b200.loss import _DTYPE_CODE, _ptr, _stream
    lib = N.load()
    B, D = a.shape
    code = _DTYPE_CODE[a.dtype]
    name = lib.crossclr_maxmargin_kernel_name(_ptr(a), _ptr(b), code, a.stride(0), b.stride(0), B, D).decode()
    nb = int(lib.crossclr_maxmargin_workspace_bytes(B, D, code))
    ws = torch.empty(nb, dtype=torch.uint8, device=a.device)
    loss = torch.empty((), dtype=torch.float64, device=a.device)
    N.check(lib.crossclr_maxmargin_fwd(_ptr(a), _ptr(b), code, a.stride(0), b.stride(0), B, D, float(margin), _ptr(ws), nb,
                                       _ptr(loss), _stream()), "fwd")
    da = torch.empty(B, D, dtype=torch.float32, device=a.device)
    db = torch.empty(B, D, dtype=torch.float32, device=a.device)
    go = None if grad_out is None else torch.tensor(grad_out, dtype=torch.float64, device=a.device)
    N.check(lib.crossclr_maxmargin_bwd(_ptr(a), _ptr(b), code, a.stride(0), b.stride(0), B, D, float(margin), _ptr(ws), nb,
                                       None if go is None else _ptr(go), _ptr(da), D, _ptr(db), D, N.F32, _stream()), "bwd")
    torch.cuda.synchronize()
    return loss.item(), da.double().cpu().numpy(), db.double().cpu().numpy(), name


TC_SHAPES = [(512, 512, torch.bfloat16), (4096, 512, torch.bfloat16), (1000, 200, torch.float16),
             (300, 72, torch.float16), (2048, 1024, torch.bfloat16), (1100, 640, torch.bfloat16), (256, 64, torch.float16),
             # fp32 inputs: staged as fp16 hi + lo pairs, score product over K = 3 D
             (512, 512, torch.float32), (1000, 200, torch.float32), (333, 77, torch.float32), (2048, 1024, torch.float32)]


@pytest.mark.gpu
@pytest.mark.parametrize("B,D,dtype", TC_SHAPES, ids=[f"{b}x{d}-{str(t)[6:]}" for b, d, t in TC_SHAPES])
def test_tensor_core_path_against_oracle(B, D, dtype):
    from oracle.maxmargin_oracle import maxmargin_loss_and_grads
    im, s = _paired(B, D, seed=B + D)
    if dtype == torch.float16:
        im, s = im.half().float(), s.half().float()
    if dtype == torch.float32:                                 # values that need the lo halves (not 16-bit representable)
        g = torch.Generator().manual_seed(1)
        im = im * (1.0 + 1e-3 * torch.randn(B, D, generator=g))
        s = s * (1.0 + 1e-3 * torch.randn(B, D, generator=g))
    rloss, rdim, rds = maxmargin_loss_and_grads(im.numpy(), s.numpy(), 0.1, grad_out=0.5)
    assert rloss > 1e-3                                        # the case exercises active hinges
    loss, da, db, name = _mm_raw(im.to(dtype).cuda(), s.to(dtype).cuda(), 0.1, grad_out=0.5)
    assert name == "mm_tc_kernel"
    # exact products, fp32 accumulation: the loss to rounding; a hinge within rounding of 0 may flip one 1/B^2 entry.  fp32
    # inputs: hi + lo products are 22-bit and the tensor cores' accumulation truncates them, a one-sided 2^-22-of-|a||b| error per
    # score that does not average out over the hinges (measured 3e-6 of the loss)
    assert abs(loss - rloss) <= (1e-5 if dtype == torch.float32 else 2e-6) * abs(rloss)
    for got, ref in ((da, rdim), (db, rds)):
        assert np.linalg.norm(got - ref) <= 1e-3 * np.linalg.norm(ref)


@pytest.mark.gpu
def test_tensor_core_and_cuda_core_paths_agree(monkeypatch):
    im, s = _paired(768, 256, seed=5)
    a, b = im.bfloat16().cuda(), s.bfloat16().cuda()
    monkeypatch.setenv("CROSSCLR_MAXMARGIN_PATH", "simt")
    l0, da0, db0, n0 = _mm_raw(a, b, 0.2)
    monkeypatch.setenv("CROSSCLR_MAXMARGIN_PATH", "tc")
    l1, da1, db1, n1 = _mm_raw(a, b, 0.2)
    assert (n0, n1) == ("mm_fwd_kernel", "mm_tc_kernel")
    assert abs(l0 - l1) <= 2e-6 * abs(l0)
    assert np.linalg.norm(da0 - da1) <= 1e-3 * np.linalg.norm(da0) and np.linalg.norm(db0 - db1) <= 1e-3 * np.linalg.norm(db0)
    # fp32 inputs: hi + lo operands on the tensor cores against the exact CUDA-core kernels
    l3, da3, db3, n3 = _mm_raw(im.cuda(), s.cuda(), 0.2)
    monkeypatch.setenv("CROSSCLR_MAXMARGIN_PATH", "simt")
    l2, da2, db2, n2 = _mm_raw(im.cuda(), s.cuda(), 0.2)
    assert (n2, n3) == ("mm_fwd_kernel", "mm_tc_kernel")
    assert abs(l2 - l3) <= 1e-5 * abs(l2) and np.linalg.norm(da2 - da3) <= 1e-3 * np.linalg.norm(da2)
    # asking for the tensor cores on a problem they do not take is an error, not a silent downgrade
    monkeypatch.setenv("CROSSCLR_MAXMARGIN_PATH", "tc")
    with pytest.raises(RuntimeError, match="tensor-core kernels need"):
        _mm_raw(a[:100], b[:100], 0.2)
    monkeypatch.delenv("CROSSCLR_MAXMARGIN_PATH")
    # a row pitch that breaks the 16-byte alignment of the TMA boxes falls back to the CUDA cores (16-bit inputs are not staged)
    wide = torch.zeros(768, 260, dtype=torch.bfloat16, device="cuda")
    wide[:, :256] = a
    assert _mm_raw(wide[:, :256], b, 0.2)[3] == "mm_fwd_kernel"
    # fp32 rows are staged, so any pitch rides the tensor cores
    widef = torch.zeros(768, 259, dtype=torch.float32, device="cuda")
    widef[:, :256] = im.cuda()
    l4, da4, _, n4 = _mm_raw(widef[:, :256], s.cuda(), 0.2)
    # (partial row-block sums meet through fp32 atomics: equal to rounding, not bit for bit)
    assert n4 == "mm_tc_kernel" and abs(l4 - l3) <= 1e-9 * abs(l3) and np.linalg.norm(da4 - da3) <= 1e-5 * np.linalg.norm(da3)


@pytest.mark.gpu
@pytest.mark.parametrize("scale", [300.0, 1.0 / 64, 1e-6])
def test_tensor_core_fp32_inputs_at_any_magnitude(scale):
    """The staged fp16 pairs carry a power-of-two scale per tensor: unnormalised embeddings far outside the fp16 range (or deep
    inside its subnormals) keep fp32-grade scores."""
    from oracle.maxmargin_oracle import maxmargin_loss_and_grads
    im, s = _paired(640, 192, seed=11)
    g = torch.Generator().manual_seed(2)
    im = im * (1.0 + 1e-3 * torch.randn(640, 192, generator=g)) * scale
    s = s * (1.0 + 1e-3 * torch.randn(640, 192, generator=g)) * (scale * 4.0)
    m = 0.1 * scale * scale * 4.0
    rloss, rdim, rds = maxmargin_loss_and_grads(im.numpy(), s.numpy(), m)
    loss, da, db, name = _mm_raw(im.cuda(), s.cuda(), m)
    assert name == "mm_tc_kernel"
    assert abs(loss - rloss) <= 1e-5 * abs(rloss)
    assert np.linalg.norm(da - rdim) <= 1e-3 * np.linalg.norm(rdim) and np.linalg.norm(db - rds) <= 1e-3 * np.linalg.norm(rds)


@pytest.mark.gpu
def test_module_on_tensor_cores_matches_closed_forms():
    import crossmodal_contrastive_learning_b200 as M
    # all rows equal: every score equals the diagonal, every hinge = m: loss = 2 m (n - 1) / n
    n = 512
    x = torch.tensor([[0.5, -0.25] * 64], dtype=torch.bfloat16).repeat(n, 1).cuda().requires_grad_()
    loss = M.MaxMargin_coot(True, 0.25)(x, x.detach().clone())
    assert abs(loss.float().item() - 2 * 0.25 * (n - 1) / n) < 4e-3        # bf16 result
    # orthogonal pairs far apart: no active hinge, zero loss and zero gradient
    e = torch.eye(512, dtype=torch.float16, device="cuda").requires_grad_()
    loss = M.MaxMargin_coot(True, 0.1)(e, e.detach().clone())
    loss.backward()
    assert loss.item() == 0.0 and not e.grad.any()


def test_retrieval_oracle_closed_forms():
    from oracle.retrieval_oracle import recall_at_k, retrieval_ranks
    ra, rb, _ = retrieval_ranks(np.eye(5), np.eye(5))
    assert not ra.any() and not rb.any() and recall_at_k(ra)[1] == 1.0
    # s = im rolled by one row: every query's partner scores 0, exactly one candidate scores 1
    ra, rb, _ = retrieval_ranks(np.eye(5), np.roll(np.eye(5), 1, axis=0))
    assert (ra == 1).all() and (rb == 1).all() and recall_at_k(ra, (1, 2)) == {1: 0.0, 2: 1.0}
    # ties do not count against the partner
    x = np.ones((4, 3))
    ra, rb, _ = retrieval_ranks(x, x)
    assert not ra.any() and not rb.any()


@pytest.mark.gpu
@pytest.mark.parametrize("B,D,dtype", [(2048, 512, torch.bfloat16), (1000, 200, torch.float16), (333, 77, torch.float32)])
def test_retrieval_ranks_against_oracle(B, D, dtype):
    import crossmodal_contrastive_learning_b200 as M
    from oracle.retrieval_oracle import recall_at_k, retrieval_ranks
    im, s = _paired(B, D, seed=B, align=0.12)
    if dtype == torch.float16:
        im, s = im.half().float(), s.half().float()
    ra, rb, (gap_a, gap_b) = retrieval_ranks(im.numpy(), s.numpy())
    assert 0.02 < recall_at_k(ra)[10] < 0.98                    # a non-trivial ranking problem
    ga, gb = M.retrieval_ranks(im.to(dtype).cuda(), s.to(dtype).cuda())
    assert ga.dtype == torch.int32 and ga.shape == (B,)
    for got, ref, gap in ((ga, ra, gap_a), (gb, rb, gap_b)):
        got = got.cpu().numpy().astype(np.int64)
        bad = got != ref
        # a rank may differ by one only where some candidate ties with the partner to within fp32 accumulation noise
        assert np.abs(got - ref).max() <= 1 and (gap[bad] < 1e-6).all()
    m = M.retrieval_metrics(im.to(dtype).cuda(), s.to(dtype).cuda())
    assert abs(m["im2s_R@10"] - recall_at_k(ra)[10]) <= 2.0 / B and abs(m["s2im_R@1"] - recall_at_k(rb)[1]) <= 2.0 / B
    assert m["im2s_MedR"] >= 1.0


# ---- CPU: properties of the checkers and of the staged fp32 representation (no GPU) ----------------------------------------
def test_oracle_symmetry_and_rank_definition():
    """Swapping the two blocks transposes the score matrix: same loss, gradients exchanged, rank arrays exchanged; the ranks
    equal the position of the partner in a stable descending sort when there are no ties."""
    from oracle.maxmargin_oracle import maxmargin_loss_and_grads
    from oracle.retrieval_oracle import retrieval_ranks
    rng = np.random.default_rng(3)
    im, s = rng.standard_normal((37, 11)), rng.standard_normal((37, 11))
    l0, da0, db0 = maxmargin_loss_and_grads(im, s, 0.3)
    l1, da1, db1 = maxmargin_loss_and_grads(s, im, 0.3)
    assert abs(l0 - l1) < 1e-12 and np.allclose(da0, db1) and np.allclose(db0, da1)
    ra, rb, _ = retrieval_ranks(im, s)
    rb2, ra2, _ = retrieval_ranks(s, im)
    assert (ra == ra2).all() and (rb == rb2).all()
    scores = im @ s.T
    order = np.argsort(-scores, axis=1, kind="stable")
    assert (ra == np.array([int(np.where(order[i] == i)[0][0]) for i in range(37)])).all()
    # the hinge indicators at margin 0 are the rank indicators (what lets one kernel serve both): with orthonormal s the diagonal
    # coefficient of the gradient, -(active hinges in row i + in column i) / B^2, reads off as (dL/dim_i . s_i)
    q, _ = np.linalg.qr(rng.standard_normal((16, 16)))
    im2 = rng.standard_normal((16, 16))
    _, da, _ = maxmargin_loss_and_grads(im2, q, 0.0)
    ra2, rb2, _ = retrieval_ranks(im2, q)
    coef = da @ q.T * 16 * 16                      # G (rows of s orthonormal): off-diagonal indicators, diagonal -(counts)
    assert np.allclose(-np.diag(coef), ra2 + rb2)


@pytest.mark.parametrize("scale", [1.0, 300.0, 1e-6, 3e4])
def test_staged_fp16_pairs_keep_fp32_grade_scores(scale):
    """numpy restatement of mm_absmax_kernel / mm_split_kernel (csrc/maxmargin_tc.cu): x 2^-e with the tensor's largest magnitude
    in [2^13, 2^14), hi = fp16, lo = fp16(rest); score = (hi.hi + lo.hi + hi.lo) 2^(e_a + e_b).  The representation error of a
    score stays ~2^-22 of |a||b| whatever the inputs' magnitude -- without the scale fp16 would overflow or flush."""
    rng = np.random.default_rng(5)
    a = (rng.standard_normal((64, 256)) * scale).astype(np.float32)
    b = (rng.standard_normal((64, 256)) * scale * 4).astype(np.float32)

    def stage(x):
        mx = np.abs(x).max()
        e = int(np.floor(np.log2(mx))) - 13
        e = max(-126, min(126, e))
        xs = (x.astype(np.float64) * 2.0 ** -e).astype(np.float32)
        assert 2.0 ** 13 <= np.abs(xs).max() < 2.0 ** 14
        hi = xs.astype(np.float16)
        lo = (xs - hi.astype(np.float32)).astype(np.float16)
        assert np.isfinite(hi).all()
        return hi.astype(np.float64), lo.astype(np.float64), 2.0 ** e

    ha, la, ca = stage(a)
    hb, lb, cb = stage(b)
    staged = (ha @ hb.T + la @ hb.T + ha @ lb.T) * ca * cb
    exact = a.astype(np.float64) @ b.astype(np.float64).T
    bound = np.linalg.norm(a.astype(np.float64), axis=1)[:, None] * np.linalg.norm(b.astype(np.float64), axis=1)[None, :]
    assert (np.abs(staged - exact) / bound).max() < 2.0 ** -21
    # a single fp16 rounding of the same rows is ~2^-11 per element: three orders of magnitude worse on the scores
    single = (ha @ hb.T) * ca * cb
    assert (np.abs(single - exact) / bound).max() > 50 * (np.abs(staged - exact) / bound).max()
