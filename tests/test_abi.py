"""CPU-side checks of the C-ABI boundary: the in-tree shared library loads, exports every symbol that
include/crossclr_b200.h declares, and its argument validation / planning entry points (which never touch a
device) behave.  No compute is launched here."""
import ctypes
import os
import re

import pytest
import torch

from conftest import ROOT


def _declared_symbols():
    text = open(os.path.join(ROOT, "include", "crossclr_b200.h")).read()
    text = re.sub(r"/\*.*?\*/", "", text, flags=re.S)
    return sorted(set(re.findall(r"CROSSCLR_API\s+[\w\s\*]+?\b(crossclr_\w+)\s*\(", text)))


def test_library_exports_every_declared_symbol():
    from crossmodal_contrastive_learning_b200 import _native as N
    lib = N.load()
    names = _declared_symbols()
    assert len(names) >= 15, names
    missing = [n for n in names if not hasattr(lib, n)]
    assert not missing, missing
    assert sorted(N.EXPORTS) == names           # the ctypes binding covers exactly the header
    assert lib.crossclr_version() == 140


def test_planning_and_validation_without_a_device():
    from crossmodal_contrastive_learning_b200 import _native as N
    lib = N.load()
    P = N.Problem
    ok = P(2, 4096, 512, 0, 8192, 0.03, 0.8)
    assert lib.crossclr_choose_path(ctypes.byref(ok), N.BF16, 0) == N.PATH_TC
    assert lib.crossclr_choose_path(ctypes.byref(ok), N.F16, 0) == N.PATH_TC
    # fp32 inputs are not rounded to fp16 unasked: hi + lo operands where the dataflow backward applies, else the exact path
    assert lib.crossclr_choose_path(ctypes.byref(ok), N.F32, 0) == N.PATH_TC_SPLIT
    assert lib.crossclr_choose_path(ctypes.byref(P(2, 256, 512, 0, 512, 0.03, 0.8)), N.F32, 0) == N.PATH_SIMT
    assert lib.crossclr_choose_path(ctypes.byref(P(2, 4096, 2048, 0, 8192, 0.03, 0.8)), N.F32, 0) == N.PATH_SIMT
    assert lib.crossclr_feature_dtype(N.PATH_TC_SPLIT) == N.F16X2
    assert lib.crossclr_feature_pitch(N.PATH_TC_SPLIT, 512) == 2 * 512 + N.ROW_TAIL
    assert lib.crossclr_feature_pitch(N.PATH_TC_SPLIT, 500) == 2 * 512 + N.ROW_TAIL
    assert lib.crossclr_feature_pitch(N.PATH_TC_SPLIT, 600) == 2 * 768 + N.ROW_TAIL
    assert lib.crossclr_segment_rows(N.PATH_TC_SPLIT, 1000) == 1024
    assert lib.crossclr_choose_path(ctypes.byref(ok), N.BF16, 1) == N.PATH_SIMT
    # ragged shapes ride the tensor-core path on a zero-padded layout; tiny ones (padding would dominate) stay exact
    ragged = P(2, 100, 72, 0, 200, 0.03, 0.8)
    assert lib.crossclr_choose_path(ctypes.byref(ragged), N.BF16, 0) == N.PATH_TC
    assert lib.crossclr_segment_rows(N.PATH_TC, 100) == 128 and lib.crossclr_segment_rows(N.PATH_SIMT, 100) == 100
    assert lib.crossclr_feature_pitch(N.PATH_TC, 72) == 128 + N.ROW_TAIL and lib.crossclr_feature_pitch(N.PATH_SIMT, 72) == 72
    assert lib.crossclr_segment_rows(N.PATH_TC, 4000) == 4096 and lib.crossclr_feature_pitch(N.PATH_TC, 500) == 512 + N.ROW_TAIL
    assert lib.crossclr_workspace_bytes(ctypes.byref(ragged), N.PATH_TC) >= 2 * 256 * 128 * 4
    tiny = P(2, 8, 4, 0, 16, 0.03, 0.8)
    assert lib.crossclr_choose_path(ctypes.byref(tiny), N.BF16, 0) == N.PATH_SIMT
    # temperatures below the tensor-core kernels' range go to the exact path's online-maximum mode
    cold = P(2, 4096, 512, 0, 8192, 0.005, 0.8)
    assert lib.crossclr_choose_path(ctypes.byref(cold), N.BF16, 0) == N.PATH_SIMT
    assert lib.crossclr_fwd(ctypes.byref(cold), N.PATH_TC, ctypes.c_void_p(8), ctypes.c_void_p(8), None, 0, None) == -1
    assert b"temperature" in lib.crossclr_last_error()
    assert lib.crossclr_feature_dtype(N.PATH_TC) == N.F16 and lib.crossclr_feature_dtype(N.PATH_SIMT) == N.F32
    assert lib.crossclr_feature_pitch(N.PATH_SIMT, 512) == 512 and lib.crossclr_feature_pitch(N.PATH_TC, 512) == 512 + N.ROW_TAIL
    assert lib.crossclr_workspace_bytes(ctypes.byref(ok), N.PATH_TC) >= 8192 * 512 * 4
    assert lib.crossclr_shift(ctypes.byref(ok)) == 0.0
    small_tau = P(2, 128, 64, 0, 256, 0.0075, 0.8)
    assert abs(lib.crossclr_shift(ctypes.byref(small_tau)) - (1.4426950408889634 / 0.0075 - 96.0)) < 1e-3
    # rank 3 of 4 owns stacked rows [6 B, 8 B)
    r3 = P(8, 256, 128, 6 * 256, 512, 0.03, 0.8)
    assert lib.crossclr_choose_path(ctypes.byref(r3), N.BF16, 0) == N.PATH_TC
    for bad in (P(3, 4, 4, 0, 8, 0.03, 0.8), P(2, 0, 4, 0, 0, 0.03, 0.8), P(2, 4, 4, 4, 8, 0.03, 0.8),
                P(2, 4, 4, 0, 8, 0.0, 0.8), P(2, 4, 4, 0, 8, 0.03, float("nan"))):
        assert lib.crossclr_choose_path(ctypes.byref(bad), N.F32, 0) == -1          # CROSSCLR_EINVAL
        assert len(lib.crossclr_last_error()) > 0
    assert lib.crossclr_fwd(ctypes.byref(ok), 7, None, None, None, 0, None) == -1
    tot, n = ctypes.c_double(), ctypes.c_int64()
    assert lib.crossclr_timing_read(99, ctypes.byref(tot), ctypes.byref(n)) == -1


def test_maxmargin_and_peer_planning_without_a_device(monkeypatch):
    """The MaxMargin / retrieval path rule, its workspace size and the peer exchange's argument checks are pure host logic."""
    from crossmodal_contrastive_learning_b200 import _native as N
    lib = N.load()
    monkeypatch.delenv("CROSSCLR_MAXMARGIN_PATH", raising=False)
    name = lambda a, b, dt, sa, sb, B, D: lib.crossclr_maxmargin_kernel_name(a, b, dt, sa, sb, B, D).decode()
    assert name(0x1000, 0x2000, N.BF16, 512, 512, 4096, 512) == "mm_tc_kernel"      # 16-bit rows read in place by TMA
    assert name(0x1000, 0x2000, N.F16, 520, 512, 300, 72) == "mm_tc_kernel"
    assert name(0x1002, 0x2000, N.BF16, 512, 512, 4096, 512) == "mm_fwd_kernel"     # rows not 16-byte aligned
    assert name(0x1000, 0x2000, N.BF16, 516, 512, 4096, 512) == "mm_fwd_kernel"     # pitch not a multiple of 16 bytes
    assert name(0x1004, 0x2004, N.F32, 513, 777, 4096, 512) == "mm_tc_kernel"       # fp32 rows are staged: any alignment
    assert name(0x1000, 0x2000, N.BF16, 512, 512, 128, 512) == "mm_fwd_kernel"      # below two tiles of rows
    assert name(0x1000, 0x2000, N.BF16, 32, 32, 4096, 32) == "mm_fwd_kernel"        # below one K chunk
    monkeypatch.setenv("CROSSCLR_MAXMARGIN_PATH", "simt")
    assert name(0x1000, 0x2000, N.BF16, 512, 512, 4096, 512) == "mm_fwd_kernel"
    monkeypatch.setenv("CROSSCLR_MAXMARGIN_PATH", "tc")
    assert name(0x1000, 0x2000, N.BF16, 512, 512, 128, 512) == "invalid"            # an error, never a silent downgrade
    monkeypatch.delenv("CROSSCLR_MAXMARGIN_PATH")
    ws16 = lib.crossclr_maxmargin_workspace_bytes(4096, 512, N.BF16)
    ws32 = lib.crossclr_maxmargin_workspace_bytes(4096, 512, N.F32)
    assert ws16 >= 16 + 2 * 4096 * 4 + 4096 * 512 * 4                               # state + a direction's fp32 accumulator
    assert ws32 - ws16 == 256 + 2 * 4096 * 2 * 512 * 2                              # + the staged fp16 [hi | lo] rows
    assert lib.crossclr_maxmargin_workspace_bytes(1000, 200, N.F16) >= 1024 * 256 * 4       # padded to 128 rows / 64 columns
    assert lib.crossclr_maxmargin_workspace_bytes(0, 512, N.BF16) == 0
    # workspace too small / bad shapes are refused before anything is launched
    one = ctypes.c_void_p(0x1000)
    assert lib.crossclr_maxmargin_fwd(one, one, N.BF16, 512, 512, 4096, 512, 0.1, one, 64, one, None) == -3      # EWORKSPACE
    assert b"workspace too small" in lib.crossclr_last_error()
    assert lib.crossclr_maxmargin_fwd(one, one, N.BF16, 100, 512, 4096, 512, 0.1, one, ws16, one, None) == -1    # stride < dim
    assert lib.crossclr_retrieval_ranks(one, one, N.BF16, 512, 512, 4096, 512, one, ws16, None, one, None) == -1
    # peer exchange: collective geometry and alignment are checked on the host
    bases = (ctypes.c_void_p * 2)(0x1000, 0x2000)
    flags = (ctypes.c_void_p * 2)(0x3000, 0x4000)
    assert lib.crossclr_peer_exchange(bases, flags, 1, 0, 0, 256, 0, one, None) == -1        # a single rank exchanges nothing
    assert lib.crossclr_peer_exchange(bases, flags, 2, 2, 0, 256, 0, one, None) == -1        # rank outside the group
    assert lib.crossclr_peer_exchange(bases, flags, 2, 0, 8, 256, 0, one, None) == -1        # offset not a multiple of 16
    assert lib.crossclr_peer_exchange(bases, flags, 17, 0, 0, 256, 0, one, None) == -1       # more ranks than one node holds
    holes = (ctypes.c_void_p * 2)(0x1000, None)
    assert lib.crossclr_peer_exchange(holes, flags, 2, 0, 0, 256, 0, one, None) == -1 and b"rank 1" in lib.crossclr_last_error()
    import crossmodal_contrastive_learning_b200 as M
    with pytest.raises(ValueError):
        M.CrossCLR_onlyIntraModality(exchange="mpi")
    assert M.CrossCLR_onlyIntraModality(exchange="peer").exchange == "peer"


def test_module_surface_on_cpu():
    import crossmodal_contrastive_learning_b200 as M
    from trainer.loss import CrossCLR_onlyIntraModality
    assert CrossCLR_onlyIntraModality is M.CrossCLR_onlyIntraModality
    crit = CrossCLR_onlyIntraModality(temperature=0.07, negative_weight=0.5, logger="log")
    assert (crit.temperature, crit.negative_w, crit.logger) == (0.07, 0.5, "log")
    sd = crit.state_dict()
    assert list(sd.keys()) == ["logit_scale"] and sd["logit_scale"].dtype == torch.float32 and sd["logit_scale"].item() == 1.0
    assert [n for n, _ in crit.named_children()] == ["criterion"]
    ref_like = {"logit_scale": torch.tensor(3.0)}
    crit.load_state_dict(ref_like, strict=True)
    # no CPU path: same exception class the reference raises for CPU inputs / bad shapes (RuntimeError)
    with pytest.raises(RuntimeError):
        crit(torch.randn(4, 8), torch.randn(4, 8))
    with pytest.raises(RuntimeError):
        crit(torch.randn(4, 8), torch.randn(5, 8))
    with pytest.raises(RuntimeError):
        crit(torch.randn(2, 4, 8), torch.randn(2, 4, 8))
    with pytest.raises(ValueError):
        CrossCLR_onlyIntraModality(path="cpu")


def test_header_is_plain_c_and_the_c_host_example_links():
    """include/crossclr_b200.h compiles as C99 and examples/c_host.c links against the in-tree library (no torch, no
    Python in the signatures): the boundary a non-Python host binds."""
    import shutil
    import subprocess
    import tempfile
    root = ROOT
    gcc = shutil.which("gcc")
    cuda = "/usr/local/cuda"
    if gcc is None or not os.path.exists(os.path.join(cuda, "include", "cuda_runtime_api.h")):
        pytest.skip("gcc / CUDA headers not available")
    from crossmodal_contrastive_learning_b200 import _native as N
    N.load()                                               # make sure the library is built
    with tempfile.TemporaryDirectory() as tmp:
        for example in ("c_host", "retrieval_host"):        # the CrossCLR step; MaxMargin_coot + retrieval ranks
            cmd = [gcc, "-std=c99", "-Wall", "-Werror", "-I", os.path.join(root, "include"), "-I", os.path.join(cuda, "include"),
                   os.path.join(root, "examples", example + ".c"), "-L", os.path.join(root, "crossmodal_contrastive_learning_b200"),
                   "-lcrossclr_b200", "-L", os.path.join(cuda, "lib64"), "-lcudart", "-Wl,--allow-shlib-undefined",
                   "-o", os.path.join(tmp, example)]
            p = subprocess.run(cmd, capture_output=True, text=True)
            assert p.returncode == 0, p.stderr
