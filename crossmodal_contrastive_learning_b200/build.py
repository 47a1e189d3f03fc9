"""Build libcrossclr_b200.so in-tree with nvcc for sm_100a (no torch involved in the build).

    python -m crossmodal_contrastive_learning_b200.build [--force] [--verbose]

The library is a plain C-ABI shared object (include/crossclr_b200.h); nvcc cross-compiles it on a
machine without a GPU.  `-gencode arch=compute_100a,code=sm_100a` (not `-arch=sm_100a`): the kernels use
tcgen05 / TMA instructions that only exist for the `a` target.
"""
from __future__ import annotations

import os
import shutil
import subprocess
import sys

HERE = os.path.dirname(os.path.abspath(__file__))
ROOT = os.path.dirname(HERE)
CSRC = os.path.join(HERE, "csrc")
SOURCES = ["api.cu", "simt_kernels.cu", "tc_kernels.cu", "flow_kernels.cu", "maxmargin.cu", "maxmargin_tc.cu", "peer.cu"]
HEADERS = ["common.cuh", "tc_ptx.cuh", "tc_common.cuh", "finalize.cuh", os.path.join(ROOT, "include", "crossclr_b200.h")]
LIB = os.path.join(HERE, "libcrossclr_b200.so")


def nvcc_path() -> str:
    for cand in (os.environ.get("NVCC"), shutil.which("nvcc"), "/usr/local/cuda/bin/nvcc"):
        if cand and os.path.exists(cand):
            return cand
    raise RuntimeError("nvcc not found (set NVCC=/path/to/nvcc)")


def is_stale() -> bool:
    if not os.path.exists(LIB):
        return True
    t = os.path.getmtime(LIB)
    deps = [os.path.join(CSRC, s) for s in SOURCES] + [h if os.path.isabs(h) else os.path.join(CSRC, h) for h in HEADERS]
    return any(os.path.getmtime(d) > t for d in deps)


def build(force: bool = False, verbose: bool = False) -> str:
    """Compile the library.  Safe under several ranks of one node: an exclusive file lock serialises the builders (the
    staleness check is repeated under the lock, so only the first one compiles) and nvcc writes to a process-unique
    temporary file that is renamed into place."""
    import fcntl
    import tempfile
    if not force and not is_stale():
        return LIB
    with open(LIB + ".lock", "w") as lock:
        fcntl.flock(lock, fcntl.LOCK_EX)
        try:
            if not force and not is_stale():
                return LIB
            fd, tmp = tempfile.mkstemp(prefix=".libcrossclr_b200.", suffix=".so.tmp", dir=HERE)
            os.close(fd)
            try:
                _compile(tmp, verbose)
                os.chmod(tmp, 0o755)
                os.replace(tmp, LIB)
            finally:
                if os.path.exists(tmp):
                    os.unlink(tmp)
        finally:
            fcntl.flock(lock, fcntl.LOCK_UN)
    return LIB


def _compile(out_path: str, verbose: bool) -> None:
    cmd = [
        nvcc_path(), "-gencode", "arch=compute_100a,code=sm_100a", "-O3", "-std=c++17", "-lineinfo",
        "-shared", "-Xcompiler", "-fPIC", "-Xcompiler", "-fvisibility=hidden", "--expt-relaxed-constexpr",
        "-I", os.path.join(ROOT, "include"), "-I", CSRC,
        "-DCROSSCLR_BUILDING=1",
    ]
    if verbose:
        cmd += ["-Xptxas", "-v"]
    cmd += [os.path.join(CSRC, s) for s in SOURCES] + ["-o", out_path]
    proc = subprocess.run(cmd, capture_output=True, text=True)
    if proc.returncode != 0:
        raise RuntimeError("nvcc failed:\n" + " ".join(cmd) + "\n" + proc.stdout + proc.stderr)
    if verbose:
        print(proc.stdout + proc.stderr)


if __name__ == "__main__":
    print(build(force="--force" in sys.argv, verbose="--verbose" in sys.argv))
