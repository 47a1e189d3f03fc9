"""Every tensor-core kernel variant against the oracle.  The library picks a forward / backward kernel by shape; the
CROSSCLR_*_VARIANT environment switches (read once per process, hence the child processes) force the others so that the
kernels a given shape would not select stay covered: full-Gram (non-symmetric) paired forward, the forward's super-tile and
row-interleaved orders (picked by themselves only beyond the L2's size), single-CTA slab backward,
1 S-CTA + G-CTA(s) cluster backward (also at a size the dataflow kernel would take), dataflow backward (variant 4) below
its default size threshold, and its row-band schedule (SYM_MAX = 0) in both producer orders."""
import os
import subprocess
import sys

import pytest

pytestmark = pytest.mark.gpu
HERE = os.path.dirname(os.path.abspath(__file__))


@pytest.mark.parametrize("env,B,D", [
    ({"CROSSCLR_FWD_SYM": "0"}, 512, 512), ({"CROSSCLR_FWD_SYM": "0"}, 384, 1024),
    ({"CROSSCLR_FWD_BLOCKED": "1"}, 2432, 1024), ({"CROSSCLR_FWD_BLOCKED": "1"}, 333, 77),
    ({"CROSSCLR_FWD_BLOCKED": "1", "CROSSCLR_FWD_SYM": "0"}, 1280, 512),
    ({"CROSSCLR_FWD_BLOCKED": "2"}, 1536, 256), ({"CROSSCLR_FWD_BLOCKED": "2", "CROSSCLR_FWD_SYM": "0"}, 1280, 512),
    ({"CROSSCLR_BWD_VARIANT": "1"}, 512, 512), ({"CROSSCLR_BWD_VARIANT": "2"}, 512, 512),
    ({"CROSSCLR_BWD_VARIANT": "2"}, 384, 1024), ({"CROSSCLR_BWD_VARIANT": "2"}, 2048, 512),
    ({"CROSSCLR_BWD_VARIANT": "4"}, 512, 512), ({"CROSSCLR_BWD_VARIANT": "4"}, 384, 384),
    ({"CROSSCLR_FLOW_JMAJOR": "1", "CROSSCLR_FLOW_SYM_MAX": "0"}, 2048, 512), ({"CROSSCLR_FLOW_JMAJOR": "1", "CROSSCLR_FLOW_SYM_MAX": "0"}, 1536, 1024),
    ({"CROSSCLR_FLOW_JMAJOR": "0", "CROSSCLR_FLOW_SYM_MAX": "0"}, 2048, 512),
], ids=lambda x: "-".join(f"{k[9:]}{v}" for k, v in x.items()) if isinstance(x, dict) else str(x))  # noqa: E501
def test_forced_variant_matches_oracle(env, B, D):
    e = dict(os.environ, **env)
    p = subprocess.run([sys.executable, os.path.join(HERE, "_variant_check.py"), str(B), str(D)], env=e,
                       capture_output=True, text=True, timeout=300)
    assert p.returncode == 0 and p.stdout.strip().startswith("OK"), p.stdout[-2000:] + p.stderr[-2000:]


@pytest.mark.parametrize("env,B,D,world", [
    ({"CROSSCLR_FWD_BLOCKED": "1", "CROSSCLR_FLOW_JMAJOR": "1"}, 2048, 512, 4),
    ({"CROSSCLR_FWD_BLOCKED": "2", "CROSSCLR_FLOW_JMAJOR": "1"}, 3072, 256, 3),
    ({"CROSSCLR_FWD_BLOCKED": "1", "CROSSCLR_FLOW_JMAJOR": "1"}, 2400, 1000, 2),
    ({"CROSSCLR_FWD_BLOCKED": "2", "CROSSCLR_FLOW_JMAJOR": "0"}, 4096, 512, 8),
], ids=lambda x: "-".join(f"{k[9:]}{v}" for k, v in x.items()) if isinstance(x, dict) else str(x))
def test_forced_tile_orders_on_row_bands(env, B, D, world):
    """The large-problem tile orders (forward super-tiles / row-interleaved, backward column-block-major producers) forced at
    test sizes on the row bands of a multi-rank job (every rank's launches on one GPU), ragged shapes included."""
    e = dict(os.environ, **env)
    p = subprocess.run([sys.executable, os.path.join(HERE, "_variant_check_ranks.py"), str(B), str(D), str(world)], env=e,
                       capture_output=True, text=True, timeout=300)
    assert p.returncode == 0 and p.stdout.strip().startswith("OK"), p.stdout[-2000:] + p.stderr[-2000:]
