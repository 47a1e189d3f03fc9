"""MaxMargin_coot / retrieval ranks on one GPU: the `next_rows` record of bench.py alone (scripts/README.md)."""
import json
import os
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import torch  # noqa: E402

import bench  # noqa: E402
import crossmodal_contrastive_learning_b200 as M  # noqa: E402
from crossmodal_contrastive_learning_b200 import _native as NAT  # noqa: E402

dev = torch.device("cuda:0")
flush = torch.empty(256 << 20, dtype=torch.uint8, device=dev)
print(json.dumps(bench.run_next_rows(M, NAT, torch, dev, flush), indent=1))
