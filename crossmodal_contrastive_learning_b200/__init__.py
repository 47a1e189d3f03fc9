"""B200-native CrossCLR criterion (drop-in for the reference's `trainer/loss.py: CrossCLR_onlyIntraModality`).

    from crossmodal_contrastive_learning_b200 import CrossCLR_onlyIntraModality
    # or, unchanged reference import path (README.md:25 of the reference):
    from trainer.loss import CrossCLR_onlyIntraModality

The compute lives in libcrossclr_b200.so (csrc/, C ABI in include/crossclr_b200.h); see DESIGN.md.
"""
from .loss import CrossCLR_onlyIntraModality, crossclr_loss  # noqa: F401
from .graph import GraphedCrossCLR, HostFedCrossCLR  # noqa: F401
from .maxmargin import MaxMargin_coot, cosine_sim  # noqa: F401
from .retrieval import recall_at_k, retrieval_metrics, retrieval_ranks  # noqa: F401
from ._native import NativeLibraryError, launch_count, load as load_native  # noqa: F401

__all__ = ["CrossCLR_onlyIntraModality", "crossclr_loss", "GraphedCrossCLR", "HostFedCrossCLR", "MaxMargin_coot", "cosine_sim", "retrieval_ranks", "recall_at_k", "retrieval_metrics", "NativeLibraryError", "launch_count", "load_native"]
