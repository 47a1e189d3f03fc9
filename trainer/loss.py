"""Import-path shim: `from trainer.loss import CrossCLR_onlyIntraModality` (the reference's README usage)
resolves to the B200-native criterion.  No arithmetic lives here."""
from crossmodal_contrastive_learning_b200.loss import CrossCLR_onlyIntraModality, crossclr_loss  # noqa: F401
from crossmodal_contrastive_learning_b200.maxmargin import MaxMargin_coot, cosine_sim  # noqa: F401

__all__ = ["CrossCLR_onlyIntraModality", "crossclr_loss", "MaxMargin_coot", "cosine_sim"]
