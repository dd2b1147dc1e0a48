# same export names as the reference's revisit_bpr/models/__init__.py (BPR parts only)
from revisit_bpr.models.bpr import Loss as BPRLoss
from revisit_bpr.models.bpr import Model as BPR

__all__ = ["BPR", "BPRLoss"]
