from revisit_bpr.models.bpr.loss import Loss
from revisit_bpr.models.bpr.model import MF, BaseLogitModel, Model

__all__ = ["Model", "MF", "BaseLogitModel", "Loss"]
