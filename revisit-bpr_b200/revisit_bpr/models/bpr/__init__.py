from revisit_bpr.models.bpr.loss import Loss
from revisit_bpr.models.bpr.model import MF, BaseLogitModel, Model
from revisit_bpr.models.bpr.knn import FreeItemKNN, ItemKNN

__all__ = ["Model", "MF", "ItemKNN", "FreeItemKNN", "BaseLogitModel", "Loss"]
