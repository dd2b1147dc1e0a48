"""Ranking metrics of the reference library, evaluated by the CUDA top-k / AUC kernels."""
from revisit_bpr.metrics import auc as _auc
from revisit_bpr.metrics import fbeta as _fbeta
from revisit_bpr.metrics import map as _map
from revisit_bpr.metrics import metric as _metric
from revisit_bpr.metrics import ndcg as _ndcg
from revisit_bpr.metrics import precision as _precision
from revisit_bpr.metrics import recall as _recall

Metric, MaskedMetric = _metric.Metric, _metric.MaskedMetric
NDCG, Recall, Precision, MAP, FBeta = _ndcg.NDCG, _recall.Recall, _precision.Precision, _map.MAP, _fbeta.FBeta
RocAucOne, RocAucMany, RocAucManySlow = _auc.RocAucOne, _auc.RocAucMany, _auc.RocAucManySlow

__all__ = ["Metric", "MaskedMetric", "NDCG", "Recall", "Precision", "MAP", "FBeta", "RocAucOne", "RocAucMany",
           "RocAucManySlow"]
