from revisit_bpr.metrics.auc import RocAucMany, RocAucManySlow, RocAucOne
from revisit_bpr.metrics.fbeta import FBeta
from revisit_bpr.metrics.map import MAP
from revisit_bpr.metrics.metric import MaskedMetric, Metric
from revisit_bpr.metrics.ndcg import NDCG
from revisit_bpr.metrics.precision import Precision
from revisit_bpr.metrics.recall import Recall

__all__ = ["Metric", "MaskedMetric", "NDCG", "Recall", "Precision", "RocAucOne", "RocAucMany",
           "RocAucManySlow", "FBeta", "MAP"]
