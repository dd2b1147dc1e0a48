"""`experiments.bpr.Experiment` — the name the jinja configs' `experiment._target_` points at."""
from experiments.bpr import exp as _exp

Experiment = _exp.BPRExperiment

__all__ = ["Experiment"]
