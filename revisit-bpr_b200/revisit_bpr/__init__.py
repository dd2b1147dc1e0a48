"""Drop-in `revisit_bpr` package for the BPR hot path, backed by librbpr.so (sm_100a).

Only the parts of the reference library that sit on the BPR training / scoring path are mirrored
(SURVEY.md §8): `revisit_bpr.models.BPR`, `revisit_bpr.models.bpr.{Model, MF, Loss}`,
`revisit_bpr.modules.{Sampler, UniformSampler, AdaptiveSampler}` and
`revisit_bpr.metrics.{Metric, NDCG, Recall, Precision}` — same constructor signatures, argument
meaning, output keys and error types as the reference classes, so configs that name them by
`_target_` and loops written like the reference's example.py keep working.  All arithmetic runs
in hand-written CUDA kernels behind the C ABI of include/rbpr.h; there is no CPU fallback.
"""
