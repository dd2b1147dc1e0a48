"""The part of the reference's `experiments` package that drives the BPR hot path:
`experiments.trainer.{Trainer, ModelEvents}` (reference experiments/trainer.py).  `ignite` and
`accelerate` are used when installed; otherwise the small stand-ins in `experiments._engine` /
`experiments._accel` provide the subset of their API the Trainer relies on (SURVEY.md §8 b)."""
