from revisit_bpr.datasets import jsonl

__all__ = ["jsonl"]
