"""Shared helpers for the parity tests (oracle side)."""
from __future__ import annotations

import ast
from pathlib import Path

import numpy as np
import torch

GOLDEN = Path(__file__).resolve().parent / "golden"
TRAIN_CASES = ["sgd_reg3", "sgd_bias_all", "sgd_noreg", "adam_all", "adam_bias_ui", "sgdm_nesterov", "rmsprop"]


def load_train_case(name: str) -> dict:
    z = np.load(GOLDEN / f"train_{name}.npz", allow_pickle=False)
    d = {k: z[k] for k in z.files}
    d["opt"] = str(d["opt"])
    d["opt_kw"] = ast.literal_eval(str(d["opt_kw"]))
    d["reg"] = ast.literal_eval(str(d["reg"]))
    d["bias"] = bool(d["bias"])
    for k in ("U", "I", "D", "B"):
        d[k] = int(d[k])
    d["coo_user"] = np.repeat(np.arange(d["U"]), np.diff(d["indptr"]))
    return d


def oracle_run(case: dict, steps: int | None = None):
    """Replay a golden case through oracle.ref_bpr with the recorded triples and negatives."""
    from oracle import ref_bpr
    bias = torch.as_tensor(case["init_item_bias"]) if case["bias"] else None
    model = ref_bpr.RefModel(torch.as_tensor(case["init_user"]), torch.as_tensor(case["init_item"]),
                             bias, case["reg"])
    opt = ref_bpr.make_optimizer(model, case["opt"].lower(), **case["opt_kw"])
    outs = []
    n = case["triples"].shape[0] if steps is None else steps
    for s in range(n):
        t = case["triples"][s]
        user = torch.as_tensor(case["coo_user"][t], dtype=torch.long)
        item = torch.as_tensor(case["indices"][t], dtype=torch.long)
        neg = torch.as_tensor(case["negs"][s], dtype=torch.long)
        outs.append(ref_bpr.train_step(model, opt, user, item, neg))
    return model, outs
