#!/usr/bin/env python
"""How much of a default bench step (ML-20M shape, D=128, SGD, 262 144 triples) is attributable to
on-device negative sampling: the same 60 steps with the uniform sampler (drawn one wave ahead on the
preparation stream) and with the SAME negatives injected (no draws).  Diagnostic, not a bench line."""
import json
import sys
from pathlib import Path

import torch

ROOT = Path(__file__).resolve().parent.parent
sys.path.insert(0, str(ROOT))
sys.path.insert(0, str(ROOT / "revisit-bpr_b200"))
import bench  # noqa: E402
from rbpr import native  # noqa: E402
from rbpr.engine import Engine  # noqa: E402

dev = torch.device("cuda:0")
B, STEPS, WARM = 262144, 60, 5
inter = bench.load_interactions("ml-20m", 1.0)
ue, ie = bench.init_tables(inter.num_users, inter.num_items, 128)
perm = torch.randperm(inter.nnz, generator=torch.Generator(device=dev).manual_seed(13), device=dev)
n = min(perm.numel() // B, STEPS + WARM) * B
ids = perm[:n]
out = {}
negs = None
for mode in ("uniform", "injected", "uniform"):
    eng = Engine(ue.to(dev), ie.to(dev))
    eng.bind_csr(torch.from_numpy(inter.indptr), torch.from_numpy(inter.indices))
    eng.set_reg(bench.REG)
    eng.set_sgd(bench.LR)
    eng.set_sampler(native.SAMPLER_UNIFORM if mode == "uniform" else native.SAMPLER_INJECTED)
    if mode == "uniform" and negs is None:
        _, negs = eng.train_steps(ids, B, 13, 0, want_neg=True)  # also the untimed first pass
        eng = None
        continue
    warm = WARM * B
    eng.train_steps(ids[:warm], B, 13, 0, neg_in=None if mode == "uniform" else negs[:warm])
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    eng.train_steps(ids[warm:], B, 13, WARM, neg_in=None if mode == "uniform" else negs[warm:])
    e1.record()
    torch.cuda.synchronize()
    steps = (n - warm) // B
    out[mode] = {"us_per_step": 1e3 * e0.elapsed_time(e1) / steps, "steps": steps}
    eng.sync_check()
out["sampling_share_of_step"] = 1.0 - out["injected"]["us_per_step"] / out["uniform"]["us_per_step"]
print(json.dumps(out))
