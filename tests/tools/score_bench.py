#!/usr/bin/env python
"""Config 5 of BASELINE.json: full-catalog scoring + NDCG@100 / Recall@20 on the ML-20M shape
(10 000 held-out users, 20 % of each user's items held out), our kernels vs the reference's eval
op sequence on the host cores (oracle port).  Diagnostic numbers for DESIGN.md, not the bench line.
usage: python tests/tools/score_bench.py [--users 10000] [--dim 128]"""
import argparse, json, os, sys, time
from pathlib import Path
import numpy as np
import torch
ROOT = Path(__file__).resolve().parent.parent.parent
sys.path.insert(0, str(ROOT)); sys.path.insert(0, str(ROOT / "revisit-bpr_b200"))
import bench
from rbpr import synth
from rbpr.engine import Engine

ap = argparse.ArgumentParser()
ap.add_argument("--users", type=int, default=10000)
ap.add_argument("--dim", type=int, default=128)
ap.add_argument("--cpu-batches", type=int, default=3)
args = ap.parse_args()
dev = torch.device("cuda:0")
inter = bench.load_interactions("ml-20m", 1.0)
ue, ie = bench.init_tables(inter.num_users, inter.num_items, args.dim)
ue, ie = ue * 30, ie * 30
eng = Engine(ue.to(dev), ie.to(dev))
users, seen, held = synth.split_heldout(inter, args.users)
u = torch.from_numpy(users).to(dev)
seen_d = (torch.from_numpy(seen[0]).to(dev), torch.from_numpy(seen[1]).to(dev))
held_d = (torch.from_numpy(held[0]).to(dev), torch.from_numpy(held[1]).to(dev))
res = eng.score_topk(u, seen_d, held_d, [20, 100], k_max=100, want_items=False)
torch.cuda.synchronize()
times = []
for _ in range(5):
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    res = eng.score_topk(u, seen_d, held_d, [20, 100], k_max=100, want_items=False)
    e1.record()
    torch.cuda.synchronize()
    times.append(e0.elapsed_time(e1))
ms = min(times)
flops = 2.0 * args.dim * inter.num_items * len(users)
out = {"users": len(users), "ms": ms, "users_per_s": len(users) / ms * 1e3, "tflops_fp32": flops / ms / 1e9,
       "ndcg@100": res["ndcg"][:, 1].mean().item(), "recall@20": res["recall"][:, 0].mean().item()}
# reference eval op sequence on the host (batches of 128 users, like the reference's eval loader)
from oracle import ref_bpr
torch.set_num_threads(os.cpu_count() or 1)
model = ref_bpr.RefModel(ue, ie)
t0 = time.perf_counter()
n_cpu = 0
nd, rc = [], []
for b in range(args.cpu_batches):
    rows = range(b * 128, (b + 1) * 128)
    us = torch.from_numpy(users[list(rows)])
    seen_pad = torch.nn.utils.rnn.pad_sequence(
        [torch.as_tensor(seen[1][seen[0][r]:seen[0][r + 1]], dtype=torch.long) for r in rows], batch_first=True)
    logits = model.eval_logits(us, seen_pad)
    target = ref_bpr.multi_hot(held[0][b * 128:(b + 1) * 128 + 1] - held[0][b * 128], held[1][held[0][b * 128]:held[0][(b + 1) * 128]],
                               inter.num_items)
    nd.append(ref_bpr.ndcg_at_k(logits, target, 100))
    rc.append(ref_bpr.recall_at_k(logits, target, 20))
    n_cpu += 128
cpu_s = time.perf_counter() - t0
out["cpu_users_per_s"] = n_cpu / cpu_s
out["cpu_cores"] = torch.get_num_threads()
out["max_abs_ndcg_diff_first_batches"] = float((torch.cat(nd) - res["ndcg"][:n_cpu, 1].cpu()).abs().max())
out["max_abs_recall_diff_first_batches"] = float((torch.cat(rc) - res["recall"][:n_cpu, 0].cpu()).abs().max())
print(json.dumps(out))
