#!/usr/bin/env python
"""GPU tuning sweep of the training path (not a bench line): per-step device time for
a list of batch sizes.
usage: python scripts/sweep.py [--dim 128] [--shape ml-20m]"""
import argparse, os, sys, time
from pathlib import Path
import torch
ROOT = Path(__file__).resolve().parent.parent
sys.path.insert(0, str(ROOT)); sys.path.insert(0, str(ROOT / "revisit-bpr_b200"))
import bench
from rbpr import native
from rbpr.engine import Engine

ap = argparse.ArgumentParser()
ap.add_argument("--dim", type=int, default=128)
ap.add_argument("--shape", default="ml-20m")
ap.add_argument("--opt", default="sgd")
ap.add_argument("--chunks", default="0", help="unused (kept for old command lines)")
ap.add_argument("--batches", default="256,4096,65536,262144")
ap.add_argument("--zipf", type=float, default=None, help="override the item-popularity exponent")
args = ap.parse_args()
dev = torch.device("cuda:0")
if args.zipf is None:
    inter = bench.load_interactions(args.shape, 1.0)
else:
    from rbpr import synth
    u0, i0, nnz0, med, mind, _ = synth.SHAPES[args.shape]
    inter = synth.generate(args.shape, u0, i0, nnz0, med, mind, args.zipf, 13)
ue, ie = bench.init_tables(inter.num_users, inter.num_items, args.dim)
eng = Engine(ue.to(dev), ie.to(dev))
eng.bind_csr(torch.from_numpy(inter.indptr), torch.from_numpy(inter.indices))
eng.set_reg(bench.REG)
if args.opt == "sgd":
    eng.set_sgd(bench.LR)
else:
    eng.set_adam(1e-3)
eng.set_sampler(native.SAMPLER_UNIFORM)
g = torch.Generator(device=dev).manual_seed(13)
perm = torch.randperm(inter.nnz, generator=g, device=dev)

def timeit(B, steps, chunk, reps=3):
    n = min(B * steps, inter.nnz)
    steps = n // B
    t = perm[:steps * B]
    eng.train_steps(t, B, 13, 0, want_stats=False)
    torch.cuda.synchronize()
    best = 1e9
    for r in range(reps):
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        eng.train_steps(t, B, 13, (r + 1) * steps, want_stats=False)
        e1.record()
        torch.cuda.synchronize()
        best = min(best, e0.elapsed_time(e1))
    return best * 1e3 / steps, steps

print(f"shape={args.shape} D={args.dim} opt={args.opt} zipf={args.zipf}")
for B in [int(b) for b in args.batches.split(",")]:
    for chunk in [int(c) for c in args.chunks.split(",")]:
        steps = max(8, min(2048, (1 << 23) // B))
        us, st = timeit(B, steps, chunk)
        print(f"B={B:7d} steps/call={st:5d} chunk={chunk or 'auto':>4} : {us:9.2f} us/step  {B / us:8.1f} Mtriples/s  "
              f"{B * 24 * args.dim / us / 1e3:8.1f} GB/s alg", flush=True)
eng.sync_check()
