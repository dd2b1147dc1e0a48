#!/usr/bin/env python
"""Multi-GPU correctness check of the in-library data-parallel step (run under torchrun on a box
with >= 2 GPUs):  W ranks x their share of a global batch  ==  the closed-form oracle on the whole
batch (sum-reduced loss), item replicas bit-identical across ranks, user rows only on their owner.
  python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 \
      --master-port 29533 tests/tools/check_multi_gpu.py
"""
import os
import sys
from pathlib import Path

import numpy as np
import torch
import torch.distributed as dist

ROOT = Path(__file__).resolve().parent.parent.parent
sys.path.insert(0, str(ROOT)); sys.path.insert(0, str(ROOT / "revisit-bpr_b200"))
from oracle import closed, philox  # noqa: E402  (checker only)
from oracle.ref_bpr import resolve_reg  # noqa: E402
from rbpr import native, synth  # noqa: E402
from rbpr.engine import Engine  # noqa: E402
from rbpr.parallel import owned_triples  # noqa: E402

rank, world, local = int(os.environ["RANK"]), int(os.environ["WORLD_SIZE"]), int(os.environ["LOCAL_RANK"])
torch.cuda.set_device(local)
dev = torch.device("cuda", local)
dist.init_process_group("nccl", device_id=dev)
inter = synth.generate("t", 4000, 900, 60000, 12, 4, 0.9, 7)
D, B, steps, lr, seed = 32, 8192, 4, 0.05, 321
reg = {"user": 0.0016, "item": 0.0001, "neg": 0.00375}
torch.manual_seed(1)
ue = (torch.rand(inter.num_users, D) - 0.5) * 1.2
ie = (torch.rand(inter.num_items, D) - 0.5) * 1.2
ue[0] = 0
ie[0] = 0
eng = Engine(ue.to(dev), ie.to(dev))
eng.bind_csr(torch.from_numpy(inter.indptr), torch.from_numpy(inter.indices))
eng.set_reg(reg)
eng.set_sgd(lr)
eng.set_sampler(native.SAMPLER_UNIFORM)
eng.init_comm()
lo, hi = owned_triples(inter.indptr, world, rank)
perm = torch.randperm(inter.nnz, generator=torch.Generator().manual_seed(3)).numpy()
coo = inter.coo_users()
ref_u, ref_i = ue.numpy().astype(np.float64), ie.numpy().astype(np.float64)
ok = True
for s in range(steps):
    t = perm[s * B:(s + 1) * B]
    mine = t[(t >= lo) & (t < hi)]
    stats, negs = eng.train_steps(torch.as_tensor(mine).to(dev), len(mine), seed, s, want_neg=True)
    eng.sync_check()
    tot = stats.clone()
    dist.all_reduce(tot)
    # oracle on the GLOBAL batch (negatives are a function of (seed, step, triple): rank-independent)
    exp_neg = philox.sample_negatives(inter.indptr, inter.indices, coo, t, seed, s, inter.num_items)
    mine_neg = exp_neg[(t >= lo) & (t < hi)]
    ok &= bool((negs.cpu().numpy() == mine_neg).all())
    bpr, l2, upd = closed.sgd_step(ref_u, ref_i, coo[t], inter.indices[t].astype(np.int64), exp_neg, lr, resolve_reg(reg))
    ref_u[upd["users"]] = upd["user_rows"]
    ref_i[upd["items"]] = upd["item_rows"]
    ok &= abs(tot[0, 0].item() - bpr) <= 1e-4 * abs(bpr) and tot[0, 3].item() == B
item = eng.item_emb.clone()
gathered = [torch.empty_like(item) for _ in range(world)]
dist.all_gather(gathered, item)
ok &= all(torch.equal(gathered[0], g) for g in gathered)  # replicas bit-identical
ok &= np.allclose(item.cpu().numpy(), ref_i, atol=1e-5, rtol=1e-4)
owned_users = np.unique(coo[lo:hi])
ok &= np.allclose(eng.user_emb.cpu().numpy()[owned_users], ref_u[owned_users], atol=1e-5, rtol=1e-4)
flag = torch.tensor([1 if ok else 0], device=dev)
dist.all_reduce(flag, op=dist.ReduceOp.MIN)
if rank == 0:
    print(f"multi-gpu check world={world}: {'OK' if flag.item() == 1 else 'FAILED'} (allreduces={eng.collective_count()})")
dist.destroy_process_group()
sys.exit(0 if flag.item() == 1 else 1)
