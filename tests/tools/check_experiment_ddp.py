#!/usr/bin/env python
"""experiments.bpr.Experiment under torchrun (>= 2 GPUs): the plugin surface runs data-parallel —
owner-sharded train loaders with the same number of steps on every rank, the library's exchange
inside every step, eval users dealt round-robin with ONE (sum, count) all-reduce, user shards gathered
before eval.  Checks: item replicas bit-identical, full user table identical after the gather, metrics
identical on every rank and equal to the oracle's eval of the final tables over ALL eval users.
  python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 \
      --master-port 29561 tests/tools/check_experiment_ddp.py
"""
import os
import sys
import tempfile
from pathlib import Path

import numpy as np
import torch
import torch.distributed as dist

ROOT = Path(__file__).resolve().parent.parent.parent
for p in (ROOT, ROOT / "revisit-bpr_b200", ROOT / "tests"):
    sys.path.insert(0, str(p))
import test_gpu_experiment as T  # noqa: E402  (config text + dataset writer)

import faulthandler  # noqa: E402

rank, world, local = int(os.environ["RANK"]), int(os.environ["WORLD_SIZE"]), int(os.environ["LOCAL_RANK"])
faulthandler.dump_traceback_later(int(os.environ.get("RBPR_HANG_DUMP_S", "90")), repeat=False, exit=True)
torch.cuda.set_device(local)
dev = torch.device("cuda", local)
dist.init_process_group("nccl", device_id=dev)
from experiments._instantiate import instantiate  # noqa: E402
from oracle import ref_bpr  # noqa: E402

code = 0
for mode, optimizer, lr in (("stock", "torch.optim.SGD", 0.05), ("fast", "torch.optim.SGD", 0.05), ("stock", "torch.optim.Adam", 0.01),
                            ("fast", "torch.optim.Adam", 0.01)):
    holder = [None]
    if rank == 0:
        holder[0] = tempfile.mkdtemp()
    dist.broadcast_object_list(holder, src=0)
    tmp = Path(holder[0])
    if rank == 0:
        inter, train_rows, test_rows = T._write_dataset(tmp, n_users=400, n_items=300, seed=4)
    dist.barrier()
    if rank != 0:
        from rbpr import synth
        inter = synth.generate("t", 400, 300, 4000, 9, 4, 0.8, 4)
    cfg = T._render(tmp, num_users=inter.num_users - 1, num_items=inter.num_items - 1, epochs=2, adaptive=False,
                    train_batch_size=64, embedding_dim=16, optimizer=optimizer, lr=lr, item_bias="true")
    if mode == "fast":
        cfg["fast_train"], cfg["fast_steps_per_chunk"] = True, 5
    exp = instantiate(cfg.pop("experiment"), exp_config=lambda: cfg, dir=None, debug=False, seed=13, trackers_params={})
    print(f"[rank {rank}] {mode} {optimizer}: run()", flush=True)
    exp.run()
    print(f"[rank {rank}] {mode} {optimizer}: done", flush=True)
    faulthandler.cancel_dump_traceback_later()
    faulthandler.dump_traceback_later(int(os.environ.get("RBPR_HANG_DUMP_S", "90")), repeat=False, exit=True)
    ok, why = True, []
    sd = {k: v.detach().clone() for k, v in exp._model.state_dict().items()}
    for k, v in sd.items():
        parts = [torch.empty_like(v) for _ in range(world)]
        dist.all_gather(parts, v.contiguous())
        if not all(torch.equal(parts[0], q) for q in parts):
            ok = False
            why.append(f"{k} differs across ranks")
    eng = exp._model.logits_model.engine()
    steps = exp._model._opt_step
    tr = exp.trainer.engines["train"].state
    per = torch.tensor([steps, tr.iteration], device=dev)
    allp = [torch.empty_like(per) for _ in range(world)]
    dist.all_gather(allp, per)
    if not all(torch.equal(allp[0], q) for q in allp):
        ok = False
        why.append(f"step counts differ: {[q.tolist() for q in allp]}")
    m = torch.tensor([float(exp.metrics[k]) for k in ("ndcg@10", "recall@20", "precision@5", "auc")], device=dev, dtype=torch.float64)
    allm = [torch.empty_like(m) for _ in range(world)]
    dist.all_gather(allm, m)
    if not all(torch.equal(allm[0], q) for q in allm):
        ok = False
        why.append("reduced metrics differ across ranks")
    if rank == 0:  # oracle eval of the final tables over ALL eval users
        ref = ref_bpr.RefModel(sd["logits_model._user_emb.weight"].cpu(), sd["logits_model._item_emb.weight"].cpu(),
                               sd["logits_model._item_bias"].cpu())
        users = sorted(test_rows)
        seen_pad = torch.nn.utils.rnn.pad_sequence([torch.as_tensor(train_rows[u]) for u in users], batch_first=True)
        logits = ref.eval_logits(torch.as_tensor(users), seen_pad)
        target = torch.zeros(len(users), inter.num_items)
        for r, u in enumerate(users):
            target[r, torch.as_tensor(test_rows[u])] = 1.0
        want = [ref_bpr.ndcg_at_k(logits, target, 10).mean().item(), ref_bpr.recall_at_k(logits, target, 20).mean().item()]
        got = [float(exp.metrics["ndcg@10"]), float(exp.metrics["recall@20"])]
        if not np.allclose(got, want, atol=1e-4):
            ok = False
            why.append(f"metrics {got} vs oracle {want}")
        moved = (sd["logits_model._user_emb.weight"].cpu()[1:].abs().sum(1) > 0).float().mean().item()
        print(f"experiment ddp world={world} mode={mode} opt={optimizer.split('.')[-1]}: {'OK' if ok else 'FAILED ' + '; '.join(why)} "
              f"(steps={steps}, exchanges fused={eng.fused_exchange_count()} nccl={eng.collective_count()}, ndcg@10={got[0]:.4f})", flush=True)
    flag = torch.tensor([1 if ok else 0], device=dev)
    dist.all_reduce(flag, op=dist.ReduceOp.MIN)
    code |= 0 if flag.item() == 1 else 1
    exp.clean()
    del exp
    dist.barrier()
dist.destroy_process_group()
sys.exit(code)
