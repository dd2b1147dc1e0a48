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

import torch
import torch.distributed as dist

ROOT = Path(__file__).resolve().parent.parent.parent
sys.path.insert(0, str(ROOT)); sys.path.insert(0, str(ROOT / "revisit-bpr_b200")); sys.path.insert(0, str(Path(__file__).resolve().parent))
import dp_parity  # noqa: E402  (checker only)

rank, world, local = int(os.environ["RANK"]), int(os.environ["WORLD_SIZE"]), int(os.environ["LOCAL_RANK"])
torch.cuda.set_device(local)
dev = torch.device("cuda", local)
dist.init_process_group("nccl", device_id=dev)
code = 0
binding = {}


def setup(eng):  # RBPR_FUSED_EXCHANGE=0: NCCL all-reduce + dense apply; else the fused exchange kernel
    if os.environ.get("RBPR_FUSED_EXCHANGE", "1") != "0":
        fused = eng.init_fused_exchange()
        binding["kind"] = ("multicast" if getattr(eng, "multicast", False) else
                           "symmetric unicast" if getattr(eng, "_symm_buf", None) is not None else "cudaIpc") if fused else "nccl (fused unavailable)"
    else:
        binding["kind"] = "nccl"


for opt in ("sgd", "adam"):
    ok, why = dp_parity.run(dev, rank, world, opt, setup=setup)
    if rank == 0:
        print(f"multi-gpu check world={world} optimizer={opt} exchange={binding.get('kind')}: {'OK' if ok else 'FAILED ' + why}", flush=True)
    code |= 0 if ok else 1
dist.destroy_process_group()
sys.exit(code)
