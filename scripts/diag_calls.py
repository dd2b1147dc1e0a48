#!/usr/bin/env python
"""Wall/device time of consecutive rbpr_train_steps calls (diagnostic, not a bench line)."""
import sys, time
from pathlib import Path
import torch
ROOT = Path(__file__).resolve().parent.parent
sys.path.insert(0, str(ROOT)); sys.path.insert(0, str(ROOT / "revisit-bpr_b200"))
import bench
from rbpr import native
from rbpr.engine import Engine
dev = torch.device("cuda:0")
inter = bench.load_interactions("ml-20m", 1.0)
ue, ie = bench.init_tables(inter.num_users, inter.num_items, 128)
eng = Engine(ue.to(dev), ie.to(dev))
eng.bind_csr(torch.from_numpy(inter.indptr), torch.from_numpy(inter.indices))
eng.set_reg(bench.REG); eng.set_sgd(bench.LR); eng.set_sampler(native.SAMPLER_UNIFORM)
perm = torch.randperm(inter.nnz, generator=torch.Generator(device=dev).manual_seed(13), device=dev)
B = 65536
def call(steps, timing, stats=True, tag=""):
    eng.kernel_timing(timing)
    torch.cuda.synchronize()
    t0 = time.perf_counter()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    eng.train_steps(perm[:steps * B], B, 13, 0, want_stats=stats)
    e1.record()
    t1 = time.perf_counter()
    torch.cuda.synchronize()
    t2 = time.perf_counter()
    k = eng.kernel_time_ms() if timing else (0, 0)
    print(f"{tag:28s} steps={steps:4d} timing={timing!s:5} host-enqueue={1e3*(t1-t0):8.2f} ms total-wall={1e3*(t2-t0):8.2f} ms "
          f"device={e0.elapsed_time(e1):8.2f} ms per-step={1e3*e0.elapsed_time(e1)/steps:7.1f} us kernel={k}", flush=True)
call(5, False, tag="warm 5")
call(100, False, tag="first 100 (no timing)")
call(100, False, tag="second 100 (no timing)")
call(100, True, tag="100 with timing")
call(100, True, tag="100 with timing again")
call(100, False, stats=False, tag="100 no stats")
call(16, False, tag="16 (one wave)")
call(25, False, tag="25 (two waves)")
call(32, False, tag="32 (two waves)")
