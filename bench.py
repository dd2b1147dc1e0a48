#!/usr/bin/env python
"""Benchmark of the B200-native BPR training hot path (BASELINE.json metric: BPR triples/sec at
dim=128 on a synthetic ML-20M-shape matrix; achieved HBM GB/s vs peak).

  python bench.py [--gpus N] [--steps K] [--warmup W]          our arm (CUDA, sm_100a)
  python bench.py --impl reference [...]                       the reference's CPU path (oracle port)

One "step" = one minibatch of --batch triples through the fused path: on-device negative
sampling, (u,i+,i-) gather, loss, exact minibatch gradients, SGD update.  Prints ONE JSON line.
Under torchrun (N>1) users are sharded by owner, the item table is replicated and the dense item
gradient is all-reduced once per step (NCCL); `value` is the whole-job aggregate.
"""
from __future__ import annotations

import argparse
import json
import os
import sys
import threading
import time
from pathlib import Path

import numpy as np
import torch

ROOT = Path(__file__).resolve().parent
sys.path.insert(0, str(ROOT / "revisit-bpr_b200"))
sys.path.insert(0, str(ROOT))

METRIC = "BPR triples/sec at dim=128 ML-20M shape"
UNIT = "triples/s"
REG = {"user": 0.0016, "item": 0.0001, "neg": 0.00375}  # configs/RQ2/neg-sampling/ada-sampling-ml-20m.yaml.j2:144-147
LR = 0.001
SEED = 13


def parse():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=60)
    ap.add_argument("--warmup", type=int, default=5)
    ap.add_argument("--impl", default="ours", choices=["ours", "reference"])
    ap.add_argument("--shape", default="ml-20m")
    ap.add_argument("--dim", type=int, default=128)
    ap.add_argument("--sampler", default="uniform", choices=["uniform", "adaptive"],
                    help="adaptive = BASELINE configs[3] (Yelp shape, dim 64): sampling_prob 1/100, statistics "
                         "refreshed every int(I ln I / batch) steps")
    ap.add_argument("--opt", default="sgd", choices=["sgd", "adam"],
                    help="adam = BASELINE configs[2] (MSD shape, dim 256): lr 1e-3, betas (0.9,0.999), reg all=0.00043")
    ap.add_argument("--batch", type=int, default=262144,
                    help="triples per step and per GPU (train_batch_size is a free jinja variable of the reference configs)")
    ap.add_argument("--scale", type=float, default=1.0, help="shrink the synthetic matrix (tests)")
    ap.add_argument("--no-cpu-baseline", action="store_true")
    ap.add_argument("--clock-period-ms", type=float, default=5.0, help="NVML sampling period (0 = off)")
    ap.add_argument("--cpu-steps", type=int, default=20)
    ap.add_argument("--e2e-steps-per-call", type=int, default=30,
                    help="steps handed to one host-buffer API call in the e2e leg")
    return ap.parse_args()


def load_interactions(shape: str, scale: float):
    from rbpr import synth
    cache = Path(os.environ.get("RBPR_CACHE", "/tmp")) / f"rbpr_synth_{shape}_{scale}_{SEED}.npz"
    if cache.exists():
        z = np.load(cache)
        return synth.Interactions(shape, int(z["U"]), int(z["I"]), z["indptr"], z["indices"])
    inter = synth.make(shape, seed=SEED, scale=scale)
    try:
        tmp = cache.with_suffix(f".{os.getpid()}.tmp.npz")
        np.savez(tmp, U=inter.num_users, I=inter.num_items, indptr=inter.indptr, indices=inter.indices)
        os.replace(tmp, cache)
    except OSError:
        pass
    return inter


def init_tables(U: int, I: int, D: int):
    """MF.reset_parameters (revisit_bpr/models/bpr/model.py:117-129) under torch.manual_seed(13)."""
    torch.manual_seed(SEED)
    ue = (torch.rand(U, D) - 0.5) / D
    ie = (torch.rand(I, D) - 0.5) / D
    ue[0] = 0
    ie[0] = 0
    return ue, ie


class ClockSampler:
    """SM clock and throttle reasons sampled DURING the timed region (B200_PROFILING.md clocks line).
    In-process NVML thread (light queries every 5 ms): a polling `nvidia-smi -lms` child process
    was seen to stall the CUDA driver for hundreds of ms at a time on these boxes."""
    NAMES = ["hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown", "sw_power_cap"]

    def __init__(self, index: int, period_ms: float = 5.0):
        self.rows: list[tuple[float, float, int]] = []
        self.h = None
        self._stop = False
        self.period = period_ms * 1e-3
        try:
            if period_ms <= 0:
                raise RuntimeError("sampling disabled")
            import pynvml
            self.nv = pynvml
            pynvml.nvmlInit()
            uuid = str(torch.cuda.get_device_properties(index).uuid)
            try:
                self.h = pynvml.nvmlDeviceGetHandleByUUID(("GPU-" + uuid).encode())
            except Exception:
                self.h = pynvml.nvmlDeviceGetHandleByIndex(index)
            self.max_sm = float(pynvml.nvmlDeviceGetMaxClockInfo(self.h, pynvml.NVML_CLOCK_SM))
            self.bits = [pynvml.nvmlClocksEventReasonHwSlowdown, pynvml.nvmlClocksEventReasonHwThermalSlowdown,
                         pynvml.nvmlClocksEventReasonSwThermalSlowdown, pynvml.nvmlClocksEventReasonSwPowerCap]
            self.th = threading.Thread(target=self._run, daemon=True)
            self.th.start()
        except Exception as e:  # noqa: BLE001
            self.h = None
            self.err = repr(e)

    def _run(self):
        nv = self.nv
        while not self._stop:
            try:
                sm = float(nv.nvmlDeviceGetClockInfo(self.h, nv.NVML_CLOCK_SM))
                reasons = int(nv.nvmlDeviceGetCurrentClocksEventReasons(self.h))
                self.rows.append((time.perf_counter(), sm, reasons))
            except Exception:  # noqa: BLE001
                pass
            time.sleep(self.period)

    def mark(self):
        return len(self.rows)

    def stop(self, start_row: int = 0) -> dict:
        if self.h is None:
            return {"sm_mhz": None, "sm_max_mhz": None, "reasons": [f"nvml unavailable: {getattr(self, 'err', '')}"]}
        self._stop = True
        self.th.join(timeout=1.0)
        rows = self.rows[max(0, start_row - 1):] or self.rows[-3:]
        sm = [r[1] for r in rows]
        mask = 0
        for r in rows:
            mask |= r[2]
        reasons = [n for n, b in zip(self.NAMES, self.bits) if mask & b]
        return {"sm_mhz": float(np.median(sm)) if sm else None, "sm_max_mhz": self.max_sm,
                "reasons": reasons, "samples": len(rows)}


def peaks() -> tuple[float, str]:
    p = ROOT / "MEASURED_PEAKS.json"
    if p.exists():
        return float(json.loads(p.read_text())["hbm_gbs"]), "measured (MEASURED_PEAKS.json hbm_gbs)"
    return 6650.0, "fallback (B200_PROFILING.md 6.65 TB/s)"


def ncu_traffic(D: int, batch: int):
    """dram bytes per launch of the dominant kernel from the committed ncu capture, if any."""
    p = ROOT / "profiles" / "traffic.json"
    if p.exists():
        try:
            t = json.loads(p.read_text())
            key = f"bpr_phase_a:D{D}:B{batch}"
            return t.get(key)
        except (OSError, ValueError):
            return None
    return None


# ------------------------------------------------------------------------------------------------
# CPU baseline: the reference's own op sequence (oracle port), all host threads
# ------------------------------------------------------------------------------------------------
def cpu_reference_run(inter, D: int, batch: int, steps: int, warmup: int) -> dict:
    from oracle import ref_bpr
    torch.set_num_threads(os.cpu_count() or 1)
    ue, ie = init_tables(inter.num_users, inter.num_items, D)
    model = ref_bpr.RefModel(ue, ie, None, REG)
    opt = ref_bpr.make_optimizer(model, "sgd", lr=LR)
    gen = torch.Generator().manual_seed(SEED)
    weights = torch.ones(inter.num_items)
    coo = inter.coo_users()
    perm = torch.randperm(inter.nnz, generator=torch.Generator().manual_seed(SEED)).numpy()
    times = []
    for s in range(warmup + steps):
        t = perm[(s * batch) % (inter.nnz - batch):][:batch]
        users = torch.as_tensor(coo[t])
        items = torch.as_tensor(inter.indices[t], dtype=torch.long)
        t0 = time.perf_counter()
        seen = ref_bpr.padded_seen(inter.indptr, inter.indices, users)  # dataset.py:175-181 (fancy index)
        t_collate = time.perf_counter() - t0
        t0 = time.perf_counter()
        neg = ref_bpr.reference_style_negatives(weights, seen, gen)
        ref_bpr.train_step(model, opt, users, items, neg)
        dt = time.perf_counter() - t0
        if s >= warmup:
            times.append(dt)  # the padded-seen gather is a precomputed matrix in the reference: not timed
        del t_collate
    total = float(np.sum(times))
    return {"value": batch * len(times) / total, "ms_per_step": 1e3 * total / len(times),
            "cores": torch.get_num_threads(), "steps": len(times), "batch": batch}


def run_reference(args) -> None:
    rank = int(os.environ.get("RANK", "0"))
    if rank != 0:
        return
    inter = load_interactions(args.shape, args.scale)
    b = 256  # the reference's own train batch (README.md:305); a step is a bounded sample
    r = cpu_reference_run(inter, args.dim, b, args.steps, args.warmup)
    line = {
        "impl": "reference", "metric": METRIC, "value": r["value"], "unit": UNIT, "n_gpus": args.gpus,
        "steps": args.steps, "warmup": args.warmup, "ms_per_step": r["ms_per_step"],
        "higher_is_better": True, "scaling": "weak", "vs_baseline": None, "dtype": "f32",
        "data": "synthetic",
        "config": {"workload": f"{args.shape} shape {inter.num_users - 1}x{inter.num_items - 1}, "
                               f"{inter.nnz} interactions, dim={args.dim}, SGD, CPU oracle port of the "
                               f"reference op sequence, batch={b} triples per step",
                   "batch": b, "dim": args.dim},
        "cpu_baseline": {"value": r["value"], "unit": UNIT, "cores": r["cores"], "kind": "port",
                         "sample": f"{r['steps']} steps of {b} triples (reference default batch), "
                                   "multinomial sampler + autograd + dense SGD"},
        "e2e": {"value": r["value"], "unit": UNIT, "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
        "gpu_launches": 0,
    }
    print(json.dumps(line), flush=True)


# ------------------------------------------------------------------------------------------------
# our arm
# ------------------------------------------------------------------------------------------------
def run_ours(args) -> None:
    import torch.distributed as dist
    from rbpr import native
    from rbpr.engine import Engine

    world = int(os.environ.get("WORLD_SIZE", "1"))
    rank = int(os.environ.get("RANK", "0"))
    local = int(os.environ.get("LOCAL_RANK", "0"))
    if world != args.gpus:
        if world == 1 and args.gpus > 1:
            raise SystemExit("launch with torchrun --nproc-per-node N for --gpus N > 1")
    if not torch.cuda.is_available():
        raise SystemExit("bench.py (our arm) needs a B200: there is no CPU fallback")
    torch.cuda.set_device(local)
    dev = torch.device("cuda", local)
    if world > 1:
        dist.init_process_group("nccl", device_id=dev)

    inter = load_interactions(args.shape, args.scale)
    D, B, K, W = args.dim, args.batch, args.steps, args.warmup
    ue, ie = init_tables(inter.num_users, inter.num_items, D)
    eng = Engine(ue.to(dev), ie.to(dev))
    eng.bind_csr(torch.from_numpy(inter.indptr), torch.from_numpy(inter.indices))
    if args.opt == "adam":  # configs/RQ2/neg-sampling/adam-ada-sampling-msd.yaml.j2:152-160
        eng.set_reg({"all": 0.00043})
        eng.set_adam(1e-3, (0.9, 0.999), 1e-8)
    else:
        eng.set_reg(REG)
        eng.set_sgd(LR)
    if args.sampler == "adaptive":  # experiments/bpr/exp.py:194-207; config default prob 1/100
        import math
        every = max(1, int(inter.num_items * math.log(inter.num_items) / args.batch))
        eng.set_adaptive(0.01, every)
        eng.adaptive_update_stats()
    else:
        eng.set_sampler(native.SAMPLER_UNIFORM)

    from rbpr.parallel import DataParallelTrainer, owned_triples
    lo, hi = owned_triples(inter.indptr, world, rank)
    n_local = hi - lo
    g = torch.Generator(device=dev).manual_seed(SEED + rank)
    need = (W + K) * B
    perms = []
    while sum(p.numel() for p in perms) < need:  # epoch permutations of the owned triples
        perms.append(torch.randperm(n_local, generator=g, device=dev) + lo)
    perm = torch.cat(perms)[:need].contiguous()
    perm_host = perm.cpu().pin_memory()

    def barrier():
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize()

    if world > 1:  # the library joins its own NCCL communicator and runs the exchange itself
        eng.init_comm()
    del DataParallelTrainer

    def run_steps(t_dev: torch.Tensor, step0: int):
        """Device-resident steps in ONE library call; with N>1 every step all-reduces the dense
        item gradient once (NCCL, inside the library) before the replicated item update."""
        return eng.train_steps(t_dev, B, SEED, step0)[0]

    # ---- warm-up (the clock sampler starts first: nvidia-smi's start-up must not overlap the timed region) ----
    clocks = ClockSampler(local, args.clock_period_ms) if rank == 0 else None
    run_steps(perm[:W * B], 0)
    barrier()
    eng.sync_check()

    # ---- timed: device-resident ----
    time.sleep(0.3)
    eng.kernel_timing(True)
    eng.kernel_time_ms()
    l0 = eng.launch_count()
    c0 = eng.collective_count()
    mark = clocks.mark() if clocks else 0
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    barrier()
    e0.record()
    stats = run_steps(perm[W * B:], W)
    e1.record()
    barrier()
    ms = e0.elapsed_time(e1)
    launches = eng.launch_count() - l0
    collectives = eng.collective_count() - c0
    k_ms, k_n = eng.kernel_time_ms()
    eng.kernel_timing(False)
    eng.sync_check()
    if world > 1:
        tm = torch.tensor([ms], device=dev, dtype=torch.float64)
        dist.all_reduce(tm, op=dist.ReduceOp.MAX)
        ms = tm.item()
    loss_last = stats[-1, 0].item() / max(stats[-1, 3].item(), 1.0)

    # ---- timed: end to end with host buffers (H2D of the step's triple ids, D2H of its stats) ----
    spc = max(1, min(args.e2e_steps_per_call, K))
    # untimed: first use of the host-buffer entry point (staging + pinned result buffers)
    eng.train_steps_host(perm_host[:spc * B], B, SEED, W + K)
    barrier()
    t0 = time.perf_counter()
    e2e_loss = 0.0
    for s in range(0, K, spc):
        th = perm_host[(W + s) * B:(W + min(s + spc, K)) * B]
        st, _ = eng.train_steps_host(th, B, SEED, W + K + spc + s)
        e2e_loss = st[-1, 0].item()
    barrier()
    e2e_s = time.perf_counter() - t0
    if world > 1:
        tm = torch.tensor([e2e_s], device=dev, dtype=torch.float64)
        dist.all_reduce(tm, op=dist.ReduceOp.MAX)
        e2e_s = tm.item()
    clk = clocks.stop(mark) if clocks else None

    if rank == 0:
        peak, peak_src = peaks()
        total = K * B * world
        value = total / (ms * 1e-3)
        alg_bytes = 24 * D * B  # 3 rows read + 3 rows written, per triple, per launch
        achieved = alg_bytes / (k_ms / max(k_n, 1) * 1e-3) / 1e9 if k_n else None
        working_set_mb = ((inter.num_users + 2 * inter.num_items) * D * 4 + inter.nnz * 8 + need * 12) / 2**20
        line = {
            "metric": METRIC, "value": value, "unit": UNIT, "n_gpus": world, "steps": K, "warmup": W,
            "ms_per_step": ms / K, "higher_is_better": True, "scaling": "weak", "vs_baseline": None,
            "dtype": "f32", "data": "synthetic",
            "config": {"workload": f"BASELINE configs[1]: synthetic {args.shape} shape "
                                   f"{inter.num_users - 1}x{inter.num_items - 1}, {inter.nnz} interactions, "
                                   f"dim={D}, {'Adam lr=0.001 (dense-Adam semantics, lazy user rows)' if args.opt == 'adam' else f'SGD lr={LR}'}, "
                                   f"{args.sampler} on-device negatives, batch={B} triples/step"
                                   + (f" per GPU, users sharded by owner over {world} GPUs, one NCCL "
                                      "all-reduce of the dense item gradient per step" if world > 1 else ""),
                       "batch": B, "dim": D, "l2_policy": f"inputs larger than L2: working set "
                                                           f"{working_set_mb:.0f} MB vs 126 MB L2 (tables, CSR, "
                                                           "epoch permutation); consecutive steps are dependent "
                                                           "training steps, no flush",
                       "final_bpr_loss_per_triple": loss_last},
            "roofline": {"bound": "hbm", "achieved": achieved, "peak": peak, "unit": "GB/s",
                         "frac": (achieved / peak) if achieved else None, "traffic": ncu_traffic(D, B),
                         "kernel": "bpr_phase_a", "algorithmic_bytes_per_launch": alg_bytes,
                         "kernel_ms_avg": (k_ms / k_n) if k_n else None, "kernel_launches_timed": k_n,
                         "kernel_share_of_step": (k_ms / k_n * K / ms) if k_n else None, "peak_source": peak_src,
                         "frac_of_nominal_8TBs": (achieved / 8000.0) if achieved else None},
            "e2e": {"value": K * B * world / e2e_s, "unit": UNIT, "h2d_bytes_per_step": B * 8,
                    "d2h_bytes_per_step": 4 * 8, "ms_per_step": 1e3 * e2e_s / K,
                    "api": f"rbpr_train_steps_host (C ABI, pinned host buffers), {spc} steps per call"},
            "gpu_launches": launches, "nccl_allreduces": collectives, "clocks": clk,
        }
        if not args.no_cpu_baseline and world == 1:
            r = cpu_reference_run(inter, D, 256, args.cpu_steps, 2)
            line["cpu_baseline"] = {"value": r["value"], "unit": UNIT, "cores": r["cores"], "kind": "port",
                                    "sample": f"{r['steps']} steps of 256 triples (reference default batch, "
                                              f"{r['ms_per_step']:.1f} ms/step) of the same workload: oracle port "
                                              "of the reference op sequence (materialised (B,I) weights + "
                                              "multinomial, autograd, dense SGD)"}
        print(json.dumps(line), flush=True)
    if world > 1:
        dist.destroy_process_group()


def main():
    args = parse()
    if args.impl == "reference":
        run_reference(args)
    else:
        run_ours(args)


if __name__ == "__main__":
    main()
