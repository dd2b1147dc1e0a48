#!/usr/bin/env python
"""Benchmark of the B200-native BPR hot path (BASELINE.json metric: BPR triples/sec at dim=128 on a
synthetic ML-20M-shape matrix; achieved GB/s vs the HBM roofline).

  python bench.py [--gpus N] [--steps K] [--warmup W]          our arm (CUDA, sm_100a)
  python bench.py --impl reference [...]                       the reference's own CPU path

Headline line (`value`): BASELINE configs[1] — ML-20M shape, D=128, SGD, on-device uniform negatives,
--batch triples per step and per GPU (default 65 536, the large batch SURVEY.md §8(d) names for C2).
One "step" = one minibatch through the fused path: negative sampling, (u,i+,i-) gather, loss, exact
minibatch gradients, optimizer update.  The same JSON line carries, under `configs`, short runs of
the other BASELINE configurations (C2 at B=256 and 262 144, C3 MSD/D=256/Adam, C4 Yelp/D=64/adaptive,
C5 full-catalog scoring, and the experiment-level plugin surface), each with its own roofline.
Under torchrun (N>1) users are sharded by owner, the item table is replicated and the dense item
gradient is exchanged once per step; `value` is the whole-job aggregate, and a data-parallel parity
check against the oracle (SGD and Adam) runs before anything is timed (`parity_check`).
"""
from __future__ import annotations

import argparse
import json
import math
import os
import subprocess
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
REG_MSD = {"all": 0.00043}                              # configs/RQ2/neg-sampling/adam-ada-sampling-msd.yaml.j2:152-160
LR = 0.001
SEED = 13
L2_BYTES = 126e6


def parse():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=60)
    ap.add_argument("--warmup", type=int, default=5)
    ap.add_argument("--impl", default="ours", choices=["ours", "reference"])
    ap.add_argument("--shape", default="ml-20m")
    ap.add_argument("--dim", type=int, default=128)
    ap.add_argument("--sampler", default="uniform", choices=["uniform", "adaptive"])
    ap.add_argument("--opt", default="sgd", choices=["sgd", "adam"])
    ap.add_argument("--batch", type=int, default=65536,
                    help="triples per step and per GPU (train_batch_size is a free jinja variable of the reference configs)")
    ap.add_argument("--configs", default="auto",
                    help="comma list of extra configurations measured into `configs` (c2_b256,c2_b262144,c3,c4,c5,experiment), "
                         "'all', 'none', or 'auto' (all at N=1; c3,c5 at N>1)")
    ap.add_argument("--scale", type=float, default=1.0, help="shrink the synthetic matrices (tests)")
    ap.add_argument("--no-cpu-baseline", action="store_true")
    ap.add_argument("--no-parity-check", action="store_true")
    ap.add_argument("--clock-period-ms", type=float, default=5.0, help="NVML sampling period (0 = off)")
    ap.add_argument("--cpu-steps", type=int, default=12)
    ap.add_argument("--ref-sample", type=int, default=512,
                    help="reference arm: triples of each batch the CPU step actually processes (bounded sample)")
    ap.add_argument("--e2e-steps-per-call", type=int, default=60,
                    help="steps handed to one host-buffer API call in the e2e leg")
    ap.add_argument("--leg", default=None, help=argparse.SUPPRESS)  # internal: cpu legs run as a child process
    return ap.parse_args()


# ------------------------------------------------------------------------------------------------
# shared helpers
# ------------------------------------------------------------------------------------------------
def load_interactions(shape: str, scale: float, rank: int = 0, barrier=None):
    """Rank 0 generates (and caches under $RBPR_CACHE, default /tmp); the other ranks wait and load."""
    from rbpr import synth
    cache = Path(os.environ.get("RBPR_CACHE", "/tmp")) / f"rbpr_synth_{shape}_{scale}_{SEED}.npz"
    if rank != 0 and barrier is not None:
        barrier()
    if cache.exists():
        z = np.load(cache)
        inter = synth.Interactions(shape, int(z["U"]), int(z["I"]), z["indptr"], z["indices"])
    else:
        inter = synth.make(shape, seed=SEED, scale=scale)
        try:
            tmp = cache.with_suffix(f".{os.getpid()}.tmp.npz")
            np.savez(tmp, U=inter.num_users, I=inter.num_items, indptr=inter.indptr, indices=inter.indices)
            os.replace(tmp, cache)
        except OSError:
            pass
    if rank == 0 and barrier is not None:
        barrier()
    return inter


def init_tables(U: int, I: int, D: int, dev=None):
    """MF.reset_parameters (revisit_bpr/models/bpr/model.py:117-129) under torch.manual_seed(13)."""
    g = torch.Generator(device=dev or "cpu").manual_seed(SEED)
    ue = (torch.rand(U, D, generator=g, device=dev) - 0.5) / D
    ie = (torch.rand(I, D, generator=g, device=dev) - 0.5) / D
    ue[0] = 0
    ie[0] = 0
    return ue, ie


def workload(shape: str, inter_dims: tuple[int, int, int], D: int, opt: str, sampler: str, B: int, world: int) -> dict:
    """The `config` object — a function of the command line only, so both arms print the same one."""
    U, I, nnz = inter_dims
    optim = "Adam lr=0.001 betas=(0.9,0.999) (dense torch.optim.Adam semantics)" if opt == "adam" else f"SGD lr={LR}"
    label = {("ml-20m", 128, "sgd", "uniform"): "BASELINE configs[1]", ("msd", 256, "adam", "uniform"): "BASELINE configs[2]",
             ("yelp", 64, "sgd", "adaptive"): "BASELINE configs[3]"}.get((shape, D, opt, sampler), "custom")
    hot_mb = ((U + I) * D * 4 * (3 if opt == "adam" else 1) + I * D * 4) / 1e6
    return {
        "workload": f"{label}: synthetic {shape} shape {U - 1}x{I - 1}, {nnz} interactions, dim={D}, {optim}, "
                    f"{sampler} negatives, batch={B} triples/step" + (f" per GPU x {world} GPUs (users sharded by owner, "
                    "item table replicated, one exchange of the dense item gradient per step)" if world > 1 else ""),
        "batch": B, "dim": D, "n_gpus": world,
        "l2_policy": (f"no flush: consecutive steps are dependent training steps on the same tables; the tables the "
                      f"step kernels read ({hot_mb:.0f} MB incl. gradient accumulator" +
                      (", Adam moments" if opt == "adam" else "") + ") " +
                      ("fit the 126 MB L2, so this configuration is L2-resident, not HBM-bound (see roofline.bound / "
                       "roofline.traffic)" if hot_mb * 1e6 < L2_BYTES else "exceed the 126 MB L2: inputs larger than L2")),
    }


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

    def window(self, a: int, b: int | None = None) -> dict:
        if self.h is None:
            return {"sm_mhz": None, "sm_max_mhz": None, "reasons": [f"nvml unavailable: {getattr(self, 'err', '')}"]}
        rows = self.rows[max(0, a - 1):b] or self.rows[-3:]
        sm = [r[1] for r in rows]
        mask = 0
        for r in rows:
            mask |= r[2]
        reasons = [n for n, bit in zip(self.NAMES, self.bits) if mask & bit]
        return {"sm_mhz": float(np.median(sm)) if sm else None, "sm_max_mhz": self.max_sm,
                "reasons": reasons, "samples": len(rows)}

    def stop(self):
        self._stop = True
        if self.h is not None:
            self.th.join(timeout=1.0)


def peaks() -> tuple[float, float, str]:
    p = ROOT / "MEASURED_PEAKS.json"
    if p.exists():
        z = json.loads(p.read_text())
        return float(z["hbm_gbs"]), float(z.get("bf16_tflops", 1593.5)), "measured (MEASURED_PEAKS.json hbm_gbs)"
    return 6650.0, 1590.0, "fallback (B200_PROFILING.md 6.65 TB/s)"


def ncu_traffic(key: str):
    """DRAM bytes (read + write) per launch of a kernel from the committed ncu --set full captures
    (profiles/traffic.json; keys kernel:shape:D<dim>:B<batch>:<opt>:N<gpus>), or None."""
    p = ROOT / "profiles" / "traffic.json"
    try:
        return json.loads(p.read_text()).get(key)
    except (OSError, ValueError):
        return None


def roofline(kernel: str, key: str, alg_bytes: float, kernel_ms: float | None, hot_bytes: float, extra: dict | None = None) -> dict:
    """`achieved` = algorithmic bytes per launch / measured launch time (24*D B per
    triple).  `bound` says what that number is a fraction OF: when the measured DRAM traffic is less
    than half the algorithmic bytes (or, without a capture, the tables fit L2) the kernel is served
    by L2 and `frac_algorithmic` may exceed 1; `frac_hbm_dram` is the real HBM utilisation."""
    peak, _, src = peaks()
    traffic = ncu_traffic(key)
    achieved = alg_bytes / (kernel_ms * 1e-3) / 1e9 if kernel_ms else None
    if traffic is not None:
        bound = "l2" if traffic < 0.5 * alg_bytes else "hbm"
    else:
        bound = "l2" if hot_bytes < L2_BYTES else "hbm"
    out = {"bound": bound, "achieved": achieved, "peak": peak, "unit": "GB/s",
           "frac": (achieved / peak) if achieved else None, "traffic": traffic, "kernel": kernel,
           "frac_algorithmic": (achieved / peak) if achieved else None,
           "frac_hbm_dram": (traffic / (kernel_ms * 1e-3) / 1e9 / peak) if (traffic is not None and kernel_ms) else None,
           "algorithmic_bytes_per_launch": alg_bytes, "kernel_ms_avg": kernel_ms, "traffic_key": key, "peak_source": src}
    if extra:
        out.update(extra)
    return out


# ------------------------------------------------------------------------------------------------
# reference arm / CPU legs (always a process of their own: the reference package shares its name
# with this repo's drop-in)
# ------------------------------------------------------------------------------------------------
def run_reference(args) -> None:
    rank = int(os.environ.get("RANK", "0"))
    if rank != 0:
        return
    from oracle import reference_arm as ra
    world = int(os.environ.get("WORLD_SIZE", str(args.gpus)))
    inter = load_interactions(args.shape, args.scale)
    if args.leg == "eval":
        from rbpr import synth
        users, seen, held = synth.split_heldout(inter, 10_000)
        r = ra.eval_throughput(users, seen, held, inter.num_users, inter.num_items, args.dim, batch=128, max_users=384)
        print(json.dumps({"leg": "eval", **r}), flush=True)
        return
    sample = min(args.ref_sample, args.batch)
    reg = REG_MSD if args.opt == "adam" else REG
    r = ra.train_throughput(inter.indptr, inter.indices, inter.coo_users(), inter.num_users, inter.num_items, args.dim,
                            sample, args.steps, args.warmup, opt=args.opt, lr=LR, reg=reg, seed=SEED,
                            sampler=args.sampler)
    what = ("UniformSampler.sample -> BPR.forward -> loss.backward() -> torch.optim step -> zero_grad "
            "(reference example.py:172-180), materialised (B,I) weights + multinomial, dense gradients, dense update")
    cpu = {"value": r["value"], "unit": UNIT, "cores": r["cores"], "kind": r["kind"],
           "sample": f"{r['steps']} steps, each the reference's own train step on a {sample}-triple sample of the "
                     f"{args.batch}-triple batch ({r['ms_per_step']:.0f} ms per sample step): {what}"}
    if args.leg == "train":
        print(json.dumps({"leg": "train", **cpu}), flush=True)
        return
    line = {
        "impl": "reference", "metric": METRIC, "value": r["value"], "unit": UNIT, "n_gpus": args.gpus,
        "steps": args.steps, "warmup": args.warmup, "ms_per_step": r["ms_per_step"] * args.batch / sample,
        "higher_is_better": True, "scaling": "weak", "vs_baseline": None, "dtype": "f32", "data": "synthetic",
        "config": workload(args.shape, (inter.num_users, inter.num_items, inter.nnz), args.dim, args.opt, args.sampler,
                           args.batch, world),
        "cpu_baseline": cpu,
        "e2e": {"value": r["value"], "unit": UNIT, "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
        "gpu_launches": 0,
    }
    print(json.dumps(line), flush=True)


def cpu_leg(args, leg: str, **over) -> dict | None:
    """Run a CPU leg of the reference arm in a child process and return its JSON object."""
    a = {"shape": args.shape, "dim": args.dim, "opt": args.opt, "sampler": args.sampler, "batch": args.batch, **over}
    cmd = [sys.executable, str(ROOT / "bench.py"), "--impl", "reference", "--leg", leg, "--steps", str(args.cpu_steps),
           "--warmup", "2", "--scale", str(args.scale), "--ref-sample", str(args.ref_sample)]
    for k, v in a.items():
        cmd += [f"--{k}", str(v)]
    env = {**os.environ, "CUDA_VISIBLE_DEVICES": ""}
    for k in ("RANK", "WORLD_SIZE", "LOCAL_RANK"):
        env.pop(k, None)
    try:
        out = subprocess.run(cmd, capture_output=True, text=True, timeout=600, env=env)
        for ln in reversed(out.stdout.strip().splitlines()):
            if ln.startswith("{"):
                d = json.loads(ln)
                d.pop("leg", None)
                return d
        return {"error": (out.stderr or out.stdout)[-300:]}
    except Exception as e:  # noqa: BLE001
        return {"error": repr(e)}


# ------------------------------------------------------------------------------------------------
# our arm
# ------------------------------------------------------------------------------------------------
class Runner:
    def __init__(self, args):
        import torch.distributed as dist
        self.args, self.dist = args, dist
        self.world = int(os.environ.get("WORLD_SIZE", "1"))
        self.rank = int(os.environ.get("RANK", "0"))
        self.local = int(os.environ.get("LOCAL_RANK", "0"))
        if self.world == 1 and args.gpus > 1:
            raise SystemExit("launch with torchrun --nproc-per-node N for --gpus N > 1")
        if not torch.cuda.is_available():
            raise SystemExit("bench.py (our arm) needs a B200: there is no CPU fallback")
        torch.cuda.set_device(self.local)
        self.dev = torch.device("cuda", self.local)
        if self.world > 1:
            dist.init_process_group("nccl", device_id=self.dev)
        self.clocks = ClockSampler(self.local, args.clock_period_ms) if self.rank == 0 else None

    def barrier(self):
        if self.world > 1:
            self.dist.barrier()
        torch.cuda.synchronize()

    def align(self):
        """After barrier(): a device-side rendezvous on the current stream with NO host wait, so the
        timed region of every rank starts at the same point of the GPU timeline (the hosts leave
        dist.barrier() hundreds of microseconds apart, which a 60-step region of ~0.1 ms steps would
        otherwise book as step time: the first exchange waits for the last host)."""
        if self.world > 1:
            if not hasattr(self, "_align_buf"):
                self._align_buf = torch.zeros(1, device=self.dev)
            self.dist.all_reduce(self._align_buf)

    def max_over_ranks(self, x: float) -> float:
        if self.world > 1:
            t = torch.tensor([x], device=self.dev, dtype=torch.float64)
            self.dist.all_reduce(t, op=self.dist.ReduceOp.MAX)
            return t.item()
        return x

    def sum_over_ranks(self, x: float) -> float:
        if self.world > 1:
            t = torch.tensor([x], device=self.dev, dtype=torch.float64)
            self.dist.all_reduce(t, op=self.dist.ReduceOp.SUM)
            return t.item()
        return x

    # ---- one training configuration ---------------------------------------------------------------
    def make_engine(self, inter, D: int, opt: str, sampler: str, B: int):
        from rbpr import native
        from rbpr.engine import Engine
        ue, ie = init_tables(inter.num_users, inter.num_items, D, self.dev)
        eng = Engine(ue, ie)
        eng.bind_csr(torch.from_numpy(inter.indptr), torch.from_numpy(inter.indices))
        if opt == "adam":
            eng.set_reg(REG_MSD)
            eng.set_adam(1e-3, (0.9, 0.999), 1e-8)
        else:
            eng.set_reg(REG)
            eng.set_sgd(LR)
        if sampler == "adaptive":  # experiments/bpr/exp.py:194-207; config default prob 1/100
            every = max(1, int(inter.num_items * math.log(inter.num_items) / B))
            eng.set_adaptive(0.01, every)
            eng.adaptive_update_stats()
        else:
            eng.set_sampler(native.SAMPLER_UNIFORM)
        if self.world > 1:
            eng.init_comm()
            if os.environ.get("RBPR_FUSED_EXCHANGE", "1") != "0" and hasattr(eng, "init_fused_exchange"):
                eng.init_fused_exchange()
        return eng

    def permutation(self, inter, need: int):
        from rbpr.parallel import owned_triples
        lo, hi = owned_triples(inter.indptr, self.world, self.rank)
        g = torch.Generator(device=self.dev).manual_seed(SEED + self.rank)
        perms, have = [], 0
        while have < need:  # epoch permutations of the owned triples
            perms.append(torch.randperm(hi - lo, generator=g, device=self.dev) + lo)
            have += hi - lo
        return torch.cat(perms)[:need].contiguous()

    def train_config(self, shape: str, D: int, opt: str, sampler: str, B: int, K: int, W: int, e2e: bool = False) -> dict:
        """W warm-up + K timed steps, device-resident, one library call; CUDA events, max over ranks."""
        args = self.args
        inter = load_interactions(shape, args.scale, self.rank, self.barrier if self.world > 1 else None)
        eng = self.make_engine(inter, D, opt, sampler, B)
        perm = self.permutation(inter, (W + K) * B)
        eng.train_steps(perm[:W * B], B, SEED, 0)
        self.barrier()
        eng.sync_check()
        time.sleep(0.2)
        eng.kernel_timing(True)
        eng.kernel_time_ms()
        l0, c0 = eng.launch_count(), eng.collective_count()
        x0 = eng.fused_exchange_count() if hasattr(eng, "fused_exchange_count") else 0
        mark = self.clocks.mark() if self.clocks else 0
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        self.barrier()
        self.align()
        e0.record()
        stats = eng.train_steps(perm[W * B:], B, SEED, W)[0]
        e1.record()
        self.barrier()
        ms = self.max_over_ranks(e0.elapsed_time(e1))
        mark1 = self.clocks.mark() if self.clocks else 0
        launches, collectives = eng.launch_count() - l0, eng.collective_count() - c0
        exchanges = (eng.fused_exchange_count() - x0) if hasattr(eng, "fused_exchange_count") else 0
        k_ms, k_n = eng.kernel_time_ms()
        eng.kernel_timing(False)
        eng.sync_check()
        loss_last = stats[-1, 0].item() / max(stats[-1, 3].item(), 1.0)
        res = {"inter": inter, "ms": ms, "value": K * B * self.world / (ms * 1e-3), "ms_per_step": ms / K,
               "launches": launches, "collectives": collectives, "fused_exchanges": exchanges,
               "exchange": (("nvswitch multicast (multimem.ld_reduce / multimem.st)" if getattr(eng, "multicast", False)
                             else "peer loads / stores") if exchanges else ("nccl all-reduce" if collectives else None)),
               "kernel_ms": (k_ms / k_n) if k_n else None, "kernel_n": k_n, "loss": loss_last,
               "clocks": self.clocks.window(mark, mark1) if self.clocks else None}
        if e2e:  # end to end with host buffers: H2D of the step's triple ids, D2H of its stats, every call
            spc = max(1, min(args.e2e_steps_per_call, K))
            perm_host = perm.cpu().pin_memory()
            eng.train_steps_host(perm_host[:spc * B], B, SEED, W + K)  # untimed: first use (staging buffers)
            self.barrier()
            t0 = time.perf_counter()
            for s in range(0, K, spc):
                th = perm_host[(W + s) * B:(W + min(s + spc, K)) * B]
                eng.train_steps_host(th, B, SEED, W + K + spc + s)
            self.barrier()
            e2e_s = self.max_over_ranks(time.perf_counter() - t0)
            res["e2e"] = {"value": K * B * self.world / e2e_s, "unit": UNIT, "h2d_bytes_per_step": B * 8,
                          "d2h_bytes_per_step": 4 * 8, "ms_per_step": 1e3 * e2e_s / K,
                          "api": f"rbpr_train_steps_host (C ABI, pinned host buffers), {spc} steps per call"}
        del eng, perm
        torch.cuda.empty_cache()
        return res

    def train_entry(self, name: str, shape: str, D: int, opt: str, sampler: str, B: int, K: int, W: int) -> dict:
        r = self.train_config(shape, D, opt, sampler, B, K, W)
        inter = r["inter"]
        cfg = workload(shape, (inter.num_users, inter.num_items, inter.nnz), D, opt, sampler, B, self.world)
        hot = ((inter.num_users + inter.num_items) * D * 4 * (3 if opt == "adam" else 1) + inter.num_items * D * 4)
        alg = 24.0 * D * B
        small = r["kernel_n"] == 0
        kern = "whole step (persistent small-batch kernel)" if small else "bpr_phase_a"
        k_ms = r["ms_per_step"] if small else r["kernel_ms"]
        key = f"bpr_phase_a:{shape}:D{D}:B{B}:{opt}:N{self.world}"
        extra = {"kernel_launches_timed": r["kernel_n"],
                 "kernel_share_of_step": (k_ms / r["ms_per_step"]) if k_ms else None,
                 "step_frac_algorithmic": alg / (r["ms_per_step"] * 1e-3) / 1e9 / peaks()[0]}
        if opt == "adam":  # Adam's m,v traffic is outside the north-star formula; state it
            extra["bytes_per_triple_incl_adam_state"] = 72 * D
        return {"metric": "BPR triples/sec", "value": r["value"], "unit": UNIT, "ms_per_step": r["ms_per_step"],
                "steps": K, "warmup": W, "config": cfg, "roofline": roofline(kern, key, alg, k_ms, hot, extra),
                "gpu_launches": r["launches"], "nccl_allreduces": r["collectives"], "fused_exchanges": r["fused_exchanges"], "exchange": r["exchange"],
                "final_bpr_loss_per_triple": r["loss"], "clocks": r["clocks"]}

    # ---- scoring configuration (BASELINE configs[4]) -------------------------------------------------
    def scoring_entry(self, reps: int = 5) -> dict:
        from rbpr import synth
        from rbpr.engine import Engine
        args = self.args
        inter = load_interactions("ml-20m", args.scale, self.rank, self.barrier if self.world > 1 else None)
        D = 128
        g = torch.Generator(device=self.dev).manual_seed(SEED)
        ue = torch.randn(inter.num_users, D, generator=g, device=self.dev) * 0.1
        ie = torch.randn(inter.num_items, D, generator=g, device=self.dev) * 0.1
        ue[0] = 0
        ie[0] = 0
        eng = Engine(ue, ie)
        users, seen, held = synth.split_heldout(inter, 10_000)
        n_all = users.size
        # eval users sharded over ranks (item table replicated): contiguous blocks
        a, b = (n_all * self.rank) // self.world, (n_all * (self.rank + 1)) // self.world
        sp, hp = seen[0][a:b + 1] - seen[0][a], held[0][a:b + 1] - held[0][a]
        si, hi = seen[1][seen[0][a]:seen[0][b]], held[1][held[0][a]:held[0][b]]
        dv = lambda x, dt: torch.from_numpy(np.ascontiguousarray(x)).to(self.dev, dt)  # noqa: E731
        u_d, seen_d, held_d = dv(users[a:b], torch.int64), (dv(sp, torch.int64), dv(si, torch.int32)), \
            (dv(hp, torch.int64), dv(hi, torch.int32))
        ks = [20, 100]
        out = eng.score_metrics(u_d, seen_d, held_d, ks, want=("ndcg", "recall"))  # warm-up (allocates scratch)
        self.barrier()
        l0 = eng.launch_count()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        for _ in range(reps):
            out = eng.score_metrics(u_d, seen_d, held_d, ks, want=("ndcg", "recall"))
        e1.record()
        self.barrier()
        ms = self.max_over_ranks(e0.elapsed_time(e1) / reps)
        launches = (eng.launch_count() - l0) // reps
        # end to end: host user ids + CSRs in, per-user metrics summed on the device, two scalars out
        hu, hseen, hheld = torch.from_numpy(users[a:b]).pin_memory(), (torch.from_numpy(sp), torch.from_numpy(np.ascontiguousarray(si))), \
            (torch.from_numpy(hp), torch.from_numpy(np.ascontiguousarray(hi)))
        t0 = time.perf_counter()
        o = eng.score_metrics(hu, hseen, hheld, ks, want=("ndcg", "recall"))
        sums = torch.stack([o["ndcg"][:, 1].sum(), o["recall"][:, 0].sum()]).cpu()
        e2e_s = self.max_over_ranks(time.perf_counter() - t0)
        ndcg = self.sum_over_ranks(float(out["ndcg"][:, 1].double().sum())) / n_all
        recall = self.sum_over_ranks(float(out["recall"][:, 0].double().sum())) / n_all
        eng.sync_check()
        flops = 2.0 * D * inter.num_items * n_all
        min_bytes = (inter.num_items * D * 4) * self.world + n_all * D * 4 + n_all * len(ks) * 2 * 4
        traffic = ncu_traffic(f"score:ml-20m:D{D}:U{n_all}:N{self.world}")
        peak_gbs, peak_bf16, _ = peaks()
        # dominant kernel: score_tc, two tcgen05 kind::tf32 passes over the padded (users x items x K) problem
        # (K = D + bias column + "never" column, padded to 32-float swizzle atoms); tf32 dense peak = half
        # the measured bf16 rate.  `achieved` divides by the WHOLE call (pack, 2 passes, select, rescore):
        # the kernel alone is in profiles/round2/z_launches_c5_score.txt.
        tc_passes, tc_ovf = eng.score_path_counts() if hasattr(eng, "score_path_counts") else (0, 0)
        up, ip = -(-(b - a) // 128) * 128, -(-inter.num_items // 128) * 128
        kp = -(-(D + 1) // 32) * 32
        tc_flops = 2 * 2.0 * up * ip * kp
        tc_roof = {"bound": "tensor" if tc_passes else "fp32 FMA (dense fallback)",
                   "kernel": "score_tc (tcgen05 tf32, 2 passes) + select_threshold + rescore_rank",
                   "achieved": tc_flops / (ms * 1e-3) / 1e12, "peak": peak_bf16 / 2, "unit": "TFLOP/s",
                   "frac": tc_flops / (ms * 1e-3) / 1e12 / (peak_bf16 / 2),
                   "peak_source": "MEASURED_PEAKS.json bf16_tflops / 2 (tf32 runs at half the bf16 rate)",
                   "executed_tf32_flops_per_call": tc_flops, "useful_fp32_flops_per_call": 2.0 * D * inter.num_items * (b - a),
                   "algorithmic_min_bytes": min_bytes, "traffic": traffic,
                   "traffic_over_algorithmic": (traffic / min_bytes) if traffic else None,
                   "hbm_view_frac": min_bytes / (ms * 1e-3) / 1e9 / peak_gbs,
                   "tensor_passes": tc_passes, "overflow_users_on_dense_path": tc_ovf}
        entry = {"metric": "users/sec scored (full catalog, NDCG@100 + Recall@20)", "value": n_all / (ms * 1e-3),
                 "unit": "users/s", "ms": ms, "users": n_all, "ndcg@100": ndcg, "recall@20": recall,
                 "effective_tflops": flops / (ms * 1e-3) / 1e12, "gpu_launches_per_pass": launches,
                 "config": {"workload": f"BASELINE configs[4]: ML-20M shape, dim={D}, {n_all} eval users (20% of each user's "
                                        "items held out, the rest masked), whole set in one call" +
                                        (f", eval users sharded over {self.world} GPUs" if self.world > 1 else ""),
                            "n_gpus": self.world},
                 "roofline": tc_roof,
                 "e2e": {"value": n_all / e2e_s, "unit": "users/s", "h2d_bytes_per_step": int((b - a) * 8 + sp.nbytes + hp.nbytes + si.nbytes + hi.nbytes),
                         "d2h_bytes_per_step": int(sums.numel() * 4)}}
        del eng
        torch.cuda.empty_cache()
        return entry

    # ---- the plugin surface: experiments.bpr.Experiment from a reference-schema config -------------
    def experiment_entry(self) -> dict:
        from experiments_bench import run_experiment_bench  # noqa: PLC0415  (scripts/, our own code)
        return run_experiment_bench(self.dev)


def run_ours(args) -> None:
    R = Runner(args)
    world, rank = R.world, R.rank
    D, B, K, W = args.dim, args.batch, args.steps, args.warmup
    parity = None
    if world > 1 and not args.no_parity_check:
        sys.path.insert(0, str(ROOT / "tests" / "tools"))
        import dp_parity  # the oracle as CHECKER of the data-parallel step, before anything is timed
        msgs = []
        for o in ("sgd", "adam"):
            def setup(eng):
                if os.environ.get("RBPR_FUSED_EXCHANGE", "1") != "0" and hasattr(eng, "init_fused_exchange"):
                    eng.init_fused_exchange()
            ok, why = dp_parity.run(R.dev, rank, world, o, setup=setup)
            msgs.append(f"{o}: {'ok' if ok else 'FAILED ' + why}")
        parity = "ok" if all(m.endswith("ok") for m in msgs) else "; ".join(msgs)

    head = R.train_config(args.shape, D, args.opt, args.sampler, B, K, W, e2e=True)
    inter = head["inter"]
    cfg = workload(args.shape, (inter.num_users, inter.num_items, inter.nnz), D, args.opt, args.sampler, B, world)
    hot = (inter.num_users + inter.num_items) * D * 4 * (3 if args.opt == "adam" else 1) + inter.num_items * D * 4
    alg = 24.0 * D * B  # 3 rows read + 3 rows written, per triple, per launch
    small = head["kernel_n"] == 0
    k_ms = head["ms_per_step"] if small else head["kernel_ms"]

    want = os.environ.get("RBPR_BENCH_CONFIGS", args.configs)  # (scripts override the default set)
    if want == "auto":
        want = "all" if world == 1 else "c3,c5"
    names = ["c2_b256", "c2_b262144", "c3", "c4", "c5", "experiment"] if want == "all" else \
        ([] if want == "none" else [w.strip() for w in want.split(",") if w.strip()])
    configs: dict[str, dict] = {}
    for name in names:
        try:
            if name == "c2_b256":  # the reference configs' own batch size (README.md:305)
                configs[name] = R.train_entry(name, "ml-20m", 128, "sgd", "uniform", 256, 2000, 200)
            elif name == "c2_b262144":
                configs[name] = R.train_entry(name, "ml-20m", 128, "sgd", "uniform", 262144, 30, 4)
            elif name == "c3":
                configs["c3_msd_d256_adam"] = R.train_entry(name, "msd", 256, "adam", "uniform", 65536, 30, 4)
            elif name == "c4":
                configs["c4_yelp_d64_adaptive"] = R.train_entry(name, "yelp", 64, "sgd", "adaptive", 65536, 32, 4)
            elif name == "c5":
                configs["c5_scoring_ml20m"] = R.scoring_entry()
            elif name == "experiment" and world == 1:
                sys.path.insert(0, str(ROOT / "scripts"))
                configs["e2e_experiment"] = R.experiment_entry()
        except Exception as e:  # noqa: BLE001  a failing side configuration must not lose the headline
            configs[name] = {"error": repr(e)[:400]}
            torch.cuda.empty_cache()
    if R.clocks:
        R.clocks.stop()

    if rank == 0:
        key = f"bpr_phase_a:{args.shape}:D{D}:B{B}:{args.opt}:N{world}"
        line = {
            "metric": METRIC, "value": head["value"], "unit": UNIT, "n_gpus": world, "steps": K, "warmup": W,
            "ms_per_step": head["ms_per_step"], "higher_is_better": True, "scaling": "weak", "vs_baseline": None,
            "dtype": "f32", "data": "synthetic", "config": cfg,
            "roofline": roofline("whole step (persistent small-batch kernel)" if small else "bpr_phase_a", key, alg, k_ms, hot,
                                 {"kernel_launches_timed": head["kernel_n"],
                                  "kernel_share_of_step": (k_ms / head["ms_per_step"]) if k_ms else None,
                                  "step_frac_algorithmic": alg / (head["ms_per_step"] * 1e-3) / 1e9 / peaks()[0],
                                  "frac_of_nominal_8TBs": (alg / (k_ms * 1e-3) / 1e9 / 8000.0) if k_ms else None}),
            "e2e": head["e2e"], "gpu_launches": head["launches"], "nccl_allreduces": head["collectives"],
            "fused_exchanges": head["fused_exchanges"], "exchange": head["exchange"], "clocks": head["clocks"],
            "final_bpr_loss_per_triple": head["loss"], "configs": configs,
        }
        if parity is not None:
            line["parity_check"] = parity
        if not args.no_cpu_baseline and world == 1:
            line["cpu_baseline"] = cpu_leg(args, "train")
            if "c5_scoring_ml20m" in configs and "error" not in configs["c5_scoring_ml20m"]:
                ev = cpu_leg(args, "eval", shape="ml-20m", dim=128)
                if ev and "value" in ev:
                    configs["c5_scoring_ml20m"]["cpu_baseline"] = {
                        "value": ev["value"], "unit": "users/s", "cores": ev["cores"], "kind": ev["kind"],
                        "sample": f"{ev['users']} of the 10 000 eval users in batches of {ev['batch']}: MF eval forward over "
                                  "all items + scatter mask + NDCG(100) + Recall(20) (reference example.py:209-221)"}
        print(json.dumps(line), flush=True)
    if world > 1:
        R.dist.destroy_process_group()


def main():
    args = parse()
    if args.impl == "reference":
        run_reference(args)
    else:
        run_ours(args)


if __name__ == "__main__":
    main()
