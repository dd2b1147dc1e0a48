"""One native context bound to a (tables, interaction matrix) pair.

PyTorch owns every tensor (so `state_dict`, checkpointing and `MF.get_features()` keep
working); this class hands raw device pointers to librbpr.so and keeps the tensors alive.
All work is enqueued on torch's current CUDA stream.
"""
from __future__ import annotations

import ctypes as C
from typing import Sequence

import numpy as np
import torch

from rbpr import native


def _ptr(t: torch.Tensor | None) -> C.c_void_p:
    return C.c_void_p(0 if t is None else t.data_ptr())


def _stream() -> C.c_void_p:
    return C.c_void_p(torch.cuda.current_stream().cuda_stream)


def resolve_reg(reg_alphas: dict | None) -> tuple[float, float, float]:
    """(user, item, neg) coefficients with the reference's defaulting rules
    (revisit_bpr/models/bpr/model.py:74-86: `all` overrides; `neg` defaults to `item`)."""
    r = reg_alphas or {}
    all_reg, user, item, neg = r.get("all"), r.get("user"), r.get("item"), r.get("neg")
    if all(v is None for v in (all_reg, user, item, neg)):
        return 0.0, 0.0, 0.0
    if all_reg is not None:
        user = item = neg = all_reg
    user = user or 0
    item = item or 0
    neg = neg or item
    return float(user), float(item), float(neg)


class Context:
    """A bare native context on one device: what the table-less entry points need (negative
    sampling from padded seen matrices, ranking metrics from dense tensors)."""

    def __init__(self, device: torch.device) -> None:
        device = torch.device(device)
        if device.type != "cuda":
            raise native.NativeError("the BPR hot path runs on a B200 only: no CPU fallback exists")
        self.lib = native.load()
        self.device = device
        dev_index = device.index if device.index is not None else torch.cuda.current_device()
        self.ctx = C.c_void_p()
        rc = self.lib.rbpr_create(dev_index, C.byref(self.ctx))
        if rc != 0:
            raise native.NativeError(f"rbpr_create failed ({rc}): needs a compute-capability 10.x device")
        self._alias = None

    def close(self) -> None:
        """Destroy the native context now (scratch, communicator, peer mappings); idempotent."""
        if getattr(self, "ctx", None) is not None and self.ctx.value:
            self.lib.rbpr_destroy(self.ctx)
            self.ctx = C.c_void_p()

    def __del__(self) -> None:
        try:
            self.close()
        except Exception:
            pass

    def _check(self, rc: int) -> None:
        native.check(self.lib, self.ctx, rc)

    def sync_check(self) -> None:
        self._check(self.lib.rbpr_sync_check(self.ctx, _stream()))

    def launch_count(self) -> int:
        return int(self.lib.rbpr_launch_count(self.ctx))

    def score_path_counts(self) -> tuple[int, int]:
        """(blocks of users scored through the tensor-core filter, users it handed to the dense path)."""
        a, b = C.c_int64(), C.c_int64()
        self._check(self.lib.rbpr_score_path_counts(self.ctx, C.byref(a), C.byref(b)))
        return a.value, b.value

    def topk_launch_count(self) -> int:
        """Launches of the ranking kernel by this context (one per eval batch on the fused path)."""
        return int(self.lib.rbpr_topk_launch_count(self.ctx))

    def bind_item_weights(self, weights: torch.Tensor) -> None:
        """Popularity weights (I,) -> Walker alias table (weights[0] is forced to 0)."""
        w = weights.detach().double().cpu().numpy().copy()
        w[0] = 0.0
        prob, alias = build_alias(w)
        self._alias = (torch.from_numpy(prob).to(self.device), torch.from_numpy(alias).to(self.device))
        self._check(self.lib.rbpr_bind_item_alias(self.ctx, _ptr(self._alias[0]), _ptr(self._alias[1])))

    def sample_padded(self, seen: torch.Tensor, num_items: int, num: int, seed: int, step: int,
                      sampler: int = native.SAMPLER_UNIFORM) -> torch.Tensor:
        """Negatives (B,num) int64 for a 0-padded seen matrix (B,S) int64 on this device."""
        if seen.dim() != 2:
            raise ValueError("seen_items must be (batch, width)")
        seen = seen.to(self.device, torch.int64).contiguous()
        out = torch.empty((seen.size(0), num), dtype=torch.int64, device=self.device)
        self._check(self.lib.rbpr_sample_negatives_padded(
            self.ctx, _ptr(seen), seen.size(0), seen.size(1), num_items, num, seed & (2**64 - 1), step,
            sampler, _ptr(out), _stream()))
        return out

    def mask_seen_padded(self, logits: torch.Tensor, seen: torch.Tensor) -> torch.Tensor:
        """In place: logits[b, seen[b,:]] = -1e13, logits[b, 0] = -1e13 (exp.py:369-374)."""
        if not (logits.is_cuda and logits.dtype == torch.float32 and logits.is_contiguous() and logits.dim() == 2):
            raise ValueError("logits must be a contiguous float32 CUDA matrix")
        seen = seen.to(self.device, torch.int64).contiguous()
        self._check(self.lib.rbpr_mask_seen_padded(self.ctx, _ptr(logits), _ptr(seen), logits.size(0),
                                                   seen.size(1), logits.size(1), _stream()))
        return logits

    def auc_dense(self, output: torch.Tensor, target: torch.Tensor, mask: torch.Tensor | None = None) -> torch.Tensor:
        output = output.to(self.device, torch.float32).contiguous()
        target = target.to(self.device, torch.float32).contiguous()
        if mask is not None:
            mask = mask.to(self.device, torch.float32).contiguous()
        out = torch.empty(output.size(0), dtype=torch.float32, device=self.device)
        self._check(self.lib.rbpr_auc_dense(self.ctx, _ptr(output), _ptr(target), _ptr(mask), output.size(0),
                                            output.size(1), _ptr(out), _stream()))
        return out

    def topk_metrics_dense(self, output: torch.Tensor, target: torch.Tensor, ks: Sequence[int],
                           linear_gain: bool = False, want_items: bool = False,
                           map_normalized: bool = True) -> dict[str, torch.Tensor]:
        """NDCG / Recall / Precision at every cut-off of `ks` from dense (B,I) scores and targets."""
        output = output.to(self.device, torch.float32).contiguous()
        target = target.to(self.device, torch.float32).contiguous()
        n, cols = output.shape
        ks = [min(int(k), cols) for k in ks]
        k_max = max(ks)
        if k_max > native.MAX_TOPK:
            raise ValueError(f"topk={k_max} exceeds the kernel limit {native.MAX_TOPK}")
        out = {name: torch.empty((n, len(ks)), dtype=torch.float32, device=self.device)
               for name in ("ndcg", "recall", "precision", "map")}
        items = torch.empty((n, k_max), dtype=torch.int32, device=self.device) if want_items else None
        ks_arr = (C.c_int32 * len(ks))(*ks)
        self._check(self.lib.rbpr_topk_metrics_dense(
            self.ctx, _ptr(output), _ptr(target), n, cols, k_max, ks_arr, len(ks), int(linear_gain),
            _ptr(out["ndcg"]), _ptr(out["recall"]), _ptr(out["precision"]), _ptr(out["map"]),
            int(map_normalized), _ptr(items), _stream()))
        if want_items:
            out["items"] = items
        return out

    # ---- item-neighbourhood logits models (reference model.py:156-251) ---------------------------
    def _knn_ids(self, item: torch.Tensor, seen: torch.Tensor) -> tuple[torch.Tensor, torch.Tensor]:
        if item.dim() != 2 or seen.dim() != 2 or item.size(0) != seen.size(0):
            raise IndexError("item must be (batch, num items) and seen_items (batch, seen items)")
        return (item.to(self.device, torch.int64).contiguous(), seen.to(self.device, torch.int64).contiguous())

    def _knn_table(self, weights: torch.Tensor, bias: torch.Tensor | None) -> None:
        for t in (weights,) + (() if bias is None else (bias,)):
            if not (t.is_cuda and t.dtype == torch.float32 and t.is_contiguous()):
                raise ValueError("weights / bias must be contiguous float32 CUDA tensors")
        if weights.dim() != 2 or (bias is not None and bias.shape != weights.shape[:1]):
            raise ValueError("weights must be (num items, width) and bias (num items,)")

    def knn_forward(self, weights: torch.Tensor, bias: torch.Tensor | None, item: torch.Tensor,
                    seen: torch.Tensor) -> tuple[torch.Tensor, torch.Tensor, torch.Tensor]:
        """ItemKNN logits (B,n) plus what the backward needs: kept-row sums (B,H), keep flags (B,S)."""
        self._knn_table(weights, bias)
        item, seen = self._knn_ids(item, seen)
        B, n = item.shape
        I, H = weights.shape
        keep = torch.empty(seen.shape, dtype=torch.uint8, device=self.device)
        profile = torch.empty((B, H), dtype=torch.float32, device=self.device)
        logits = torch.empty((B, n), dtype=torch.float32, device=self.device)
        self._check(self.lib.rbpr_knn_forward(self.ctx, _ptr(weights), I, H, _ptr(bias), _ptr(item), B, n,
                                              _ptr(seen), seen.size(1), _ptr(keep), _ptr(profile), _ptr(logits),
                                              _stream()))
        return logits, profile, keep

    def knn_backward(self, weights: torch.Tensor, item: torch.Tensor, seen: torch.Tensor, keep: torch.Tensor,
                     profile: torch.Tensor, grad_logits: torch.Tensor, grad_weights: torch.Tensor,
                     grad_bias: torch.Tensor | None) -> None:
        item, seen = self._knn_ids(item, seen)
        I, H = weights.shape
        self._check(self.lib.rbpr_knn_backward(self.ctx, _ptr(weights), I, H, _ptr(item), item.size(0), item.size(1),
                                               _ptr(seen), seen.size(1), _ptr(keep), _ptr(profile),
                                               _ptr(grad_logits), _ptr(grad_weights), _ptr(grad_bias), _stream()))

    def freeknn_forward(self, weights: torch.Tensor, bias: torch.Tensor | None, item: torch.Tensor,
                        seen: torch.Tensor) -> tuple[torch.Tensor, torch.Tensor]:
        self._knn_table(weights, bias)
        if weights.size(0) != weights.size(1):
            raise ValueError("FreeItemKNN weights must be (num items, num items)")
        item, seen = self._knn_ids(item, seen)
        B, n = item.shape
        keep = torch.empty(seen.shape, dtype=torch.uint8, device=self.device)
        logits = torch.empty((B, n), dtype=torch.float32, device=self.device)
        self._check(self.lib.rbpr_freeknn_forward(self.ctx, _ptr(weights), weights.size(0), _ptr(bias), _ptr(item), B,
                                                  n, _ptr(seen), seen.size(1), _ptr(keep), _ptr(logits), _stream()))
        return logits, keep

    def freeknn_backward(self, num_items: int, item: torch.Tensor, seen: torch.Tensor, keep: torch.Tensor,
                         grad_logits: torch.Tensor, grad_weights: torch.Tensor,
                         grad_bias: torch.Tensor | None) -> None:
        item, seen = self._knn_ids(item, seen)
        self._check(self.lib.rbpr_freeknn_backward(self.ctx, num_items, _ptr(item), item.size(0), item.size(1),
                                                   _ptr(seen), seen.size(1), _ptr(keep), _ptr(grad_logits),
                                                   _ptr(grad_weights), _ptr(grad_bias), _stream()))


class Engine(Context):
    def __init__(self, user_emb: torch.Tensor, item_emb: torch.Tensor,
                 item_bias: torch.Tensor | None = None) -> None:
        if not user_emb.is_cuda:
            raise native.NativeError("the BPR hot path runs on a B200 only: tables must be CUDA tensors")
        for t in (user_emb, item_emb) + ((item_bias,) if item_bias is not None else ()):
            if t.dtype != torch.float32 or not t.is_contiguous():
                raise ValueError("tables must be contiguous float32")
        if user_emb.size(1) != item_emb.size(1):
            raise ValueError("user and item tables must share the embedding dim")
        super().__init__(user_emb.device)
        self.user_emb, self.item_emb, self.item_bias = user_emb, item_emb, item_bias
        self.U, self.D = user_emb.shape
        self.I = item_emb.size(0)
        self._check(self.lib.rbpr_bind_tables(self.ctx, _ptr(user_emb), self.U, _ptr(item_emb),
                                              self.I, self.D, _ptr(item_bias)))
        self.hp = native.HParams()
        self.hp.optimizer = native.OPT_SGD
        self.hp.sampler = native.SAMPLER_UNIFORM
        self.hp.lr = 0.0
        self.hp.beta1, self.hp.beta2, self.hp.eps = 0.9, 0.999, 1e-8
        self.adam_state: dict[str, torch.Tensor] | None = None
        self.opt_state: dict[str, torch.Tensor] | None = None
        self.opt_state_kind: str | None = None
        self.indptr = self.indices = None
        self.nnz = 0
        self._pin_stats = self._pin_neg = None

    # ---- configuration -------------------------------------------------------------------
    def set_reg(self, reg_alphas: dict | None) -> None:
        self.hp.reg_user, self.hp.reg_item, self.hp.reg_neg = resolve_reg(reg_alphas)

    def set_sgd(self, lr: float) -> None:
        self.hp.optimizer = native.OPT_SGD
        self.hp.lr = lr

    def set_adam(self, lr: float, betas: Sequence[float] = (0.9, 0.999), eps: float = 1e-8,
                 state: dict[str, torch.Tensor] | None = None) -> dict[str, torch.Tensor]:
        """Bind (or create zeroed) Adam state: user_m/v, user_last (int32), item_m/v, bias_m/v."""
        self.hp.optimizer = native.OPT_ADAM
        self.hp.lr, self.hp.beta1, self.hp.beta2, self.hp.eps = lr, betas[0], betas[1], eps
        if state is None:
            state = self.adam_state
        if state is None:
            z = torch.zeros_like
            state = {"user_m": z(self.user_emb), "user_v": z(self.user_emb),
                     "user_last": torch.zeros(self.U, dtype=torch.int32, device=self.device),
                     "item_m": z(self.item_emb), "item_v": z(self.item_emb)}
            if self.item_bias is not None:
                state["bias_m"], state["bias_v"] = z(self.item_bias), z(self.item_bias)
        self.adam_state = state
        self._check(self.lib.rbpr_bind_adam_state(
            self.ctx, _ptr(state["user_m"]), _ptr(state["user_v"]), _ptr(state["user_last"]),
            _ptr(state["item_m"]), _ptr(state["item_v"]), _ptr(state.get("bias_m")),
            _ptr(state.get("bias_v"))))
        return state

    def _set_state1(self, kind: int, key: str, state: dict[str, torch.Tensor] | None) -> dict[str, torch.Tensor]:
        """Bind (or create zeroed) single-state optimizer state: user_s, user_last (int32), item_s, bias_s."""
        self.hp.optimizer = kind
        if state is None:
            state = self.opt_state if getattr(self, "opt_state_kind", None) == key else None
        if state is None:
            z = torch.zeros_like
            state = {"user_s": z(self.user_emb), "item_s": z(self.item_emb),
                     "user_last": torch.zeros(self.U, dtype=torch.int32, device=self.device)}
            if self.item_bias is not None:
                state["bias_s"] = z(self.item_bias)
        self.opt_state, self.opt_state_kind = state, key
        self._check(self.lib.rbpr_bind_state1(self.ctx, _ptr(state["user_s"]), _ptr(state["user_last"]),
                                              _ptr(state["item_s"]), _ptr(state.get("bias_s"))))
        return state

    def set_sgd_momentum(self, lr: float, momentum: float, nesterov: bool = False,
                         state: dict[str, torch.Tensor] | None = None) -> dict[str, torch.Tensor]:
        """torch.optim.SGD(lr, momentum, nesterov) with dampening 0 (dense semantics, lazy user rows)."""
        self.hp.lr, self.hp.beta1, self.hp.beta2 = lr, momentum, 1.0 if nesterov else 0.0
        return self._set_state1(native.OPT_SGDM, "sgdm", state)

    def set_rmsprop(self, lr: float, alpha: float = 0.99, eps: float = 1e-8,
                    state: dict[str, torch.Tensor] | None = None) -> dict[str, torch.Tensor]:
        """torch.optim.RMSprop(lr, alpha, eps) with momentum 0, not centered."""
        self.hp.lr, self.hp.beta2, self.hp.eps = lr, alpha, eps
        return self._set_state1(native.OPT_RMSPROP, "rmsprop", state)

    def set_sampler(self, kind: int) -> None:
        self.hp.sampler = kind

    def bind_csr(self, indptr: torch.Tensor, indices: torch.Tensor) -> None:
        """indptr (U+1,) int64, indices (nnz,) int32 ascending per row; moved to the device."""
        indptr = indptr.to(self.device, torch.int64).contiguous()
        indices = indices.to(self.device, torch.int32).contiguous()
        if indptr.numel() != self.U + 1:
            raise ValueError(f"indptr has {indptr.numel()} entries, expected num_users+1={self.U + 1}")
        self.indptr, self.indices, self.nnz = indptr, indices, indices.numel()
        self._check(self.lib.rbpr_bind_csr(self.ctx, _ptr(indptr), _ptr(indices), self.U, self.nnz,
                                           _stream()))

    # ---- hot path --------------------------------------------------------------------------
    def sample(self, triple_idx: torch.Tensor, seed: int, step: int, sampler: int | None = None) -> torch.Tensor:
        triple_idx = triple_idx.to(self.device, torch.int64).contiguous()
        out = torch.empty_like(triple_idx)
        kind = self.hp.sampler if sampler is None else sampler
        self._check(self.lib.rbpr_sample_negatives(self.ctx, _ptr(triple_idx), triple_idx.numel(),
                                                   seed, step, kind, _ptr(out), _stream()))
        return out

    def train_steps(self, triple_idx: torch.Tensor, batch: int, seed: int, step0: int,
                    neg_in: torch.Tensor | None = None, want_neg: bool = False,
                    want_stats: bool = True):
        """Device-resident call. Returns (stats (steps,4) float64 device tensor | None, negs | None)."""
        n = triple_idx.numel()
        steps = (n + batch - 1) // batch
        stats = torch.empty((steps, native.STATS_PER_STEP), dtype=torch.float64, device=self.device) \
            if want_stats else None
        neg_out = torch.empty(n, dtype=torch.int64, device=self.device) if want_neg else None
        self._check(self.lib.rbpr_train_steps(self.ctx, _ptr(triple_idx), n, batch, seed, step0,
                                              C.byref(self.hp), _ptr(neg_in), _ptr(neg_out),
                                              _ptr(stats), _stream()))
        return stats, neg_out

    def train_steps_host(self, triple_idx: torch.Tensor, batch: int, seed: int, step0: int,
                         neg_in: torch.Tensor | None = None, want_neg: bool = False):
        """Host-buffer call (pinned CPU int64 in, CPU float64 stats out); synchronises."""
        assert not triple_idx.is_cuda and triple_idx.dtype == torch.int64
        n = triple_idx.numel()
        steps = (n + batch - 1) // batch
        # pinned result buffers are kept across calls (page-locking per call costs milliseconds)
        if self._pin_stats is None or self._pin_stats.size(0) < steps:
            self._pin_stats = torch.empty((max(steps, 256), native.STATS_PER_STEP),
                                          dtype=torch.float64).pin_memory()
        stats = self._pin_stats[:steps]
        neg_out = None
        if want_neg:
            if self._pin_neg is None or self._pin_neg.numel() < n:
                self._pin_neg = torch.empty(n, dtype=torch.int64).pin_memory()
            neg_out = self._pin_neg[:n]
        self._check(self.lib.rbpr_train_steps_host(self.ctx, _ptr(triple_idx), n, batch, seed, step0,
                                                   C.byref(self.hp), _ptr(neg_in), _ptr(neg_out),
                                                   _ptr(stats), _stream()))
        return stats, neg_out

    # ---- reference-shaped calls (explicit id batches) ---------------------------------------
    def train_step_triples(self, users: torch.Tensor, items: torch.Tensor, negs: torch.Tensor,
                           step: int, want_logits: bool = True):
        """One fused step on explicit (user, positive, negative) ids. Returns (logits (n,2) | None,
        stats (4,) float64 device tensor)."""
        dev = self.device
        users = users.to(dev, torch.int64).contiguous()
        items = items.to(dev, torch.int64).contiguous()
        negs = negs.to(dev, torch.int64).contiguous()
        n = users.numel()
        if items.numel() != n or negs.numel() != n:
            raise ValueError("user, item and neg must hold the same number of ids")
        logits = torch.empty((n, 2), dtype=torch.float32, device=dev) if want_logits else None
        stats = torch.empty(native.STATS_PER_STEP, dtype=torch.float64, device=dev)
        self._check(self.lib.rbpr_train_step_triples(self.ctx, _ptr(users), _ptr(items), _ptr(negs), n,
                                                     step, C.byref(self.hp), _ptr(logits), _ptr(stats),
                                                     _stream()))
        return logits, stats

    # ---- adaptive sampler ---------------------------------------------------------------------
    def set_adaptive(self, sampling_prob: float, every: int = 0) -> None:
        self.hp.sampler = native.SAMPLER_ADAPTIVE
        self.hp.adaptive_prob = float(sampling_prob)
        self.hp.adaptive_every = int(every)

    def adaptive_update_stats(self) -> bool:
        """Snapshot the item table, per-factor std, sorted factor orders (context-owned)."""
        self._check(self.lib.rbpr_adaptive_update_stats(self.ctx, _stream()))
        return True

    def adaptive_stats(self) -> dict[str, torch.Tensor]:
        """Views of the context-owned adaptive state: std (D,), order (D,I), pos (D,I)."""
        std, order, pos = C.c_void_p(), C.c_void_p(), C.c_void_p()
        self._check(self.lib.rbpr_adaptive_stats(self.ctx, C.byref(std), C.byref(order), C.byref(pos)))
        return {"std": _wrap_device(std.value, (self.D,), "<f4", self.device),
                "order": _wrap_device(order.value, (self.D, self.I), "<i4", self.device),
                "pos": _wrap_device(pos.value, (self.D, self.I), "<i4", self.device)}

    def sample_adaptive_padded(self, users: torch.Tensor, seen: torch.Tensor, num: int,
                               sampling_prob: float, seed: int, step: int, _stats: object = None,
                               opt_step: int | None = None) -> torch.Tensor:
        """`opt_step`: optimizer steps applied so far — with a stateful optimizer the lazily updated user
        rows are caught up in registers before the draw (None: rows are read as stored)."""
        users = users.to(self.device, torch.int64).contiguous()
        seen = seen.to(self.device, torch.int64).contiguous()
        if seen.dim() != 2 or seen.size(0) != users.numel():
            raise ValueError("seen_items must be (batch, width) with one row per user")
        out = torch.empty((users.numel(), num), dtype=torch.int64, device=self.device)
        self._check(self.lib.rbpr_sample_adaptive_padded(
            self.ctx, _ptr(users), _ptr(seen), users.numel(), seen.size(1), num, float(sampling_prob),
            seed & (2**64 - 1), step, _ptr(out), int(opt_step or 0),
            C.byref(self.hp) if opt_step is not None else None, _stream()))
        return out

    def pair_logits(self, users: torch.Tensor, items: torch.Tensor, mask: torch.Tensor | None = None,
                    user_bias: torch.Tensor | None = None) -> torch.Tensor:
        """logits (B, K) for users (B,) and items (B, K) [+ biases]; mask==0 entries -> -1e13."""
        dev = self.device
        users = users.to(dev, torch.int64).contiguous()
        items = items.to(dev, torch.int64).contiguous()
        per_user = items.numel() // max(users.numel(), 1)
        out = torch.empty(items.shape, dtype=torch.float32, device=dev)
        if mask is not None:
            mask = mask.to(dev, torch.float32).contiguous()
        self._check(self.lib.rbpr_pair_logits(self.ctx, _ptr(users), _ptr(items), _ptr(mask),
                                              users.numel(), per_user, _ptr(user_bias), _ptr(out),
                                              _stream()))
        return out

    # ---- data-parallel communicator -------------------------------------------------------------
    def init_comm(self, group=None) -> None:
        """Join the library's own NCCL communicator (one rank per process of the default
        torch.distributed group): rank 0's unique id travels through torch.distributed.  After
        this, train_steps / train_steps_host run the data-parallel step inside the library."""
        import importlib.util
        import os
        import torch.distributed as dist
        world, rank = dist.get_world_size(group), dist.get_rank(group)
        if "RBPR_NCCL_LIB" not in os.environ:  # the libnccl PyTorch itself ships
            spec = importlib.util.find_spec("nvidia")
            for root in (spec.submodule_search_locations if spec else []):
                cand = os.path.join(root, "nccl", "lib", "libnccl.so.2")
                if os.path.exists(cand):
                    os.environ["RBPR_NCCL_LIB"] = cand
                    break
        uid = torch.zeros(128, dtype=torch.uint8)
        if rank == 0:
            self._check(self.lib.rbpr_comm_unique_id(self.ctx, C.c_void_p(uid.data_ptr())))
        dev_uid = uid.to(self.device)
        dist.broadcast(dev_uid, src=dist.get_global_rank(group, 0) if group is not None else 0, group=group)
        uid = dev_uid.cpu()
        self._check(self.lib.rbpr_comm_init(self.ctx, world, rank, C.c_void_p(uid.data_ptr())))
        self.world, self.rank = world, rank

    def init_fused_exchange(self, group=None, strict: bool = False) -> bool:
        """Replace the per-step all-reduce + dense apply by the fused reduce + update + broadcast kernel
        over NVLink peer memory (csrc/exchange.cu).  Two bindings, tried in this order:
          * symmetric memory (torch.distributed._symmetric_memory; RBPR_FX_SYMM=0 skips it): one buffer
            per rank holding the gradient accumulators AND the item table / bias, mapped into every
            peer and aliased by an NVSwitch multicast address, so the reduction runs inside the switch
            (multimem.ld_reduce) and updated rows leave their owner once (multimem.st).  The item
            table and bias MOVE into that buffer: `self.item_emb` / `self.item_bias` are re-pointed in
            place (`Tensor.set_`), so the tensor objects handed to the constructor follow; callers
            holding other aliases (an nn.Parameter's `.data`) re-point them at `self.item_emb`;
          * cudaIpc handles of the library's accumulators and of the item table / bias storages,
            exchanged as blobs through torch.distributed (unicast peer loads / stores).
        Returns False (and keeps the NCCL path) when peer memory cannot be shared, unless `strict`.
        With a stateful optimizer the item table's optimizer state becomes SHARDED by row block
        (`item_cuts`); call `gather_item_state` before reading it (checkpoints)."""
        import os
        import torch.distributed as dist
        world, rank = dist.get_world_size(group), dist.get_rank(group)
        if world < 2 or world > native.MAX_PEERS:
            if strict:
                raise native.NativeError(f"the fused exchange needs 2..{native.MAX_PEERS} ranks on one node")
            return False
        if os.environ.get("RBPR_FX_SYMM", "1") != "0" and self._bind_symmetric(group, world, rank):
            self.world, self.rank, self.fused_exchange = world, rank, True
            return True
        blob = torch.zeros(native.IPC_BLOB_BYTES, dtype=torch.uint8)
        ok = torch.ones(1, dtype=torch.int32, device=self.device)
        err = ""
        try:
            self._check(self.lib.rbpr_comm_ipc_export(self.ctx, C.c_void_p(blob.data_ptr())))
        except native.NativeError as e:
            ok.zero_()
            err = str(e)
        dist.all_reduce(ok, op=dist.ReduceOp.MIN, group=group)
        if ok.item() == 0:
            if strict:
                raise native.NativeError(f"fused exchange unavailable: {err or 'a peer could not export its memory'}")
            return False
        gathered = [torch.zeros(native.IPC_BLOB_BYTES, dtype=torch.uint8, device=self.device) for _ in range(world)]
        dist.all_gather(gathered, blob.to(self.device), group=group)
        blobs = torch.cat(gathered).cpu().contiguous()
        try:
            self._check(self.lib.rbpr_comm_ipc_bind(self.ctx, C.c_void_p(blobs.data_ptr()), world, rank, _stream()))
        except native.NativeError as e:
            ok.zero_()
            err = str(e)
        dist.all_reduce(ok, op=dist.ReduceOp.MIN, group=group)
        if ok.item() == 0:
            raise native.NativeError(f"fused exchange: binding peer memory failed on some rank ({err}); the ranks are "
                                     "now inconsistent, restart without it (RBPR_FUSED_EXCHANGE=0)")
        self.world, self.rank, self.fused_exchange = world, rank, True
        return True

    def _bind_symmetric(self, group, world: int, rank: int) -> bool:
        """Allocate the symmetric buffer, rendezvous, bind (rbpr_comm_symm_bind) and re-point the item
        table / bias into it.  False (nothing changed) when any rank cannot allocate or map it."""
        import os
        import torch.distributed as dist
        ok = torch.ones(1, dtype=torch.int32, device=self.device)
        buf = hdl = None
        nbytes = int(self.lib.rbpr_comm_symm_bytes(self.ctx))
        try:
            import torch.distributed._symmetric_memory as symm
            pg = group if group is not None else dist.group.WORLD
            buf = symm.empty(nbytes // 4, dtype=torch.float32, device=self.device)
            buf.zero_()
            torch.cuda.synchronize(self.device)
            hdl = symm.rendezvous(buf, group=pg)
            bases = [int(p) for p in hdl.buffer_ptrs]
            if len(bases) != world or any(b == 0 or b % 256 for b in bases) or bases[rank] != buf.data_ptr():
                raise RuntimeError(f"unexpected symmetric mapping {bases}")
        except Exception as e:  # noqa: BLE001  (any failure means "not available here": fall back, all ranks together)
            ok.zero_()
            self._symm_error = repr(e)[:300]
        dist.all_reduce(ok, op=dist.ReduceOp.MIN, group=group)
        if ok.item() == 0:
            return False
        mc = int(getattr(hdl, "multicast_ptr", 0) or 0)
        # In-switch reduction pays from 4 ranks on (measured: N=2 unicast 68 us/step vs multicast 78;
        # N=4 equal; N=8 multicast 95 vs unicast 101): with 2 ranks the switch reads BOTH copies over
        # NVLink, the own one included.  RBPR_FX_MULTICAST=1 / 0 forces it on / off; off also gives
        # the rank-order (bit-reproducible) sum of the unicast path.
        want = os.environ.get("RBPR_FX_MULTICAST", "auto")
        if want == "0" or (want == "auto" and world < 4):
            mc = 0
        mcs = torch.tensor([mc != 0], dtype=torch.int32, device=self.device)
        dist.all_reduce(mcs, op=dist.ReduceOp.MIN, group=group)
        if mcs.item() == 0:
            mc = 0
        arr = (C.c_uint64 * world)(*bases)
        new_item, new_bias = C.c_uint64(0), C.c_uint64(0)
        err = ""
        try:
            self._check(self.lib.rbpr_comm_symm_bind(self.ctx, arr, C.c_uint64(mc), world, rank,
                                                      C.byref(new_item), C.byref(new_bias), _stream()))
        except native.NativeError as e:
            ok.zero_()
            err = str(e)
        dist.all_reduce(ok, op=dist.ReduceOp.MIN, group=group)
        if ok.item() == 0:
            raise native.NativeError(f"fused exchange: binding the symmetric buffer failed on some rank ({err}); the "
                                     "ranks are now inconsistent, restart with RBPR_FX_SYMM=0")
        torch.cuda.synchronize(self.device)
        off = (new_item.value - buf.data_ptr()) // 4
        with torch.no_grad():
            self.item_emb.set_(buf[off:off + self.I * self.D].view(self.I, self.D))
            if self.item_bias is not None:
                offb = (new_bias.value - buf.data_ptr()) // 4
                self.item_bias.set_(buf[offb:offb + self.I].view(self.item_bias.shape))
        self._symm_buf, self._symm_hdl = buf, hdl  # keep the allocation and its peer mappings alive
        self.multicast = bool(self.lib.rbpr_fused_exchange_multicast(self.ctx))
        return True

    def item_cuts(self) -> list[int]:
        """Row blocks of the item table owned by each rank under the fused exchange."""
        w = getattr(self, "world", 1)
        return [self.I * r // w for r in range(w + 1)]

    def gather_item_state(self, group=None) -> None:
        """Fused exchange + stateful optimizer: make every rank hold the whole optimizer state of the
        item table again (each rank only keeps its own row block current)."""
        if not getattr(self, "fused_exchange", False) or self.hp.optimizer == native.OPT_SGD:
            return
        from rbpr.parallel import sync_row_shards
        state = self.adam_state if self.hp.optimizer == native.OPT_ADAM else self.opt_state
        names = ("item_m", "item_v", "bias_m", "bias_v") if self.hp.optimizer == native.OPT_ADAM else ("item_s", "bias_s")
        sync_row_shards([state[k] for k in names if state.get(k) is not None], self.item_cuts(), group)

    def fused_exchange_count(self) -> int:
        return int(self.lib.rbpr_fused_exchange_count(self.ctx))

    def collective_count(self) -> int:
        return int(self.lib.rbpr_collective_count(self.ctx))

    # ---- data-parallel split ---------------------------------------------------------------
    def grad_step(self, triple_idx: torch.Tensor, seed: int, step: int,
                  neg_in: torch.Tensor | None = None, want_neg: bool = False):
        n = triple_idx.numel()
        stats = torch.empty((1, native.STATS_PER_STEP), dtype=torch.float64, device=self.device)
        neg_out = torch.empty(n, dtype=torch.int64, device=self.device) if want_neg else None
        self._check(self.lib.rbpr_grad_step(self.ctx, _ptr(triple_idx), n, seed, step,
                                            C.byref(self.hp), _ptr(neg_in), _ptr(neg_out),
                                            _ptr(stats), _stream()))
        return stats, neg_out

    def item_grad_tensor(self) -> torch.Tensor:
        """The context's dense item(+bias) gradient accumulator viewed as a torch tensor
        (for torch.distributed.all_reduce). The memory stays owned by the context."""
        ptr, numel = C.c_void_p(), C.c_int64()
        self._check(self.lib.rbpr_item_grad_buffer(self.ctx, C.byref(ptr), C.byref(numel)))
        return _wrap_device_f32(ptr.value, numel.value, self.device)

    def apply_item_grads(self, step: int) -> None:
        self._check(self.lib.rbpr_apply_item_grads(self.ctx, step, C.byref(self.hp), _stream()))

    def flush_lazy(self, step: int) -> None:
        self._check(self.lib.rbpr_flush_lazy(self.ctx, step, C.byref(self.hp), _stream()))

    # ---- scoring ---------------------------------------------------------------------------
    def score_topk(self, users: torch.Tensor, seen: tuple[torch.Tensor, torch.Tensor] | None,
                   held: tuple[torch.Tensor, torch.Tensor] | None, ks: Sequence[int],
                   k_max: int | None = None, want_items: bool = True) -> dict[str, torch.Tensor]:
        users = users.to(self.device, torch.int64).contiguous()
        n = users.numel()
        ks = [int(k) for k in ks]
        k_max = int(k_max or max(ks))
        dev = self.device
        seen_p = seen_i = held_p = held_i = None
        if seen is not None:
            seen_p, seen_i = seen[0].to(dev, torch.int64).contiguous(), seen[1].to(dev, torch.int32).contiguous()
        if held is not None:
            held_p, held_i = held[0].to(dev, torch.int64).contiguous(), held[1].to(dev, torch.int32).contiguous()
        out: dict[str, torch.Tensor] = {}
        items = scores = ndcg = recall = None
        if want_items:
            items = out["items"] = torch.empty((n, k_max), dtype=torch.int32, device=dev)
            scores = out["scores"] = torch.empty((n, k_max), dtype=torch.float32, device=dev)
        if held is not None and ks:
            ndcg = out["ndcg"] = torch.empty((n, len(ks)), dtype=torch.float32, device=dev)
            recall = out["recall"] = torch.empty((n, len(ks)), dtype=torch.float32, device=dev)
        ks_arr = (C.c_int32 * max(len(ks), 1))(*ks)
        self._check(self.lib.rbpr_score_topk(self.ctx, _ptr(users), n, _ptr(seen_p), _ptr(seen_i),
                                             _ptr(held_p), _ptr(held_i), k_max, ks_arr, len(ks),
                                             _ptr(items), _ptr(scores), _ptr(ndcg), _ptr(recall),
                                             _stream()))
        return out

    def score_metrics(self, users: torch.Tensor, seen: tuple[torch.Tensor, torch.Tensor] | None,
                      held: tuple[torch.Tensor, torch.Tensor], ks: Sequence[int], want: Sequence[str] = ("ndcg", "recall"),
                      map_normalized: bool = True, want_items: bool = False) -> dict[str, torch.Tensor]:
        """ONE scoring + ranking pass: every requested metric family (`want` from ndcg, ndcg_linear,
        recall, precision, map) at every cut-off of `ks` -> {name: (n_users, len(ks)) float32}."""
        dev = self.device
        users = users.to(dev, torch.int64).contiguous()
        n = users.numel()
        ks = [int(k) for k in ks]
        k_max = max(ks)
        if k_max > native.MAX_TOPK:
            raise ValueError(f"topk={k_max} exceeds the kernel limit {native.MAX_TOPK}")
        seen_p = seen_i = None
        if seen is not None:
            seen_p, seen_i = seen[0].to(dev, torch.int64).contiguous(), seen[1].to(dev, torch.int32).contiguous()
        held_p, held_i = held[0].to(dev, torch.int64).contiguous(), held[1].to(dev, torch.int32).contiguous()
        out = {name: torch.empty((n, len(ks)), dtype=torch.float32, device=dev) for name in want}
        mo = native.MetricOutputs()
        for name in ("ndcg", "ndcg_linear", "recall", "precision", "map"):
            setattr(mo, name, out[name].data_ptr() if name in out else None)
        if want_items:
            out["items"] = torch.empty((n, k_max), dtype=torch.int32, device=dev)
            mo.topk_items = out["items"].data_ptr()
        mo.map_normalized = int(map_normalized)
        ks_arr = (C.c_int32 * len(ks))(*ks)
        self._check(self.lib.rbpr_score_metrics(self.ctx, _ptr(users), n, _ptr(seen_p), _ptr(seen_i), _ptr(held_p),
                                                _ptr(held_i), k_max, ks_arr, len(ks), C.byref(mo), _stream()))
        return out

    def score_dense(self, users: torch.Tensor,
                    seen: tuple[torch.Tensor, torch.Tensor] | None = None) -> torch.Tensor:
        users = users.to(self.device, torch.int64).contiguous()
        out = torch.empty((users.numel(), self.I), dtype=torch.float32, device=self.device)
        seen_p = seen_i = None
        if seen is not None:
            seen_p = seen[0].to(self.device, torch.int64).contiguous()
            seen_i = seen[1].to(self.device, torch.int32).contiguous()
        self._check(self.lib.rbpr_score_dense(self.ctx, _ptr(users), users.numel(), _ptr(seen_p),
                                              _ptr(seen_i), _ptr(out), _stream()))
        return out

    # ---- instrumentation -------------------------------------------------------------------
    def kernel_timing(self, enable: bool) -> None:
        self._check(self.lib.rbpr_kernel_timing(self.ctx, int(enable)))

    def kernel_time_ms(self) -> tuple[float, int]:
        ms, n = C.c_double(), C.c_int64()
        self._check(self.lib.rbpr_kernel_time_ms(self.ctx, C.byref(ms), C.byref(n)))
        return ms.value, n.value


def _wrap_device(ptr: int, shape: tuple, typestr: str, device: torch.device) -> torch.Tensor:
    class _Arr:  # __cuda_array_interface__ carrier
        pass
    a = _Arr()
    a.__cuda_array_interface__ = {"shape": tuple(shape), "typestr": typestr, "data": (ptr, False),
                                  "version": 2, "strides": None}
    return torch.as_tensor(a, device=device)


def _wrap_device_f32(ptr: int, numel: int, device: torch.device) -> torch.Tensor:
    return _wrap_device(ptr, (numel,), "<f4", device)


def build_alias(w: np.ndarray) -> tuple[np.ndarray, np.ndarray]:
    """Walker/Vose alias table for weights w (float64, w>=0, sum>0): column k is accepted
    with probability prob[k], else alias[k] is returned.  Deterministic (stack order)."""
    n = w.size
    p = w * (n / w.sum())
    prob = np.zeros(n, dtype=np.float64)
    alias = np.zeros(n, dtype=np.int32)
    small = [i for i in range(n - 1, -1, -1) if p[i] < 1.0]
    large = [i for i in range(n - 1, -1, -1) if p[i] >= 1.0]
    while small and large:
        s, l = small.pop(), large.pop()
        prob[s], alias[s] = p[s], l
        p[l] = (p[l] + p[s]) - 1.0
        (small if p[l] < 1.0 else large).append(l)
    for i in large:
        prob[i], alias[i] = 1.0, i
    for i in small:  # numerical leftovers
        prob[i], alias[i] = 1.0, i
    # a zero-weight column must never return itself
    heavy = int(np.argmax(w))
    zero = w <= 0
    prob[zero] = 0.0
    alias[zero & (alias == np.arange(n))] = heavy
    return prob.astype(np.float32), alias
