"""`Model` (exported as `revisit_bpr.models.BPR`) and `MF` with the reference's constructor
signatures, parameter names, output keys and error types (reference:
revisit_bpr/models/bpr/model.py:13-153), executed by librbpr.so.

What changes underneath, and nothing else:

* train-mode `Model.forward(batch)` runs the WHOLE step — logits, loss, L2, exact minibatch
  gradients and the optimizer update of the touched rows — in the fused CUDA path
  (rbpr_train_step_triples).  The returned dict has the reference's keys (`logits_pos`,
  `logits_neg`, `logits`, `bpr_loss`, `l2_reg`, `loss`); `loss.backward()`, `optimizer.step()` and
  `optimizer.zero_grad()` stay legal and become no-ops (no `.grad` is ever materialised), so loops
  written like the reference's example.py:172-180 / experiments/trainer.py:64-83 run unchanged
  once `model.bind_optimizer(optimizer)` has been called (our Trainer does it).
* eval-mode `Model.forward(batch)` returns `{"logits": ...}` from the pair-logits kernel.
* `MF` keeps `nn.Embedding` tables and `nn.Parameter` biases, so `state_dict()`, `parameters()`,
  `get_features()` and checkpoints are interchangeable with the reference's.

There is no CPU path: calling forward with CPU tensors raises.
"""
from __future__ import annotations

from typing import Any

import torch

from rbpr import native
from rbpr.engine import Engine
from revisit_bpr.models.bpr.loss import Loss


def _f32(x: float) -> float:
    import ctypes
    return ctypes.c_float(float(x)).value


class BaseLogitModel(torch.nn.Module):
    def get_features(self) -> dict[str, torch.Tensor]:
        return {}


class MF(BaseLogitModel):
    """Matrix-factorisation logits: <user row, item row> (+ item bias) (+ user bias)."""

    def __init__(self, user_emb: torch.nn.Embedding, item_emb: torch.nn.Embedding,
                 item_bias: bool = False, user_bias: bool = False) -> None:
        super().__init__()
        if user_emb.embedding_dim != item_emb.embedding_dim:
            raise ValueError("user and item embeddings must share embedding_dim")
        self._user_emb = user_emb
        self._item_emb = item_emb
        if item_bias:
            self._item_bias = torch.nn.Parameter(torch.empty(item_emb.num_embeddings))
        else:
            self.register_parameter("_item_bias", None)
        if user_bias:
            self._user_bias = torch.nn.Parameter(torch.empty(user_emb.num_embeddings))
        else:
            self.register_parameter("_user_bias", None)
        self._engine: Engine | None = None
        self._engine_key: tuple | None = None
        self.reset_parameters()

    @torch.no_grad()
    def reset_parameters(self) -> None:
        # U(-0.5, 0.5) / dim for both tables, padding row zeroed, biases zero
        # (reference model.py:117-129; same RNG consumption order: user table first)
        for emb in (self._user_emb, self._item_emb):
            emb.weight.uniform_().sub_(0.5).div_(emb.embedding_dim)
            if emb.padding_idx is not None:
                emb.weight[emb.padding_idx].zero_()
        for bias in (self._item_bias, self._user_bias):
            if bias is not None:
                bias.zero_()

    def get_features(self) -> dict[str, torch.Tensor]:
        return {"user": self._user_emb.weight, "item": self._item_emb.weight,
                "user_bias": self._user_bias, "item_bias": self._item_bias}

    # ---- native context bound to the parameter storages --------------------------------------
    def engine(self) -> Engine:
        uw, iw, ib = self._user_emb.weight, self._item_emb.weight, self._item_bias
        if not uw.is_cuda:
            raise native.NativeError(
                "revisit_bpr.models.bpr.MF runs on a B200 only (librbpr.so, sm_100a): move the model "
                "to CUDA; there is no CPU fallback")
        for emb in (self._user_emb, self._item_emb):
            if emb.padding_idx != 0:
                # the kernels hard-wire row 0 as the padding row (never sampled, never updated, always
                # masked), which is what every reference config sets; with padding_idx=None the
                # reference would TRAIN row 0 — refusing beats silently freezing it
                raise NotImplementedError("the CUDA BPR path needs nn.Embedding(..., padding_idx=0) on both tables "
                                          f"(got padding_idx={emb.padding_idx})")
        key = (uw.data_ptr(), iw.data_ptr(), None if ib is None else ib.data_ptr(), tuple(uw.shape),
               tuple(iw.shape))
        if self._engine is None or self._engine_key != key:
            self._engine = Engine(uw.data, iw.data, None if ib is None else ib.data)
            self._engine_key = key
        return self._engine

    @torch.no_grad()
    def adopt_engine_tables(self) -> None:
        """The symmetric-memory binding of the fused exchange moves the item table / bias into a buffer
        every peer maps (Engine.init_fused_exchange): re-point the parameters at the engine's tensors
        so that module, optimizer and checkpoints keep seeing the live storage."""
        eng = self._engine
        if eng is None:
            return
        iw, ib = self._item_emb.weight, self._item_bias
        if iw.data_ptr() != eng.item_emb.data_ptr():
            iw.data = eng.item_emb
        if ib is not None and eng.item_bias is not None and ib.data_ptr() != eng.item_bias.data_ptr():
            ib.data = eng.item_bias
        uw = self._user_emb.weight
        self._engine_key = (uw.data_ptr(), iw.data_ptr(), None if ib is None else ib.data_ptr(), tuple(uw.shape),
                            tuple(iw.shape))

    @torch.no_grad()
    def forward(self, user: torch.Tensor, item: torch.Tensor, _: dict[str, torch.Tensor] | None = None,
                mask: torch.Tensor | None = None) -> torch.Tensor:
        # user (B,), item (B, ...) -> logits (B, ...)
        if user.dim() != 1 or item.size(0) != user.size(0):
            raise IndexError(f"user must be (batch,), item (batch, ...): got {tuple(user.shape)}, {tuple(item.shape)}")
        eng = self.engine()
        ub = None if self._user_bias is None else self._user_bias.data
        return eng.pair_logits(user, item, mask, ub)


class AllItemsEval(dict):
    """Eval-mode output of `Model.forward` for a batch that scores EVERY item for each user (the
    `AllItemsCollator` batches of the reference, experiments/bpr/dataset.py:274-296).

    It behaves like the reference's `{"logits": (B, I)}` dict, but the matrix is only built when
    somebody reads `["logits"]` (one fp32 scoring kernel, rbpr_score_dense).  The metric hook of
    `experiments.options.attach_metrics` does not: it asks `ranking_metrics` for NDCG / Recall /
    Precision / MAP at every cut-off in ONE fused scoring + ranking call (rbpr_score_metrics), which is
    what replaces reference model.py:43-47 + exp.py:369-374 + one full sort per metric object."""

    def __init__(self, model: "Model", users: torch.Tensor) -> None:
        super().__init__()
        self._model, self._users = model, users
        self._seen: tuple[torch.Tensor, torch.Tensor] | None = None
        self.masked = False

    @property
    def fused(self) -> bool:
        """True until the dense logits have been materialised (after that they are authoritative:
        a handler may have edited them in place)."""
        return not dict.__contains__(self, "logits")

    def mask_seen(self, seen_csr: tuple[torch.Tensor, torch.Tensor]) -> None:
        """`_remove_seen_items` (reference exp.py:369-374) for the fused path: remember the rows to
        push to -1e13 (plus column 0) instead of editing a matrix."""
        self._seen, self.masked = seen_csr, True

    def ranking_metrics(self, held_csr: tuple[torch.Tensor, torch.Tensor], ks: list[int], want: tuple[str, ...],
                        map_normalized: bool = True) -> dict[str, torch.Tensor]:
        eng = self._model.logits_model.engine()
        return eng.score_metrics(self._users, self._seen, held_csr, ks, want=want, map_normalized=map_normalized)

    def __missing__(self, key: str) -> torch.Tensor:
        if key != "logits":
            raise KeyError(key)
        lm = self._model.logits_model
        eng = lm.engine()
        if self.masked:
            logits = eng.score_dense(self._users, self._seen)
        else:  # nothing masked (skip_seen=False): column 0 keeps its real score, as in the reference
            items = torch.arange(eng.I, device=eng.device).unsqueeze(0).expand(self._users.numel(), -1)
            logits = lm(self._users.to(eng.device), items.contiguous())
        self["logits"] = logits
        return logits

    def __contains__(self, key: object) -> bool:
        return key == "logits" or dict.__contains__(self, key)

    def get(self, key: str, default: Any = None) -> Any:
        try:
            return self[key]
        except KeyError:
            return default


class _AppliedStep(torch.autograd.Function):
    """Gives the already-applied loss a grad_fn so that `loss.backward()` stays legal."""

    @staticmethod
    def forward(ctx: Any, anchor: torch.Tensor, value: torch.Tensor) -> torch.Tensor:  # noqa: ARG004
        return value.clone()

    @staticmethod
    def backward(ctx: Any, grad: torch.Tensor):  # noqa: ARG004
        return None, None


class Model(torch.nn.Module):
    """The BPR model: `Model(logits_model, reg_alphas=None, fuse_forward=False)`.

    reg_alphas keys: user, item, neg, all (`all` overrides; `neg` defaults to `item`) — reference
    model.py:70-86.  `fuse_forward` is accepted for signature compatibility: the CUDA path always
    evaluates both logits of a triple in one pass.
    """

    def __init__(self, logits_model: BaseLogitModel, reg_alphas: dict[str, float] | None = None,
                 fuse_forward: bool = False) -> None:
        super().__init__()
        self.logits_model = logits_model
        self._reg_alphas = reg_alphas or {}
        self._fuse_forward = fuse_forward
        self._loss = Loss(size_average=False)
        self._optimizer: torch.optim.Optimizer | None = None
        self._opt_kind: int | None = None
        self._opt_step = 0
        self._adam_last: torch.Tensor | None = None
        self._anchor = torch.zeros((), requires_grad=True)

    # ---- optimizer binding ---------------------------------------------------------------------
    def bind_optimizer(self, optimizer: torch.optim.Optimizer) -> None:
        """Tell the fused step which torch optimizer it stands in for (hyper-parameters are read
        from its param_groups at every step, so LR schedulers keep working; Adam moments live in
        `optimizer.state`, so `optimizer.state_dict()` stays meaningful)."""
        opt = getattr(optimizer, "optimizer", optimizer)  # accelerate's AcceleratedOptimizer wrapper
        if not isinstance(self.logits_model, MF):
            return  # ItemKNN / FreeItemKNN: kernel-backed autograd, the optimizer does its own step
        if len(opt.param_groups) != 1:
            raise NotImplementedError("the fused BPR step supports a single param group")
        g = opt.param_groups[0]
        if g.get("weight_decay", 0) != 0 or g.get("maximize", False):
            raise NotImplementedError("weight_decay / maximize are not supported by the fused BPR step")
        if isinstance(opt, torch.optim.SGD):
            if g.get("dampening", 0) != 0:
                raise NotImplementedError("SGD dampening is not supported by the fused BPR step")
            self._opt_kind = native.OPT_SGDM if g.get("momentum", 0) != 0 else native.OPT_SGD
        elif isinstance(opt, torch.optim.RMSprop):
            if g.get("momentum", 0) != 0 or g.get("centered", False):
                raise NotImplementedError("RMSprop momentum / centered are not supported by the fused BPR step")
            self._opt_kind = native.OPT_RMSPROP
        elif isinstance(opt, torch.optim.Adam) and not isinstance(opt, torch.optim.AdamW):
            if g.get("amsgrad", False):
                raise NotImplementedError("amsgrad is not supported by the fused BPR step")
            self._opt_kind = native.OPT_ADAM
        else:
            raise NotImplementedError(f"{type(opt).__name__} is not implemented in the fused BPR step "
                                      "(supported: torch.optim.SGD [momentum, nesterov], Adam, RMSprop [momentum=0])")
        self._optimizer = opt
        self._resync_step()
        if hasattr(opt, "register_state_dict_pre_hook"):
            opt.register_state_dict_pre_hook(lambda _opt: self.flush())
        if hasattr(opt, "register_load_state_dict_post_hook"):
            # optimizer.load_state_dict AFTER binding (accelerate.load_state on a built trainer): the
            # state tensors were replaced and every row is again consistent with the loaded step
            opt.register_load_state_dict_post_hook(lambda _opt: self._resync_step())

    def _resync_step(self) -> None:
        steps = [int(s["step"]) for s in self._optimizer.state.values() if "step" in s]
        self._opt_step = max(steps) if steps else 0
        self._adam_last = None

    def restore_step(self, step: int) -> None:
        """Set the number of optimizer steps already applied (checkpoint resume: plain SGD keeps no
        state that carries it; it numbers the fast path's sampler stream and Adam's bias correction)."""
        if self._adam_last is not None and step != self._opt_step:
            raise RuntimeError("restore_step after training started with lazily updated rows in flight")
        self._opt_step = int(step)

    def _adam_state(self, eng: Engine) -> dict[str, torch.Tensor]:
        opt = self._optimizer
        feats = self.logits_model.get_features()

        def moments(p: torch.Tensor) -> tuple[torch.Tensor, torch.Tensor]:
            st = opt.state[p]
            if "exp_avg" not in st:
                st["step"] = torch.tensor(float(self._opt_step))
                st["exp_avg"] = torch.zeros_like(p, memory_format=torch.preserve_format)
                st["exp_avg_sq"] = torch.zeros_like(p, memory_format=torch.preserve_format)
            return st["exp_avg"], st["exp_avg_sq"]

        um, uv = moments(feats["user"])
        im, iv = moments(feats["item"])
        state = {"user_m": um, "user_v": uv, "item_m": im, "item_v": iv}
        if feats["item_bias"] is not None:
            state["bias_m"], state["bias_v"] = moments(feats["item_bias"])
        if self._adam_last is None or self._adam_last.device != eng.device:
            # every row is consistent with `_opt_step` dense steps when binding (fresh or resumed)
            self._adam_last = torch.full((eng.U,), self._opt_step, dtype=torch.int32, device=eng.device)
        state["user_last"] = self._adam_last
        return state

    def _state1(self, eng: Engine, key: str, with_step: bool) -> dict[str, torch.Tensor]:
        """Single-state optimizers: the state tensors live in `optimizer.state` under torch's own
        names (`momentum_buffer` / `square_avg`), so optimizer.state_dict() stays interchangeable."""
        opt = self._optimizer
        feats = self.logits_model.get_features()

        def slot(p: torch.Tensor) -> torch.Tensor:
            st = opt.state[p]
            if st.get(key) is None:
                st[key] = torch.zeros_like(p, memory_format=torch.preserve_format)
                if with_step:
                    st["step"] = torch.tensor(float(self._opt_step))
            return st[key]

        state = {"user_s": slot(feats["user"]), "item_s": slot(feats["item"])}
        if feats["item_bias"] is not None:
            state["bias_s"] = slot(feats["item_bias"])
        if self._adam_last is None or self._adam_last.device != eng.device:
            self._adam_last = torch.full((eng.U,), self._opt_step, dtype=torch.int32, device=eng.device)
        state["user_last"] = self._adam_last
        return state

    def _configure(self, eng: Engine) -> None:
        if self._optimizer is None:
            raise RuntimeError(
                "train-mode forward of the CUDA BPR model applies the optimizer update inside the fused "
                "kernel: call model.bind_optimizer(optimizer) once after creating the optimizer "
                "(experiments.trainer.Trainer does this for you)")
        g = self._optimizer.param_groups[0]
        eng.set_reg(self._reg_alphas)
        if self._opt_kind == native.OPT_SGD:
            eng.set_sgd(float(g["lr"]))
        elif self._opt_kind == native.OPT_SGDM:
            # the zero-gradient catch-up of lazily updated user rows replays the missed steps with the
            # learning rate in force NOW (only Adam keeps a per-step table): when a scheduler moved lr,
            # bring every row up to date with the old value first
            if self._adam_last is not None and eng.hp.optimizer == native.OPT_SGDM and eng.hp.lr != 0.0 \
                    and eng.hp.lr != _f32(g["lr"]):
                eng.flush_lazy(self._opt_step)
            eng.set_sgd_momentum(float(g["lr"]), float(g["momentum"]), bool(g.get("nesterov", False)),
                                 state=self._state1(eng, "momentum_buffer", with_step=False))
        elif self._opt_kind == native.OPT_RMSPROP:
            eng.set_rmsprop(float(g["lr"]), float(g["alpha"]), float(g["eps"]),
                            state=self._state1(eng, "square_avg", with_step=True))
        else:
            eng.set_adam(float(g["lr"]), tuple(g["betas"]), float(g["eps"]), state=self._adam_state(eng))

    @torch.no_grad()
    def flush(self) -> None:
        """Materialise lazily-deferred optimizer work (dense-Adam catch-up of user rows) and mirror
        the step counter into `optimizer.state`.  Called before eval, state_dict and checkpoints."""
        lm = self.logits_model
        if self._optimizer is None or self._opt_kind in (None, native.OPT_SGD) or not isinstance(lm, MF):
            return
        if lm._engine is None or self._adam_last is None:
            return
        eng = lm.engine()
        self._configure(eng)
        eng.flush_lazy(self._opt_step)
        for st in self._optimizer.state.values():
            if "step" in st:
                st["step"].fill_(float(self._opt_step))

    def state_dict(self, *args: Any, **kwargs: Any):  # noqa: ANN201
        self.flush()
        return super().state_dict(*args, **kwargs)

    # ---- data parallel (one process per GPU) ------------------------------------------------------
    def enable_data_parallel(self, user_cuts: Any, group: Any = None) -> None:
        """Join the library's NCCL communicator: from now on every training step of this model ends in
        ONE all-reduce of the dense item gradient (rbpr_train_steps / rbpr_train_step_triples), so all
        ranks must run the same number of steps.  `user_cuts` (world+1,) are the owner blocks of user
        rows (rbpr.parallel.shard_bounds): each rank must only be fed triples of its own users.
        Replaces Distributed(...) + launcher.DDP + accelerator.prepare of the reference
        (experiments/decorator.py:30-54, launcher.py:35-73, bpr/exp.py:101-105)."""
        import numpy as np
        eng = self.logits_model.engine()
        if getattr(eng, "world", 1) <= 1:
            eng.init_comm(group)
            import os
            if os.environ.get("RBPR_FUSED_EXCHANGE", "1") != "0":
                # the exchange as one kernel over NVLink peer memory; NCCL stays when it is unavailable
                eng.init_fused_exchange(group)
                self.logits_model.adopt_engine_tables()
        self._dp = {"cuts": np.asarray(user_cuts, dtype=np.int64), "group": group}

    @torch.no_grad()
    def sync_user_shards(self) -> None:
        """All-gather the owner-sharded user rows (and their optimizer state) so that every rank can
        evaluate any user and rank 0 can write a complete checkpoint."""
        dp = getattr(self, "_dp", None)
        if dp is None:
            return
        from rbpr.parallel import sync_row_shards
        self.flush()
        feats = self.logits_model.get_features()
        tensors = [feats["user"].data]
        if self._optimizer is not None:
            st = self._optimizer.state.get(feats["user"], {})
            tensors += [st[k] for k in ("exp_avg", "exp_avg_sq", "momentum_buffer", "square_avg")
                        if torch.is_tensor(st.get(k))]
        sync_row_shards(tensors, dp["cuts"], dp["group"])
        lm = self.logits_model
        if lm._engine is not None and self._optimizer is not None and self._opt_kind not in (None, native.OPT_SGD):
            # fused exchange: the item table's optimizer state is sharded by row block between eval passes
            lm._engine.gather_item_state(dp["group"])

    # ---- forward ---------------------------------------------------------------------------------
    def forward(self, inputs: dict[str, torch.Tensor]) -> dict[str, torch.Tensor]:
        lm = self.logits_model
        if not isinstance(lm, MF):
            return self._forward_unfused(inputs)
        if not self.training:
            self.flush()
            if inputs.get("all_items") is True and lm._user_bias is None and inputs.get("mask") is None:
                return AllItemsEval(self, inputs["user"])
            return {"logits": lm(inputs["user"], inputs["item"], inputs, mask=inputs.get("mask"))}
        if "triple_idx" in inputs:
            return self._forward_triple_ids(inputs)
        user, item, neg = inputs["user"], inputs["item"], inputs["neg"]
        if item.dim() < 2:
            item = item.unsqueeze(-1)
        if neg.dim() < 2:
            neg = neg.unsqueeze(-1)
        if item.shape != neg.shape or item.dim() != 2:
            raise IndexError(f"item and neg must both be (batch, num items): got {tuple(item.shape)}, {tuple(neg.shape)}")
        if not (user.size(0) == item.size(0) == neg.size(0)):
            raise IndexError("user, item and neg must share the batch dimension")
        eng = lm.engine()
        self._configure(eng)
        width = item.size(-1)
        users = user
        if width > 1:
            # several (positive, negative) pairs per row (reference model.py:41-42,48-57: logits are
            # (batch, num items), the loss sums over all of them): B*K triples of the same fused step.
            # The reference counts the user's L2 term once per ROW, the kernel once per triple, so the
            # user coefficient is divided by K for this call.
            users = user.repeat_interleave(width)
            eng.hp.reg_user = eng.hp.reg_user / width
        logits, stats = eng.train_step_triples(users, item.reshape(-1), neg.reshape(-1), self._opt_step)
        self._opt_step += 1
        stats32 = stats.to(torch.float32)
        pos, ng = logits[:, 0].reshape(-1, width), logits[:, 1].reshape(-1, width)
        if lm._user_bias is not None:
            # the user bias cancels in pos - neg (zero BPR gradient, so it never trains); it only
            # shifts the two reported logits, like MF.forward does (reference model.py:139-144)
            ub = lm._user_bias.detach()[user.to(lm._user_bias.device)].unsqueeze(-1)
            pos, ng = pos + ub, ng + ub
        if self._anchor.device != stats32.device:
            self._anchor = torch.zeros((), requires_grad=True, device=stats32.device)
        bpr_loss, l2_reg = stats32[0], stats32[1]
        return {"logits_pos": pos, "logits_neg": ng, "logits": pos - ng, "bpr_loss": bpr_loss,
                "l2_reg": l2_reg, "loss": _AppliedStep.apply(self._anchor, bpr_loss + l2_reg)}

    def _forward_unfused(self, inputs: dict[str, torch.Tensor]) -> dict[str, torch.Tensor]:
        """Logits models other than MF (ItemKNN, FreeItemKNN — reference model.py:156-251): their
        forward/backward are CUDA kernels behind autograd, the loss and L2 terms of model.py:43-68 are
        elementwise torch on (B,1) tensors, and the caller's `loss.backward()` / `optimizer.step()`
        do the update, as in the reference."""
        lm = self.logits_model
        user, item = inputs["user"], inputs["item"]
        if not self.training:
            logits = lm(user, item, inputs)
            if (mask := inputs.get("mask")) is not None:
                logits = logits.masked_fill(mask.to(logits.device) == 0, -1e13)
            return {"logits": logits}
        neg = inputs["neg"]
        if self._fuse_forward:
            width = item.size(-1)
            both = lm(user, torch.cat((item, neg), dim=-1), inputs)
            pos, ng = both[:, :width], both[:, width:]
        else:
            pos, ng = lm(user, item, inputs), lm(user, neg, inputs)
        out = {"logits_pos": pos, "logits_neg": ng, "logits": pos - ng}
        out["bpr_loss"] = self._loss(out["logits"]).sum()
        out["l2_reg"] = self._l2_rows(inputs).sum()
        out["loss"] = out["bpr_loss"] + out["l2_reg"]
        return out

    # ---- fast path: whole runs of steps from triple ids (our extension, not in the reference) -----
    def bind_interactions(self, indptr: torch.Tensor, indices: torch.Tensor, sampler: int = native.SAMPLER_UNIFORM,
                          seed: int = 13, item_weights: torch.Tensor | None = None,
                          adaptive_prob: float | None = None, adaptive_every: int = 0) -> None:
        """Give the model the training interaction matrix (CSR) and the negative-sampler settings, so
        that `forward({"triple_idx": ids, "batch_size": B})` can run ceil(len(ids)/B) complete training
        steps (device sampling, update) in ONE library call (rbpr_train_steps)."""
        eng = self.logits_model.engine()
        eng.bind_csr(indptr, indices)
        if item_weights is not None:
            eng.bind_item_weights(item_weights)
        self._fast = {"sampler": sampler, "seed": int(seed), "adaptive_prob": adaptive_prob,
                      "adaptive_every": int(adaptive_every), "stats_ready": False}

    def _forward_triple_ids(self, inputs: dict[str, torch.Tensor]) -> dict[str, torch.Tensor]:
        fast = getattr(self, "_fast", None)
        if fast is None:
            raise RuntimeError("forward with 'triple_idx' needs model.bind_interactions(indptr, indices) first")
        eng = self.logits_model.engine()
        self._configure(eng)
        if fast["sampler"] == native.SAMPLER_ADAPTIVE:
            eng.set_adaptive(fast["adaptive_prob"], fast["adaptive_every"])
            if not fast["stats_ready"]:
                eng.adaptive_update_stats()
                fast["stats_ready"] = True
        else:
            eng.set_sampler(fast["sampler"])
        ids = inputs["triple_idx"].to(eng.device, torch.int64).contiguous()
        batch = int(inputs["batch_size"])
        stats, _ = eng.train_steps(ids, batch, fast["seed"], self._opt_step)
        steps = stats.size(0)
        self._opt_step += steps
        st = stats.to(torch.float32)
        if self._anchor.device != st.device:
            self._anchor = torch.zeros((), requires_grad=True, device=st.device)
        # per-step means over the call, so running means over "iterations" keep the per-step scale
        bpr_loss, l2_reg = st[:, 0].mean(), st[:, 1].mean()
        return {"bpr_loss": bpr_loss, "l2_reg": l2_reg, "loss": _AppliedStep.apply(self._anchor, bpr_loss + l2_reg),
                "logits": (st[:, 2].sum() / st[:, 3].sum()).reshape(1, 1), "steps": steps, "step_stats": stats}

    def regularization(self, inputs: dict[str, torch.Tensor]) -> torch.Tensor:
        """Per-row L2 term (B,) of the reference (model.py:70-93) for callers that want it on its
        own; the MF training path computes it inside the fused kernel (so no graph is kept there)."""
        if isinstance(self.logits_model, MF):
            with torch.no_grad():
                return self._l2_rows(inputs)
        return self._l2_rows(inputs)

    def _l2_rows(self, inputs: dict[str, torch.Tensor]) -> torch.Tensor:
        from rbpr.engine import resolve_reg
        feats = self.logits_model.get_features()
        if not feats or all(self._reg_alphas.get(k) is None for k in ("all", "user", "item", "neg")):
            return torch.tensor(0)
        ru, ri, rn = resolve_reg(self._reg_alphas)
        term = (ri * feats["item"][inputs["item"]].pow(2).flatten(1).sum(1)
                + rn * feats["item"][inputs["neg"]].pow(2).flatten(1).sum(1))
        if feats.get("user") is not None:
            term = term + ru * feats["user"][inputs["user"]].pow(2).flatten(1).sum(1)
        return term / 2


def __getattr__(name: str) -> Any:
    # the reference defines ItemKNN / FreeItemKNN in this module (model.py:156-251); ours live in knn.py
    if name in ("ItemKNN", "FreeItemKNN"):
        from revisit_bpr.models.bpr import knn
        return getattr(knn, name)
    raise AttributeError(f"module {__name__!r} has no attribute {name!r}")
