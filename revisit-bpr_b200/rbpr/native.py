"""ctypes binding of include/rbpr.h.  Loads the in-tree librbpr.so; raises if it is absent."""
from __future__ import annotations

import ctypes as C
import os
from pathlib import Path

LIB_PATH = Path(__file__).resolve().parent.parent / "librbpr.so"

OPT_SGD, OPT_ADAM, OPT_SGDM, OPT_RMSPROP = 0, 1, 2, 3
SAMPLER_UNIFORM, SAMPLER_WEIGHTED, SAMPLER_INJECTED, SAMPLER_ADAPTIVE = 0, 1, 2, 3
STATS_PER_STEP = 4
MAX_TOPK = 128
ABI_VERSION = 9
IPC_BLOB_BYTES = 512
MAX_PEERS = 8

# Every symbol include/rbpr.h declares (tests check the library exports all of them).
SYMBOLS = (
    "rbpr_abi_version", "rbpr_create", "rbpr_destroy", "rbpr_last_error", "rbpr_bind_tables",
    "rbpr_bind_adam_state", "rbpr_bind_state1", "rbpr_bind_csr", "rbpr_bind_item_alias", "rbpr_adaptive_update_stats", "rbpr_adaptive_stats",
    "rbpr_sample_adaptive_padded", "rbpr_sample_negatives",
    "rbpr_train_steps", "rbpr_sync_check", "rbpr_train_steps_host", "rbpr_grad_step",
    "rbpr_item_grad_buffer", "rbpr_apply_item_grads", "rbpr_flush_lazy", "rbpr_score_topk",
    "rbpr_score_dense", "rbpr_train_step_triples", "rbpr_pair_logits", "rbpr_sample_negatives_padded",
    "rbpr_topk_metrics_dense", "rbpr_mask_seen_padded", "rbpr_auc_dense",
    "rbpr_knn_forward", "rbpr_knn_backward", "rbpr_freeknn_forward", "rbpr_freeknn_backward",
    "rbpr_comm_unique_id", "rbpr_comm_init", "rbpr_comm_allreduce_item_grads", "rbpr_collective_count",
    "rbpr_ingest_pairs", "rbpr_ingest_lists", "rbpr_ingest_free", "rbpr_ingest_last_error",
    "rbpr_launch_count", "rbpr_kernel_timing", "rbpr_kernel_time_ms",
    "rbpr_score_metrics", "rbpr_topk_launch_count",
    "rbpr_score_path_counts", "rbpr_comm_ipc_export", "rbpr_comm_ipc_bind", "rbpr_fused_exchange_count",
    "rbpr_comm_symm_bytes", "rbpr_comm_symm_bind", "rbpr_fused_exchange_multicast",
)


class HParams(C.Structure):
    _fields_ = [
        ("optimizer", C.c_int32), ("sampler", C.c_int32), ("lr", C.c_float),
        ("beta1", C.c_float), ("beta2", C.c_float), ("eps", C.c_float),
        ("reg_user", C.c_float), ("reg_item", C.c_float), ("reg_neg", C.c_float),
        ("adaptive_prob", C.c_float), ("adaptive_every", C.c_int32), ("reserved0", C.c_int32),
    ]


class MetricOutputs(C.Structure):
    _fields_ = [
        ("ndcg", C.c_void_p), ("ndcg_linear", C.c_void_p), ("recall", C.c_void_p), ("precision", C.c_void_p),
        ("map", C.c_void_p), ("topk_items", C.c_void_p), ("map_normalized", C.c_int32), ("reserved0", C.c_int32),
    ]


class NativeError(RuntimeError):
    pass


_lib = None


def load() -> C.CDLL:
    """Load librbpr.so (once).  No fallback: a missing library is an error."""
    global _lib
    if _lib is not None:
        return _lib
    path = Path(os.environ.get("RBPR_LIB", LIB_PATH))
    if not path.exists():
        raise NativeError(
            f"{path} not found: build it with `python -c 'import __graft_entry__ as g; g.build()'` "
            "(nvcc, sm_100a). There is no CPU or PyTorch fallback for the BPR hot path."
        )
    lib = C.CDLL(str(path))
    vp, i64, i32, u64 = C.c_void_p, C.c_int64, C.c_int32, C.c_uint64
    hp = C.POINTER(HParams)
    sig = {
        "rbpr_abi_version": (C.c_int, []),
        "rbpr_create": (C.c_int, [C.c_int, C.POINTER(vp)]),
        "rbpr_destroy": (None, [vp]),
        "rbpr_last_error": (C.c_char_p, [vp]),
        "rbpr_bind_tables": (C.c_int, [vp, vp, i64, vp, i64, i32, vp]),
        "rbpr_bind_adam_state": (C.c_int, [vp] * 8),
        "rbpr_bind_state1": (C.c_int, [vp] * 5),
        "rbpr_bind_csr": (C.c_int, [vp, vp, vp, i64, i64, vp]),
        "rbpr_bind_item_alias": (C.c_int, [vp, vp, vp]),
        "rbpr_adaptive_update_stats": (C.c_int, [vp, vp]),
        "rbpr_adaptive_stats": (C.c_int, [vp, C.POINTER(vp), C.POINTER(vp), C.POINTER(vp)]),
        "rbpr_sample_adaptive_padded": (C.c_int, [vp, vp, vp, i64, i64, i64, C.c_double, u64, u64, vp, u64, hp, vp]),
        "rbpr_sample_negatives": (C.c_int, [vp, vp, i64, u64, u64, i32, vp, vp]),
        "rbpr_train_steps": (C.c_int, [vp, vp, i64, i64, u64, u64, hp, vp, vp, vp, vp]),
        "rbpr_sync_check": (C.c_int, [vp, vp]),
        "rbpr_train_steps_host": (C.c_int, [vp, vp, i64, i64, u64, u64, hp, vp, vp, vp, vp]),
        "rbpr_grad_step": (C.c_int, [vp, vp, i64, u64, u64, hp, vp, vp, vp, vp]),
        "rbpr_item_grad_buffer": (C.c_int, [vp, C.POINTER(vp), C.POINTER(i64)]),
        "rbpr_apply_item_grads": (C.c_int, [vp, u64, hp, vp]),
        "rbpr_flush_lazy": (C.c_int, [vp, u64, hp, vp]),
        "rbpr_score_topk": (C.c_int, [vp, vp, i64, vp, vp, vp, vp, i32, C.POINTER(i32), i32,
                                      vp, vp, vp, vp, vp]),
        "rbpr_score_dense": (C.c_int, [vp, vp, i64, vp, vp, vp, vp]),
        "rbpr_score_metrics": (C.c_int, [vp, vp, i64, vp, vp, vp, vp, i32, C.POINTER(i32), i32,
                                         C.POINTER(MetricOutputs), vp]),
        "rbpr_topk_launch_count": (i64, [vp]),
        "rbpr_score_path_counts": (C.c_int, [vp, C.POINTER(i64), C.POINTER(i64)]),
        "rbpr_comm_ipc_export": (C.c_int, [vp, vp]),
        "rbpr_comm_ipc_bind": (C.c_int, [vp, vp, i32, i32, vp]),
        "rbpr_fused_exchange_count": (i64, [vp]),
        "rbpr_comm_symm_bytes": (i64, [vp]),
        "rbpr_comm_symm_bind": (C.c_int, [vp, C.POINTER(u64), u64, i32, i32, C.POINTER(u64), C.POINTER(u64), vp]),
        "rbpr_fused_exchange_multicast": (i32, [vp]),
        "rbpr_train_step_triples": (C.c_int, [vp, vp, vp, vp, i64, u64, hp, vp, vp, vp]),
        "rbpr_pair_logits": (C.c_int, [vp, vp, vp, vp, i64, i64, vp, vp, vp]),
        "rbpr_sample_negatives_padded": (C.c_int, [vp, vp, i64, i64, i64, i64, u64, u64, i32, vp, vp]),
        "rbpr_topk_metrics_dense": (C.c_int, [vp, vp, vp, i64, i64, i32, C.POINTER(i32), i32, i32,
                                              vp, vp, vp, vp, i32, vp, vp]),
        "rbpr_mask_seen_padded": (C.c_int, [vp, vp, vp, i64, i64, i64, vp]),
        "rbpr_auc_dense": (C.c_int, [vp, vp, vp, vp, i64, i64, vp, vp]),
        "rbpr_knn_forward": (C.c_int, [vp, vp, i64, i32, vp, vp, i64, i64, vp, i64, vp, vp, vp, vp]),
        "rbpr_knn_backward": (C.c_int, [vp, vp, i64, i32, vp, i64, i64, vp, i64, vp, vp, vp, vp, vp, vp]),
        "rbpr_freeknn_forward": (C.c_int, [vp, vp, i64, vp, vp, i64, i64, vp, i64, vp, vp, vp]),
        "rbpr_freeknn_backward": (C.c_int, [vp, i64, vp, i64, i64, vp, i64, vp, vp, vp, vp, vp]),
        "rbpr_comm_unique_id": (C.c_int, [vp, vp]),
        "rbpr_comm_init": (C.c_int, [vp, i32, i32, vp]),
        "rbpr_comm_allreduce_item_grads": (C.c_int, [vp, vp]),
        "rbpr_collective_count": (i64, [vp]),
        "rbpr_ingest_pairs": (C.c_int, [C.c_char_p, C.c_char_p, C.c_char_p, C.POINTER(vp), C.POINTER(vp),
                                        C.POINTER(i64)]),
        "rbpr_ingest_lists": (C.c_int, [C.c_char_p, C.c_char_p, C.c_char_p, C.POINTER(vp), C.POINTER(vp),
                                        C.POINTER(vp), C.POINTER(i64)]),
        "rbpr_ingest_free": (None, [vp]),
        "rbpr_ingest_last_error": (C.c_char_p, []),
        "rbpr_launch_count": (i64, [vp]),
        "rbpr_kernel_timing": (C.c_int, [vp, i32]),
        "rbpr_kernel_time_ms": (C.c_int, [vp, C.POINTER(C.c_double), C.POINTER(i64)]),
    }
    for name, (res, args) in sig.items():
        fn = getattr(lib, name)
        fn.restype = res
        fn.argtypes = args
    if lib.rbpr_abi_version() != ABI_VERSION:
        raise NativeError(f"librbpr.so ABI {lib.rbpr_abi_version()} != binding {ABI_VERSION}")
    _lib = lib
    return lib


def check(lib: C.CDLL, ctx, rc: int) -> None:
    if rc != 0:
        msg = lib.rbpr_last_error(ctx)
        raise NativeError(f"librbpr error {rc}: {msg.decode() if msg else '?'}")
