"""Host-side pieces of the fused eval path and of the data-parallel experiment wiring (no GPU)."""
import numpy as np
import torch

from experiments.bpr.dataset import AllItemsBatch, AllItemsCollator, EpochChunks


def test_all_items_batch_is_the_reference_batch_built_lazily():
    insts = [{"user": 3, "item": [5, 2, 5], "seen_items": [1, 7]},
             {"user": 9, "item": [], "seen_items": [4]},
             {"user": 1, "item": [8], "seen_items": [2, 3, 6]}]
    b = AllItemsCollator(10)(insts)
    assert isinstance(b, AllItemsBatch) and b["all_items"] is True
    assert set(dict.keys(b)) == {"user", "all_items", "target_csr", "seen_csr"}      # nothing wide yet
    assert "target" in b and "item" in b and "seen_items" in b and "mask" not in b  # ... but advertised
    assert b.get("mask") is None
    assert b["user"].tolist() == [3, 9, 1]
    indptr, indices = b["target_csr"]
    assert indptr.tolist() == [0, 2, 2, 3] and indices.tolist() == [2, 5, 8]
    sp, si = b["seen_csr"]
    assert sp.tolist() == [0, 2, 3, 6] and si.tolist() == [1, 7, 4, 2, 3, 6]
    # reference collator (experiments/bpr/dataset.py:274-296) restated
    target = torch.zeros(3, 10)
    for r, inst in enumerate(insts):
        target[r, torch.as_tensor(inst["item"], dtype=torch.long)] = 1.0
    assert torch.equal(b["target"], target)
    assert torch.equal(b["item"], torch.arange(10).unsqueeze(0).repeat(3, 1))
    assert torch.equal(b["seen_items"], torch.tensor([[1, 7, 0], [4, 0, 0], [2, 3, 6]]))
    assert "target" in dict.keys(b)  # cached after the first read


def test_owner_sharded_loaders_run_the_same_number_of_steps_on_every_rank():
    from rbpr.parallel import OwnerBatchSampler, RoundRobinBatches, owned_triples
    rng = np.random.default_rng(0)
    deg = rng.integers(1, 30, size=200)
    deg[0] = 0
    indptr = np.concatenate([[0], np.cumsum(deg)]).astype(np.int64)
    nnz, world, B = int(indptr[-1]), 3, 16
    seen = []
    for rank in range(world):
        s = OwnerBatchSampler(indptr, world, rank, B, generator=torch.Generator().manual_seed(rank))
        batches = list(s)
        assert len(batches) == len(s) == (nnz + B * world - 1) // (B * world)
        assert all(len(b) == B for b in batches)
        lo, hi = owned_triples(indptr, world, rank)
        flat = np.concatenate(batches)
        assert flat.min() >= lo and flat.max() < hi
        seen.append(set(flat.tolist()))
    assert not (seen[0] & seen[1]) and not (seen[1] & seen[2])

    class DS:
        def __len__(self):
            return nnz
    for rank in range(world):
        lo, hi = owned_triples(indptr, world, rank)
        ch = EpochChunks(DS(), B, steps_per_chunk=5, generator=torch.Generator().manual_seed(rank), owned=(lo, hi), world=world)
        chunks = list(ch)
        assert len(chunks) == len(ch)
        total = sum(c["triple_idx"].numel() for c in chunks)
        assert total == ch.steps_per_epoch * B == ((nnz + B * world - 1) // (B * world)) * B
        assert all(lo <= int(c["triple_idx"].min()) and int(c["triple_idx"].max()) < hi for c in chunks)
    parts = [list(RoundRobinBatches(list(range(10)), 3, r)) for r in range(3)]
    assert parts == [[0, 3, 6, 9], [1, 4, 7], [2, 5, 8]]
    assert [len(RoundRobinBatches(list(range(10)), 3, r)) for r in range(3)] == [4, 3, 3]


def _reduce_worker(rank, world, port, q):
    import os
    import torch.distributed as dist
    os.environ.update(MASTER_ADDR="127.0.0.1", MASTER_PORT=str(port), RANK=str(rank), WORLD_SIZE=str(world))
    dist.init_process_group("gloo", rank=rank, world_size=world)
    try:
        from experiments._accel import Accelerator
        from experiments.options import _reduce_metrics
        from rbpr.parallel import sync_row_shards

        class FakeMetric:  # the streaming state of revisit_bpr.metrics._TopkMean
            def __init__(self, total, count):
                self._total, self._total_count = torch.tensor(float(total)), torch.tensor(float(count))
        # rank 0 saw 3 users (sum 1.5), rank 1 saw 1 user (sum 1.0): exact mean 2.5/4, not mean of means 0.75
        m = FakeMetric(1.5, 3) if rank == 0 else FakeMetric(1.0, 1)
        values = {"ndcg@10": m._total / m._total_count, "loss": torch.tensor(2.0 + rank), "epoch_tag": "x"}
        out = _reduce_metrics(values, {"ndcg@10": m}, Accelerator(device="cpu"))
        t = torch.arange(12, dtype=torch.float32).reshape(6, 2) * (rank + 1)
        sync_row_shards([t], np.array([0, 4, 6]))
        q.put((rank, float(out["ndcg@10"]), float(out["loss"]), out["epoch_tag"], t.tolist()))
    finally:
        dist.destroy_process_group()


def test_metric_reduction_is_one_exact_sum_count_allreduce_gloo():
    import socket
    import torch.multiprocessing as mp
    with socket.socket() as s:
        s.bind(("127.0.0.1", 0))
        port = s.getsockname()[1]
    ctx = mp.get_context("spawn")
    q = ctx.Queue()
    procs = [ctx.Process(target=_reduce_worker, args=(r, 2, port, q)) for r in range(2)]
    for p in procs:
        p.start()
    res = sorted(q.get(timeout=120) for _ in procs)
    for p in procs:
        p.join(timeout=60)
        assert p.exitcode == 0
    base = torch.arange(12, dtype=torch.float32).reshape(6, 2)
    want = torch.cat([base[:4], 2 * base[4:]]).tolist()  # rows 0..3 from rank 0, rows 4..5 from rank 1
    for rank, ndcg, loss, tag, rows in res:
        assert abs(ndcg - 2.5 / 4) < 1e-12 and abs(loss - 2.5) < 1e-12 and tag == "x"
        assert rows == want


def test_parameters_follow_the_engine_into_the_symmetric_buffer():
    """Engine.init_fused_exchange (symmetric-memory binding) moves the item table / bias into a buffer
    every peer maps; MF.adopt_engine_tables re-points the parameters there WITHOUT replacing the
    parameter objects (the torch optimizer and its state are keyed by them) and keeps the engine."""
    from types import SimpleNamespace
    from revisit_bpr.models.bpr.model import MF
    U, I, D = 7, 9, 4
    mf = MF(torch.nn.Embedding(U, D, padding_idx=0), torch.nn.Embedding(I, D, padding_idx=0), item_bias=True)
    with torch.no_grad():
        mf._item_bias.copy_(torch.arange(mf._item_bias.numel(), dtype=torch.float32).view_as(mf._item_bias))
    feats = mf.get_features()
    iw, ib = feats["item"], feats["item_bias"]
    opt = torch.optim.Adam(mf.parameters(), lr=1e-3)
    opt.state[iw]["exp_avg"] = torch.ones_like(iw)
    # what the engine holds after the bind: views of ONE buffer, filled with the old values
    buf = torch.zeros(I * D + I + 16)
    new_item = buf[8:8 + I * D].view(I, D)
    new_bias = buf[8 + I * D:8 + I * D + I].view(ib.shape)
    new_item.copy_(iw.data)
    new_bias.copy_(ib.data)
    mf._engine = SimpleNamespace(item_emb=new_item, item_bias=new_bias)
    before = iw.detach().clone()
    mf.adopt_engine_tables()
    assert mf.get_features()["item"] is iw and "exp_avg" in opt.state[iw]        # same parameter object
    assert iw.data_ptr() == new_item.data_ptr() and ib.data_ptr() == new_bias.data_ptr()
    assert torch.equal(iw.detach(), before)
    buf[8] = 42.0                                                                  # a peer's store lands ...
    assert iw[0, 0].item() == 42.0                                                 # ... in the parameter
    key = mf._engine_key
    assert key[1] == iw.data_ptr() and key[2] == ib.data_ptr()                    # engine() will not rebuild
    mf.adopt_engine_tables()                                                       # idempotent
    assert mf._engine_key == key
