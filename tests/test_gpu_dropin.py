"""GPU parity tests of the reference-shaped Python surface (revisit_bpr.models.BPR / MF,
revisit_bpr.modules samplers, revisit_bpr.metrics, experiments.trainer.Trainer) — loops written the
way the reference's example.py:157-230 writes them, checked against trajectories minted from the
reference itself (tests/golden) and against the oracle.  Tolerances as in test_gpu_train.py."""
import numpy as np
import pytest
import torch

from helpers import GOLDEN, TRAIN_CASES, load_train_case

pytestmark = pytest.mark.gpu
DEV = "cuda:0"


def _model_from_case(case):
    from revisit_bpr.models import BPR
    from revisit_bpr.models.bpr import MF
    model = BPR(MF(torch.nn.Embedding(case["U"], case["D"], padding_idx=0),
                   torch.nn.Embedding(case["I"], case["D"], padding_idx=0), item_bias=case["bias"]),
                reg_alphas=case["reg"], fuse_forward=True)
    with torch.no_grad():
        f = model.logits_model.get_features()
        f["user"].copy_(torch.as_tensor(case["init_user"]))
        f["item"].copy_(torch.as_tensor(case["init_item"]))
        if case["bias"]:
            f["item_bias"].copy_(torch.as_tensor(case["init_item_bias"]))
    model = model.to(DEV)
    opt = getattr(torch.optim, case["opt"])(model.parameters(), **case["opt_kw"])
    return model, opt


@pytest.mark.parametrize("name", TRAIN_CASES)
def test_example_style_loop_reproduces_reference_trajectory(name):
    case = load_train_case(name)
    model, opt = _model_from_case(case)
    model.bind_optimizer(opt)
    model.train()
    for s in range(case["triples"].shape[0]):
        t = case["triples"][s]
        batch = {"user": torch.as_tensor(case["coo_user"][t], device=DEV),
                 "item": torch.as_tensor(case["indices"][t], dtype=torch.long, device=DEV),
                 "neg": torch.as_tensor(case["negs"][s], dtype=torch.long, device=DEV).unsqueeze(-1)}
        if batch["item"].dim() < 2:  # example.py:173-174
            batch["item"].unsqueeze_(-1)
        out = model(batch)
        out["loss"].backward()
        opt.step()
        opt.zero_grad()
        assert out["logits_pos"].shape == (case["B"], 1) and out["logits"].shape == (case["B"], 1)
        np.testing.assert_allclose(out["bpr_loss"].item(), case["bpr_loss"][s], rtol=1e-4)
        np.testing.assert_allclose(out["l2_reg"].item(), case["l2_reg"][s], rtol=1e-4, atol=1e-7)
        np.testing.assert_allclose(out["loss"].item(), case["bpr_loss"][s] + case["l2_reg"][s], rtol=1e-4)
        np.testing.assert_allclose(out["logits"].abs().mean().item(), case["logits_abs_mean"][s], rtol=1e-4)
    sd = model.state_dict()  # flushes lazy Adam rows
    np.testing.assert_allclose(sd["logits_model._user_emb.weight"].cpu().numpy(), case["final_user"], atol=1e-5, rtol=1e-4)
    np.testing.assert_allclose(sd["logits_model._item_emb.weight"].cpu().numpy(), case["final_item"], atol=1e-5, rtol=1e-4)
    if case["bias"]:
        np.testing.assert_allclose(sd["logits_model._item_bias"].cpu().numpy(), case["final_item_bias"], atol=1e-5, rtol=1e-4)
    if case["opt"] == "Adam":  # moments live in the torch optimizer's state
        st = opt.state_dict()["state"]
        assert all(float(v["step"]) == case["triples"].shape[0] for v in st.values())


def test_trainer_events_loss_and_eval_logits():
    from experiments._accel import Accelerator
    from experiments.trainer import ModelEvents, Trainer
    from oracle import ref_bpr
    case = load_train_case("sgd_reg3")
    model, opt = _model_from_case(case)
    trainer = Trainer(model, opt, Accelerator(DEV))
    steps = case["triples"].shape[0]
    train_batches = []
    for s in range(steps):
        t = case["triples"][s]
        train_batches.append({"user": torch.as_tensor(case["coo_user"][t], device=DEV),
                              "item": torch.as_tensor(case["indices"][t], dtype=torch.long, device=DEV).unsqueeze(-1),
                              "neg": torch.as_tensor(case["negs"][s], dtype=torch.long, device=DEV).unsqueeze(-1)})
    users = torch.arange(1, 9, device=DEV)
    eval_batches = [{"user": users, "item": torch.arange(case["I"], device=DEV).unsqueeze(0).repeat(8, 1)}]
    seen = []
    trainer.add_event("train", ModelEvents.FORWARD_COMPLETED, lambda e: seen.append(("fwd", e.state.forward_iteration)))
    trainer.add_event("train", ModelEvents.OPTIMIZER_COMPLETED, lambda e: seen.append(("opt", e.state.optimizer_iteration)))
    evals = []
    trainer.add_event("eval", ModelEvents.FORWARD_COMPLETED, lambda e: evals.append(e.state.output["logits"].clone()))
    state = trainer.run({"train": train_batches, "eval": eval_batches}, epochs=1)
    assert seen == [(k, s + 1) for s in range(steps) for k in ("fwd", "opt")]
    assert len(evals) == 2  # before the epoch and after training (trainer.py: EPOCH_STARTED | COMPLETED)
    tr = trainer.engines["train"].state
    np.testing.assert_allclose(tr.metrics["loss"].item(), np.mean(case["bpr_loss"] + case["l2_reg"]), rtol=1e-4)
    assert state is trainer.engines["eval"].state
    # eval logits before training == the oracle's all-item logits of the initial tables
    ref = ref_bpr.RefModel(torch.as_tensor(case["init_user"]), torch.as_tensor(case["init_item"]))
    np.testing.assert_allclose(evals[0].cpu().numpy(), ref.eval_logits(users.cpu(), None).numpy(), atol=1e-5, rtol=1e-5)
    # ... and after training == the golden final tables
    ref = ref_bpr.RefModel(torch.as_tensor(case["final_user"]), torch.as_tensor(case["final_item"]))
    np.testing.assert_allclose(evals[1].cpu().numpy(), ref.eval_logits(users.cpu(), None).numpy(), atol=2e-5, rtol=1e-4)


def test_eval_forward_with_mask_and_biases():
    from revisit_bpr.models import BPR
    from revisit_bpr.models.bpr import MF
    torch.manual_seed(3)
    mf = MF(torch.nn.Embedding(30, 24, padding_idx=0), torch.nn.Embedding(41, 24, padding_idx=0),
            item_bias=True, user_bias=True)
    with torch.no_grad():
        mf._user_emb.weight.mul_(30)
        mf._item_emb.weight.mul_(30)
        mf._item_bias.normal_()
        mf._user_bias.normal_()
    model = BPR(mf).to(DEV).eval()
    users = torch.tensor([3, 7, 7, 29], device=DEV)
    items = torch.randint(0, 41, (4, 13), device=DEV)
    mask = (torch.rand(4, 13, device=DEV) > 0.3).float()
    out = model({"user": users, "item": items, "mask": mask})["logits"]
    f = {k: (None if v is None else v.detach().cpu()) for k, v in mf.get_features().items()}
    exp = torch.einsum("bh,bkh->bk", f["user"][users.cpu()], f["item"][items.cpu()])
    exp = exp + f["item_bias"][items.cpu()] + f["user_bias"][users.cpu()].unsqueeze(-1)
    exp = exp.masked_fill(mask.cpu().eq(0), -1e13)
    np.testing.assert_allclose(out.cpu().numpy(), exp.numpy(), rtol=1e-5, atol=1e-5)
    with pytest.raises(IndexError):
        mf(users[:2], items)


def test_uniform_sampler_contract():
    from oracle import philox
    from revisit_bpr.modules import Sampler, UniformSampler
    z = np.load(GOLDEN / "sampler.npz")
    I = int(z["I"])
    rows = [torch.as_tensor(z["indices"][z["indptr"][u]:z["indptr"][u + 1]], dtype=torch.long) for u in z["users"]]
    seen = torch.nn.utils.rnn.pad_sequence(rows, batch_first=True, padding_value=0).to(DEV)
    gen = torch.Generator(device=DEV).manual_seed(99)
    s = UniformSampler(I, gen)
    assert isinstance(s, Sampler)
    batch = {"item": torch.zeros(seen.size(0), 1, dtype=torch.long, device=DEV), "seen_items": seen}
    a = s.sample(batch)
    b = s.sample(batch)
    assert a.shape == (seen.size(0), 1) and a.dtype == torch.int64 and not torch.equal(a, b)
    for call, got in enumerate((a, b)):  # bit-exact against the CPU restatement of the spec
        exp = philox.sample_negatives_padded(seen.cpu().numpy(), I, 1, 99, call)
        assert got.cpu().numpy().tolist() == exp.tolist()
    for r in range(seen.size(0)):  # never the padding item, never a seen item
        assert a[r, 0].item() > 0 and a[r, 0].item() not in set(seen[r].tolist())
    # distribution over many calls for one row: uniform over the unseen items
    row = seen[:1].repeat(512, 1)
    counts = np.zeros(I)
    for _ in range(40):
        np.add.at(counts, s.sample({"item": torch.zeros(512, 1, dtype=torch.long, device=DEV), "seen_items": row}).cpu().numpy().ravel(), 1)
    allowed = np.setdiff1d(np.arange(1, I), seen[0].cpu().numpy())
    assert counts[np.setdiff1d(np.arange(I), allowed)].sum() == 0
    expc = counts.sum() / allowed.size
    chi2 = ((counts[allowed] - expc) ** 2 / expc).sum()
    assert chi2 < (allowed.size - 1) + 5 * np.sqrt(2 * (allowed.size - 1)), chi2
    # a row that has seen every item: torch.multinomial raises RuntimeError in the reference
    full = torch.arange(I, device=DEV).unsqueeze(0)
    with pytest.raises(RuntimeError):
        s.sample({"item": torch.zeros(1, 1, dtype=torch.long, device=DEV), "seen_items": full})


def test_metrics_match_reference_golden_and_error_types():
    from revisit_bpr.metrics import NDCG, Metric, Precision, Recall
    z = np.load(GOLDEN / "metrics.npz")
    out, tgt = torch.as_tensor(z["output"]).to(DEV), torch.as_tensor(z["target"]).to(DEV)
    for k in (1, 5, 20, 100):
        np.testing.assert_allclose(NDCG(k).compute(out, tgt).cpu().numpy(), z[f"ndcg@{k}"], atol=1e-5)
        np.testing.assert_allclose(Recall(k).compute(out, tgt).cpu().numpy(), z[f"recall@{k}"], atol=1e-6)
    m = NDCG(20)
    assert isinstance(m, Metric)
    m(out[:10], tgt[:10])
    m(out[10:], tgt[10:])
    np.testing.assert_allclose(m.get_metric().item(), z["ndcg@20_stream"], rtol=1e-5)
    assert set(m.state_dict()) == {"total_ndcg", "total_count"}
    m.reset()
    assert m.get_metric().isnan()  # 0/0 like the reference after reset
    o, t = torch.as_tensor(z["kat_output"]).to(DEV), torch.as_tensor(z["kat_target"]).to(DEV)
    np.testing.assert_allclose(NDCG(3).compute(o, t).cpu().numpy(), [0.38685283, 0.0, 1.0], atol=1e-6)
    np.testing.assert_allclose(NDCG(3, "linear").compute(o, t).cpu().numpy(), [1 / 3, 0.0, 1.0], atol=1e-6)
    np.testing.assert_allclose(Recall(3).compute(o, t).cpu().numpy(), [0.5, 0.0, 1.0], atol=1e-6)
    np.testing.assert_allclose(Precision(3).compute(o, t).cpu().numpy(), [1 / 3, 0.0, 2 / 3], atol=1e-6)
    with pytest.raises(IndexError):
        Recall(3).compute(o, t[:, :4])
    with pytest.raises(ValueError):
        Recall(3).compute(o, t * 2)


def test_padding_rows_receive_no_embedding_gradient():
    """nn.Embedding(padding_idx=0): a triple that names row 0 leaves row 0 of the table untouched."""
    from oracle import ref_bpr
    from revisit_bpr.models import BPR
    from revisit_bpr.models.bpr import MF
    torch.manual_seed(5)
    mf = MF(torch.nn.Embedding(12, 16, padding_idx=0), torch.nn.Embedding(9, 16, padding_idx=0), item_bias=True)
    with torch.no_grad():
        mf._user_emb.weight.mul_(40)
        mf._item_emb.weight.mul_(40)
        mf._item_bias.normal_()
    init = {k: v.detach().clone() for k, v in mf.get_features().items() if v is not None}
    model = BPR(mf, reg_alphas={"all": 0.01}).to(DEV).train()
    opt = torch.optim.SGD(model.parameters(), lr=0.1)
    model.bind_optimizer(opt)
    user = torch.tensor([0, 3, 3, 5, 0])
    item = torch.tensor([2, 0, 4, 4, 1])
    neg = torch.tensor([5, 6, 0, 7, 0])
    out = model({"user": user.to(DEV), "item": item.to(DEV).unsqueeze(-1), "neg": neg.to(DEV).unsqueeze(-1)})
    ref = ref_bpr.RefModel(init["user"], init["item"], init["item_bias"], {"all": 0.01})
    ropt = ref_bpr.make_optimizer(ref, "sgd", lr=0.1)
    rout = ref_bpr.train_step(ref, ropt, user, item, neg)
    np.testing.assert_allclose(out["bpr_loss"].item(), rout["bpr_loss"].item(), rtol=1e-5)
    f = mf.get_features()
    np.testing.assert_allclose(f["user"].detach().cpu().numpy(), ref.user_emb.detach().numpy(), atol=1e-6)
    np.testing.assert_allclose(f["item"].detach().cpu().numpy(), ref.item_emb.detach().numpy(), atol=1e-6)
    np.testing.assert_allclose(f["item_bias"].detach().cpu().numpy(), ref.item_bias.detach().numpy(), atol=1e-6)
    assert f["user"][0].abs().sum().item() == 0 and f["item"][0].abs().sum().item() == 0


def test_auc_one_fbeta_and_one_pos_collator():
    """RQ1 protocol: OnePosCollator batch -> eval logits -> RocAucOne; FBeta from the top-k pass."""
    from experiments.bpr.dataset import OnePosCollator
    from revisit_bpr.metrics import FBeta, Precision, Recall, RocAucMany, RocAucManySlow, RocAucOne
    from revisit_bpr.models import BPR
    from revisit_bpr.models.bpr import MF
    torch.manual_seed(9)
    I = 50
    mf = MF(torch.nn.Embedding(20, 12, padding_idx=0), torch.nn.Embedding(I, 12, padding_idx=0))
    with torch.no_grad():
        mf._user_emb.weight.mul_(40)
        mf._item_emb.weight.mul_(40)
    model = BPR(mf).to(DEV).eval()
    batch = OnePosCollator(I)([{"user": 7, "item": 2, "seen_items": [4, 9, 17, 30]}])
    assert batch["item"].shape == (1, 1 + (I - 1 - 4)) and batch["item"][0, 0].item() == 17
    out = model({k: v.to(DEV) for k, v in batch.items()})["logits"]
    got = RocAucOne().compute(out, batch["target"].to(DEV)).cpu()
    lo = out.cpu()[0]
    np.testing.assert_allclose(got.item(), (lo[0] > lo[1:]).float().mean().item(), rtol=1e-6)
    mask = torch.ones_like(lo).unsqueeze(0)
    mask[0, 5:20] = 0
    got = RocAucOne().compute(out, None, mask.to(DEV)).cpu()
    keep = mask[0, 1:] != 0
    np.testing.assert_allclose(got.item(), (lo[0] > lo[1:][keep]).float().mean().item(), rtol=1e-6)
    z = np.load(GOLDEN / "metrics.npz")
    o, t = torch.as_tensor(z["output"]).to(DEV), torch.as_tensor(z["target"]).to(DEV)
    a, b = RocAucMany().compute(o, t), RocAucManySlow().compute(o, t)
    assert torch.equal(a.nan_to_num(-1), b.nan_to_num(-1))
    for beta in (1.0, 0.5):
        p, r = Precision(10).compute(o, t), Recall(10).compute(o, t)
        exp = (1 + beta ** 2) * p * r / (beta ** 2 * p + r + 1e-13)
        np.testing.assert_allclose(FBeta(10, beta).compute(o, t).cpu().numpy(), exp.cpu().numpy(), rtol=1e-6)
    from revisit_bpr.metrics import MAP
    for k, normalized in ((3, True), (10, True), (10, False), (100, True)):  # map.py:45-64 restated
        srt = torch.gather(t.cpu(), 1, torch.argsort(-o.cpu(), dim=-1))[:, :k]
        prec = srt.cumsum(-1) / (torch.arange(srt.size(1)) + 1.0)
        denom = t.cpu().sum(-1).clamp(max=srt.size(1)) if normalized else srt.sum(-1)
        exp = torch.nan_to_num((prec * srt).sum(-1) / denom)
        np.testing.assert_allclose(MAP(k, normalized).compute(o, t).cpu().numpy(), exp.numpy(), atol=1e-6)
    ko, kt = torch.as_tensor(z["kat_output"]).to(DEV), torch.as_tensor(z["kat_target"]).to(DEV)
    np.testing.assert_allclose(MAP(3).compute(ko, kt).cpu().numpy(), [0.25, 0.0, 1.0], atol=1e-6)  # SURVEY §4 KAT
    m = FBeta(10)
    m(o, t)
    assert set(m.state_dict()) == {"total_f", "total_count", "precision", "recall"}


@pytest.mark.parametrize("name", ["adam_all", "adam_bias_ui", "sgd_reg3"])
def test_checkpoint_resume_reproduces_reference_trajectory(name, tmp_path):
    """Stop after half of the reference-minted trajectory, torch.save model + optimizer state_dicts
    (what accelerate.save_state stores, reference options.py:391-400), rebuild everything from the
    files, finish the trajectory: same losses and final tables as the uninterrupted reference run.
    Exercises the lazily-updated Adam user rows (flushed by state_dict) and the step counter."""
    case = load_train_case(name)
    steps = case["triples"].shape[0]
    half = steps // 2

    def batch_of(s):
        t = case["triples"][s]
        return {"user": torch.as_tensor(case["coo_user"][t], device=DEV),
                "item": torch.as_tensor(case["indices"][t], dtype=torch.long, device=DEV).unsqueeze(-1),
                "neg": torch.as_tensor(case["negs"][s], dtype=torch.long, device=DEV).unsqueeze(-1)}

    model, opt = _model_from_case(case)
    model.bind_optimizer(opt)
    model.train()
    for s in range(half):
        out = model(batch_of(s))
        np.testing.assert_allclose(out["bpr_loss"].item(), case["bpr_loss"][s], rtol=1e-4)
    torch.save({"model": model.state_dict(), "optimizer": opt.state_dict()}, tmp_path / "ckpt.pt")
    del model, opt
    ckpt = torch.load(tmp_path / "ckpt.pt", map_location=DEV)
    model2, opt2 = _model_from_case(case)  # fresh objects (initial tables), then restore
    model2.load_state_dict(ckpt["model"])
    opt2.load_state_dict(ckpt["optimizer"])
    model2.bind_optimizer(opt2)
    model2.train()
    for s in range(half, steps):
        out = model2(batch_of(s))
        np.testing.assert_allclose(out["bpr_loss"].item(), case["bpr_loss"][s], rtol=1e-4)
    sd = model2.state_dict()
    np.testing.assert_allclose(sd["logits_model._user_emb.weight"].cpu().numpy(), case["final_user"], atol=1e-5, rtol=1e-4)
    np.testing.assert_allclose(sd["logits_model._item_emb.weight"].cpu().numpy(), case["final_item"], atol=1e-5, rtol=1e-4)
    if case["opt"] == "Adam":
        assert all(float(v["step"]) == steps for v in opt2.state_dict()["state"].values())


def test_prepare_target_helper_matches_argsort_gather():
    """revisit_bpr.metrics.metric.prepare_target (reference metric.py:110-113): target re-ordered by
    descending output; ours ranks with the top-k kernel (at most 128 columns)."""
    from revisit_bpr.metrics.metric import prepare_target
    g = torch.Generator().manual_seed(4)
    out = torch.randn(7, 90, generator=g)
    tgt = (torch.rand(7, 90, generator=g) < 0.2).float()
    want = torch.gather(tgt, -1, torch.argsort(-out, dim=-1))
    got = prepare_target(out.to(DEV), tgt.to(DEV))
    assert got.shape == (7, 90) and torch.equal(got.cpu(), want)
    wide_out = torch.randn(5, 700, generator=g)
    wide_tgt = torch.rand(5, 700, generator=g)  # not binary: the helper only re-orders
    got = prepare_target(wide_out.to(DEV), wide_tgt.to(DEV), topk=25)
    assert torch.equal(got.cpu(), torch.gather(wide_tgt, -1, torch.argsort(-wide_out, dim=-1))[:, :25])
    with pytest.raises(NotImplementedError):
        prepare_target(wide_out.to(DEV), wide_tgt.to(DEV))
    with pytest.raises(IndexError):
        prepare_target(out.to(DEV), tgt[:, :5].to(DEV))
    from revisit_bpr.metrics.metric import _context
    _context(torch.device(DEV)).sync_check()  # the non-binary target raised no device flag


@pytest.mark.parametrize("opt_name,bias", [("sgd", True), ("adam", False)])
def test_several_positives_and_negatives_per_row_match_the_oracle(opt_name, bias):
    """`item` / `neg` of shape (batch, K > 1) — the general form of the reference's train forward
    (model.py:41-42,48-57: logits (batch, num items), loss summed over all of them, the user's L2
    term once per row) — through the fused step: K triples per row, user coefficient / K."""
    from oracle import ref_bpr
    from revisit_bpr.models import BPR
    from revisit_bpr.models.bpr import MF
    dev = torch.device("cuda:0")
    torch.manual_seed(3)
    U, I, D, B, K, steps = 40, 30, 12, 16, 3, 5
    reg = {"user": 0.02, "item": 0.01, "neg": 0.03}
    model = BPR(MF(torch.nn.Embedding(U, D, padding_idx=0), torch.nn.Embedding(I, D, padding_idx=0), item_bias=bias),
                reg_alphas=reg, fuse_forward=True)
    with torch.no_grad():
        model.logits_model._user_emb.weight.mul_(D * 2.0)
        model.logits_model._item_emb.weight.mul_(D * 2.0)
    feats = model.logits_model.get_features()
    ref = ref_bpr.RefModel(feats["user"].detach().clone(), feats["item"].detach().clone(),
                           None if not bias else feats["item_bias"].detach().clone(), reg)
    kw = {"lr": 0.05} if opt_name == "sgd" else {"lr": 0.01, "betas": (0.9, 0.999)}
    ropt = ref_bpr.make_optimizer(ref, opt_name, **kw)
    model = model.to(dev)
    opt = (torch.optim.SGD if opt_name == "sgd" else torch.optim.Adam)(model.parameters(), **kw)
    model.bind_optimizer(opt)
    model.train()
    g = torch.Generator().manual_seed(9)
    for _ in range(steps):
        user = torch.randint(1, U, (B,), generator=g)
        user[1] = user[0]  # a repeated user inside the batch
        item = torch.randint(1, I, (B, K), generator=g)
        neg = torch.randint(1, I, (B, K), generator=g)
        out = model({"user": user.to(dev), "item": item.to(dev), "neg": neg.to(dev)})
        out["loss"].backward()
        opt.step()
        opt.zero_grad()
        exp = ref_bpr.train_step(ref, ropt, user, item, neg)
        assert out["logits_pos"].shape == (B, K) and out["logits"].shape == (B, K)
        np.testing.assert_allclose(out["logits_pos"].cpu().numpy(), exp["logits_pos"].numpy(), atol=1e-5, rtol=1e-4)
        np.testing.assert_allclose(out["logits"].cpu().numpy(), exp["logits"].numpy(), atol=1e-5, rtol=1e-4)
        np.testing.assert_allclose(out["bpr_loss"].item(), exp["bpr_loss"].item(), rtol=1e-4)
        np.testing.assert_allclose(out["l2_reg"].item(), exp["l2_reg"].item(), rtol=1e-4)
    sd = model.state_dict()
    np.testing.assert_allclose(sd["logits_model._user_emb.weight"].cpu().numpy(), ref.user_emb.detach().numpy(), atol=1e-5, rtol=1e-4)
    np.testing.assert_allclose(sd["logits_model._item_emb.weight"].cpu().numpy(), ref.item_emb.detach().numpy(), atol=1e-5, rtol=1e-4)
