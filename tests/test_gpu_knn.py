"""GPU parity of the ItemKNN / FreeItemKNN kernels (rbpr_knn_forward/backward,
rbpr_freeknn_forward/backward) and of the reference-shaped classes built on them
(revisit_bpr.models.bpr.ItemKNN / FreeItemKNN inside revisit_bpr.models.BPR): against
tests/golden/knn.npz minted from the unmodified reference, and against oracle/knn.py on larger
random inputs.  fp32 kernels vs float64 oracle: 1e-4 relative (sums of up to a few thousand terms)."""
import numpy as np
import pytest
import torch

from helpers import GOLDEN

pytestmark = pytest.mark.gpu
DEV = "cuda:0"
CASES = ["itemknn_bias", "itemknn_fused", "itemknn_noreg", "freeknn_bias", "freeknn_fused"]


def load(name):
    z = np.load(GOLDEN / "knn.npz")
    return {k.split("/", 1)[1]: z[k] for k in z.files if k.startswith(name + "/")}


def dev(a, dtype=None):
    return None if a is None else torch.as_tensor(a, dtype=dtype).to(DEV)


def kernel_forward(ctx, kind, w, b, item, seen):
    if kind == "itemknn":
        logits, profile, keep = ctx.knn_forward(w, b, item, seen)
        return logits, (keep, profile)
    logits, keep = ctx.freeknn_forward(w, b, item, seen)
    return logits, (keep,)


def kernel_backward(ctx, kind, w, with_bias, item, seen, saved, grad):
    gw = torch.zeros_like(w)
    gb = torch.zeros(w.size(0), device=DEV) if with_bias else None
    if kind == "itemknn":
        ctx.knn_backward(w, item, seen, saved[0], saved[1], grad, gw, gb)
    else:
        ctx.freeknn_backward(w.size(0), item, seen, saved[0], grad, gw, gb)
    return gw, gb


@pytest.mark.parametrize("name", CASES)
def test_kernels_reproduce_reference_logits(name):
    from rbpr.engine import Context
    g = load(name)
    kind, fused = name.split("_")[0], name.endswith("fused")
    ctx = Context(DEV)
    w, b, seen = dev(g["w0"]), dev(g.get("b0")), dev(g["seen"])
    if fused:
        pairs = [(np.concatenate([g["item"], g["neg"]], 1), np.concatenate([g["logits_pos"], g["logits_neg"]], 1))]
    else:
        pairs = [(g["item"], g["logits_pos"]), (g["neg"], g["logits_neg"])]
    pairs.append((g["wide"], g["eval_logits"]))
    for ids, want in pairs:
        got, _ = kernel_forward(ctx, kind, w, b, dev(ids), seen)
        ctx.sync_check()
        np.testing.assert_allclose(got.cpu().numpy(), want, rtol=1e-4, atol=1e-5)


@pytest.mark.parametrize("name", CASES)
def test_reference_shaped_model_trains_like_reference(name):
    """revisit_bpr.models.BPR(ItemKNN | FreeItemKNN) + torch.optim.SGD, loop as in example.py:172-180."""
    from revisit_bpr.models import BPR
    from revisit_bpr.models.bpr import FreeItemKNN, ItemKNN
    g = load(name)
    kind, fused = name.split("_")[0], name.endswith("fused")
    I, H = g["w0"].shape
    with_bias = "b0" in g
    lm = ItemKNN(I, H, bias=with_bias) if kind == "itemknn" else FreeItemKNN(I, bias=with_bias)
    with torch.no_grad():
        lm._weights.copy_(torch.as_tensor(g["w0"]))
        if with_bias:
            lm._bias.copy_(torch.as_tensor(g["b0"]))
    model = BPR(lm, reg_alphas={"item": float(g["reg"][0]), "neg": float(g["reg"][1])}, fuse_forward=fused).to(DEV)
    opt = torch.optim.SGD(model.parameters(), lr=float(g["lr"]))
    model.bind_optimizer(opt)  # a no-op for these models
    batch = {"user": torch.zeros(g["item"].shape[0], dtype=torch.long, device=DEV), "item": dev(g["item"]),
             "neg": dev(g["neg"]), "seen_items": dev(g["seen"])}
    model.eval()
    with torch.no_grad():
        ev = model({"user": batch["user"], "item": dev(g["wide"]), "seen_items": batch["seen_items"]})["logits"]
    np.testing.assert_allclose(ev.cpu().numpy(), g["eval_logits"], rtol=1e-4, atol=1e-5)
    model.train()
    for s, want_loss in enumerate(g["losses"]):
        out = model(batch)
        assert set(out) == {"logits_pos", "logits_neg", "logits", "bpr_loss", "l2_reg", "loss"}
        opt.zero_grad()
        out["loss"].backward()
        if s == 0:
            np.testing.assert_allclose(out["logits_pos"].detach().cpu().numpy(), g["logits_pos"], rtol=1e-4, atol=1e-5)
            np.testing.assert_allclose(out["logits_neg"].detach().cpu().numpy(), g["logits_neg"], rtol=1e-4, atol=1e-5)
            np.testing.assert_allclose(out["bpr_loss"].item(), g["bpr_loss"], rtol=1e-4)
            np.testing.assert_allclose(out["l2_reg"].item(), g["l2_reg"], rtol=1e-4, atol=1e-7)
            np.testing.assert_allclose(lm._weights.grad.cpu().numpy(), g["grad_w"], rtol=1e-4, atol=1e-5)
            if with_bias:
                np.testing.assert_allclose(lm._bias.grad.cpu().numpy(), g["grad_b"], rtol=1e-4, atol=1e-6)
        np.testing.assert_allclose(out["loss"].item(), want_loss, rtol=1e-4)
        opt.step()
    np.testing.assert_allclose(lm._weights.detach().cpu().numpy(), g["w_end"], rtol=1e-4, atol=1e-5)
    if with_bias:
        np.testing.assert_allclose(lm._bias.detach().cpu().numpy(), g["b_end"], rtol=1e-4, atol=1e-6)


def _random_case(rng, I, B, S, n):
    seen = np.zeros((B, S), dtype=np.int64)
    for b in range(B):
        k = int(rng.integers(0, S + 1))  # rows without any seen item included
        seen[b, :k] = np.sort(rng.choice(np.arange(1, I), size=k, replace=False))
    item = rng.integers(0, I, size=(B, n))
    item[:, 0] = np.where(seen[:, 0] > 0, seen[:, 0], item[:, 0])  # collisions with the seen list
    return seen, item


@pytest.mark.parametrize("H", [5, 128, 300])
@pytest.mark.parametrize("n", [2, 50])
def test_itemknn_kernels_vs_oracle_random(H, n):
    """n = 2 takes the scan, n = 50 the shared-memory bitmap; H = 5 / 128 / 300 the three profile layouts."""
    from oracle import knn
    from rbpr.engine import Context
    rng = np.random.default_rng(100 + H + n)
    I, B, S = 700, 40, 60
    w = ((rng.random((I, H)) - 0.3) * 0.2).astype(np.float32)
    bias = (rng.standard_normal(I) * 0.1).astype(np.float32)
    seen, item = _random_case(rng, I, B, S, n)
    grad = rng.standard_normal((B, n)).astype(np.float32)
    ctx = Context(DEV)
    dw, db, ditem, dseen = dev(w), dev(bias), dev(item), dev(seen)
    got, saved = kernel_forward(ctx, "itemknn", dw, db, ditem, dseen)
    want = knn.itemknn_forward(w, bias, item, seen)
    np.testing.assert_allclose(got.cpu().numpy(), want, rtol=1e-4, atol=1e-4)
    assert np.array_equal(saved[0].cpu().numpy().astype(bool), knn.keep_mask(item, seen))
    gw, gb = kernel_backward(ctx, "itemknn", dw, True, ditem, dseen, saved, dev(grad))
    ctx.sync_check()
    ow, ob = knn.itemknn_backward(w, item, seen, grad.astype(np.float64), True)
    np.testing.assert_allclose(gw.cpu().numpy(), ow, rtol=1e-4, atol=1e-4)
    np.testing.assert_allclose(gb.cpu().numpy(), ob, rtol=1e-4, atol=1e-5)


@pytest.mark.parametrize("n", [1, 40])
def test_freeknn_kernels_vs_oracle_random(n):
    from oracle import knn
    from rbpr.engine import Context
    rng = np.random.default_rng(7 + n)
    I, B, S = 300, 30, 45
    W = ((rng.random((I, I)) - 0.5) * 0.2).astype(np.float32)
    seen, item = _random_case(rng, I, B, S, n)
    grad = rng.standard_normal((B, n)).astype(np.float32)
    ctx = Context(DEV)
    dW, ditem, dseen = dev(W), dev(item), dev(seen)
    got, saved = kernel_forward(ctx, "freeknn", dW, None, ditem, dseen)
    np.testing.assert_allclose(got.cpu().numpy(), knn.freeknn_forward(W, None, item, seen), rtol=1e-4, atol=1e-5)
    gW, gb = kernel_backward(ctx, "freeknn", dW, False, ditem, dseen, saved, dev(grad))
    ctx.sync_check()
    oW, _ = knn.freeknn_backward(I, item, seen, grad.astype(np.float64), False)
    np.testing.assert_allclose(gW.cpu().numpy(), oW, rtol=1e-4, atol=1e-5)
    assert gb is None


def test_whole_catalog_item_list_masks_every_seen_item():
    """Eval over all items (AllItemsCollator): every seen id occurs in the item list, so the reference's
    mask drops the whole profile and the logits reduce to the bias (model.py:184-196)."""
    from rbpr.engine import Context
    rng = np.random.default_rng(3)
    I, H, B, S = 5000, 64, 16, 30
    w = dev(rng.random((I, H)).astype(np.float32))
    bias = dev(rng.standard_normal(I).astype(np.float32))
    seen, _ = _random_case(rng, I, B, S, 1)
    item = np.tile(np.arange(I), (B, 1))
    ctx = Context(DEV)
    got, saved = kernel_forward(ctx, "itemknn", w, bias, dev(item), dev(seen))
    ctx.sync_check()
    assert int(saved[0].sum().item()) == 0
    assert torch.equal(got, bias.unsqueeze(0).expand(B, I))


def test_out_of_range_ids_are_reported_not_dereferenced():
    from rbpr import native
    from rbpr.engine import Context
    ctx = Context(DEV)
    w = torch.rand(10, 4, device=DEV)
    seen = torch.tensor([[1, 2, 0]], device=DEV)
    ctx.knn_forward(w, None, torch.tensor([[10]], device=DEV), seen)
    with pytest.raises(native.NativeError, match="item id outside"):
        ctx.sync_check()
    ctx.freeknn_forward(torch.rand(10, 10, device=DEV), None, torch.tensor([[3]], device=DEV),
                        torch.tensor([[1, -2, 0]], device=DEV))
    with pytest.raises(native.NativeError, match="item id outside"):
        ctx.sync_check()
    ctx.knn_forward(w, None, torch.tensor([[3]], device=DEV), seen)  # the context stays usable
    ctx.sync_check()


def test_trainer_runs_itemknn_through_autograd_and_optimizer():
    """experiments.trainer.Trainer with a non-fused logits model: accelerator.backward + optimizer.step do
    the update (reference trainer.py:64-83); the loss goes down on a repeated batch."""
    from experiments._accel import Accelerator
    from experiments.trainer import Trainer
    from revisit_bpr.models import BPR
    from revisit_bpr.models.bpr import ItemKNN
    rng = np.random.default_rng(1)
    I, B, S = 200, 64, 12
    seen, _ = _random_case(rng, I, B, S, 1)
    seen[:, 0] = rng.integers(1, I, size=B)
    batch = {"user": torch.zeros(B, dtype=torch.long, device=DEV), "item": dev(seen[:, :1].copy()),
             "neg": dev(rng.integers(1, I, size=(B, 1))), "seen_items": dev(seen)}
    torch.manual_seed(0)
    model = BPR(ItemKNN(I, 16, bias=True), reg_alphas={"all": 1e-4}).to(DEV)
    opt = torch.optim.Adam(model.parameters(), lr=0.05)
    before = model.train()(batch)["loss"].item()  # unfused models: a forward alone does not step
    trainer = Trainer(model, opt, Accelerator(DEV))
    state = trainer.run({"train": [batch] * 20}, epochs=2)
    after = model.train()(batch)["loss"].item()
    assert np.isfinite(after) and after < 0.9 * before
    assert state.metrics["loss"].item() > 0
