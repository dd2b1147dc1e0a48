"""Host-side wiring of the ItemKNN / FreeItemKNN classes without a GPU: the native context is
replaced by a stand-in that answers with oracle/knn.py (test infrastructure only), so that the
autograd Functions, `Model._forward_unfused`, the L2 term and the Trainer's unfused step are
exercised on CPU against tests/golden/knn.npz.  The kernels themselves are checked in
tests/test_gpu_knn.py."""
import numpy as np
import pytest
import torch

from helpers import GOLDEN
from oracle import knn as oracle_knn

CASES = ["itemknn_bias", "itemknn_fused", "itemknn_noreg", "freeknn_bias", "freeknn_fused"]


class OracleContext:
    """Same methods as rbpr.engine.Context's KNN group, computed by the numpy oracle."""

    @staticmethod
    def _np(t):
        return None if t is None else t.detach().numpy()

    def knn_forward(self, weights, bias, item, seen):
        w, it, se = self._np(weights), self._np(item), self._np(seen)
        keep = oracle_knn.keep_mask(it, se)
        profile = np.stack([w[se[b][keep[b]]].astype(np.float64).sum(0) for b in range(it.shape[0])])
        logits = oracle_knn.itemknn_forward(w, self._np(bias), it, se)
        return (torch.as_tensor(logits, dtype=torch.float32), torch.as_tensor(profile, dtype=torch.float32),
                torch.as_tensor(keep.astype(np.uint8)))

    def knn_backward(self, weights, item, seen, keep, profile, grad, grad_w, grad_b):
        gw, gb = oracle_knn.itemknn_backward(self._np(weights), self._np(item), self._np(seen),
                                             self._np(grad).astype(np.float64), grad_b is not None)
        grad_w += torch.as_tensor(gw, dtype=torch.float32)
        if grad_b is not None:
            grad_b += torch.as_tensor(gb, dtype=torch.float32)

    def freeknn_forward(self, weights, bias, item, seen):
        it, se = self._np(item), self._np(seen)
        logits = oracle_knn.freeknn_forward(self._np(weights), self._np(bias), it, se)
        return torch.as_tensor(logits, dtype=torch.float32), torch.as_tensor(oracle_knn.keep_mask(it, se).astype(np.uint8))

    def freeknn_backward(self, num_items, item, seen, keep, grad, grad_w, grad_b):
        gw, gb = oracle_knn.freeknn_backward(num_items, self._np(item), self._np(seen),
                                             self._np(grad).astype(np.float64), grad_b is not None)
        grad_w += torch.as_tensor(gw, dtype=torch.float32)
        if grad_b is not None:
            grad_b += torch.as_tensor(gb, dtype=torch.float32)


@pytest.fixture()
def oracle_backed(monkeypatch):
    from revisit_bpr.models.bpr import knn
    monkeypatch.setattr(knn, "_context", lambda _t: OracleContext())


def load(name):
    z = np.load(GOLDEN / "knn.npz")
    return {k.split("/", 1)[1]: z[k] for k in z.files if k.startswith(name + "/")}


def build(g, kind, fused):
    from revisit_bpr.models import BPR
    from revisit_bpr.models.bpr import FreeItemKNN, ItemKNN
    I, H = g["w0"].shape
    with_bias = "b0" in g
    lm = ItemKNN(I, H, bias=with_bias) if kind == "itemknn" else FreeItemKNN(I, bias=with_bias)
    with torch.no_grad():
        lm._weights.copy_(torch.as_tensor(g["w0"]))
        if with_bias:
            lm._bias.copy_(torch.as_tensor(g["b0"]))
    return BPR(lm, reg_alphas={"item": float(g["reg"][0]), "neg": float(g["reg"][1])}, fuse_forward=fused), lm


@pytest.mark.parametrize("name", CASES)
def test_unfused_model_path_reproduces_reference(oracle_backed, name):
    g = load(name)
    model, lm = build(g, name.split("_")[0], name.endswith("fused"))
    opt = torch.optim.SGD(model.parameters(), lr=float(g["lr"]))
    model.bind_optimizer(opt)
    t = torch.as_tensor
    batch = {"user": torch.zeros(g["item"].shape[0], dtype=torch.long), "item": t(g["item"]), "neg": t(g["neg"]),
             "seen_items": t(g["seen"])}
    model.eval()
    with torch.no_grad():
        ev = model({"user": batch["user"], "item": t(g["wide"]), "seen_items": batch["seen_items"]})
    assert set(ev) == {"logits"}
    np.testing.assert_allclose(ev["logits"].numpy(), g["eval_logits"], rtol=1e-5, atol=1e-5)
    model.train()
    for s, want in enumerate(g["losses"]):
        out = model(batch)
        assert set(out) == {"logits_pos", "logits_neg", "logits", "bpr_loss", "l2_reg", "loss"}
        opt.zero_grad()
        out["loss"].backward()
        if s == 0:
            np.testing.assert_allclose(out["logits_pos"].detach().numpy(), g["logits_pos"], rtol=1e-5, atol=1e-5)
            np.testing.assert_allclose(out["logits_neg"].detach().numpy(), g["logits_neg"], rtol=1e-5, atol=1e-5)
            np.testing.assert_allclose(lm._weights.grad.numpy(), g["grad_w"], rtol=1e-4, atol=1e-5)
            if "b0" in g:
                np.testing.assert_allclose(lm._bias.grad.numpy(), g["grad_b"], rtol=1e-4, atol=1e-6)
        np.testing.assert_allclose(out["loss"].item(), want, rtol=1e-5)
        opt.step()
    np.testing.assert_allclose(lm._weights.detach().numpy(), g["w_end"], rtol=1e-4, atol=1e-5)


def test_eval_mask_and_regularization_keep_their_graph(oracle_backed):
    g = load("itemknn_bias")
    model, lm = build(g, "itemknn", False)
    t = torch.as_tensor
    wide = t(g["wide"])
    mask = torch.ones_like(wide)
    mask[:, 1] = 0
    model.eval()
    out = model({"user": None, "item": wide, "seen_items": t(g["seen"]), "mask": mask})["logits"]
    assert (out[:, 1] == -1e13).all() and (out[:, 0] > -1e12).all()
    reg = model.regularization({"item": t(g["item"]), "neg": t(g["neg"])})
    assert reg.shape == (g["item"].shape[0],) and reg.requires_grad  # trains through autograd, unlike the fused MF path
    np.testing.assert_allclose(reg.sum().item(), g["l2_reg"], rtol=1e-5)


def test_trainer_drives_the_unfused_step(oracle_backed):
    from experiments._accel import Accelerator
    from experiments.trainer import Trainer
    g = load("freeknn_bias")
    model, lm = build(g, "freeknn", False)
    opt = torch.optim.SGD(model.parameters(), lr=float(g["lr"]))
    t = torch.as_tensor
    batch = {"user": torch.zeros(g["item"].shape[0], dtype=torch.long), "item": t(g["item"]), "neg": t(g["neg"]),
             "seen_items": t(g["seen"])}
    trainer = Trainer(model, opt, Accelerator("cpu"))
    trainer.run({"train": [batch] * len(g["losses"])}, epochs=1)
    np.testing.assert_allclose(lm._weights.detach().numpy(), g["w_end"], rtol=1e-4, atol=1e-5)
    np.testing.assert_allclose(lm._bias.detach().numpy(), g["b_end"], rtol=1e-4, atol=1e-6)
    np.testing.assert_allclose(trainer.engines["train"].state.metrics["loss"].item(), g["losses"].mean(), rtol=1e-5)
