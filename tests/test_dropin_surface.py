"""CPU checks of the reference-shaped Python surface: names, constructor signatures, parameter
names / state_dict keys, init rule, error behaviour without a GPU.  Where /root/reference exists
(build container) signatures and init values are compared with the reference itself."""
import importlib.util
import inspect
import sys
import types
from pathlib import Path

import numpy as np
import pytest
import torch

REF = Path("/root/reference")


def _ours():
    from revisit_bpr import metrics, models, modules
    from revisit_bpr.models import bpr
    return models, bpr, modules, metrics


def _load_ref(rel: str, name: str):
    """Import ONE reference source file under a private module name (read-only, build container)."""
    spec = importlib.util.spec_from_file_location(name, REF / rel)
    mod = importlib.util.module_from_spec(spec)
    spec.loader.exec_module(mod)
    return mod


def test_exported_names():
    models, bpr, modules, metrics = _ours()
    assert models.BPR is bpr.Model
    for n in ("Model", "MF", "Loss", "BaseLogitModel"):
        assert hasattr(bpr, n)
    for n in ("Sampler", "UniformSampler", "AdaptiveSampler"):
        assert hasattr(modules, n)
    for n in ("Metric", "MaskedMetric", "NDCG", "Recall", "Precision", "MAP", "FBeta", "RocAucOne", "RocAucMany",
              "RocAucManySlow"):
        assert hasattr(metrics, n)
    from experiments.trainer import ModelEvents, Trainer
    assert [e.value for e in ModelEvents] == ["forward_started", "forward_completed", "optimizer_started",
                                              "optimizer_completed"]
    assert list(inspect.signature(Trainer.__init__).parameters) == ["self", "model", "optimizer", "accelerator",
                                                                    "custom_engines"]


def test_constructor_signatures():
    _, bpr, modules, metrics = _ours()
    sig = lambda c: list(inspect.signature(c.__init__).parameters)[1:]  # noqa: E731
    assert sig(bpr.Model) == ["logits_model", "reg_alphas", "fuse_forward"]
    assert sig(bpr.MF) == ["user_emb", "item_emb", "item_bias", "user_bias"]
    assert sig(bpr.Loss) == ["size_average"]
    assert sig(modules.UniformSampler) == ["num_items", "neg_gen"]
    assert sig(modules.AdaptiveSampler) == ["model", "num_items", "sampling_prob", "neg_gen", "every"]
    assert sig(metrics.NDCG) == ["topk", "gain_function"]
    assert sig(metrics.Recall) == ["topk"] and sig(metrics.Precision) == ["topk"]


@pytest.mark.skipif(not REF.exists(), reason="reference tree only exists in the build container")
def test_signatures_equal_reference():
    _, bpr, _, _ = _ours()
    stub = types.ModuleType("revisit_bpr_ref_loss")
    ref_loss = _load_ref("revisit_bpr/models/bpr/loss.py", "revisit_bpr_ref_loss")
    src = (REF / "revisit_bpr/models/bpr/model.py").read_text().replace(
        "from revisit_bpr.models.bpr.loss import Loss", "from revisit_bpr_ref_loss import Loss")
    sys.modules["revisit_bpr_ref_loss"] = ref_loss
    ref_model = types.ModuleType("revisit_bpr_ref_model")
    exec(compile(src, "ref_model", "exec"), ref_model.__dict__)  # noqa: S102
    del stub
    for cls in ("Model", "MF", "ItemKNN", "FreeItemKNN"):
        ours, ref = getattr(bpr, cls), getattr(ref_model, cls)
        po, pr = inspect.signature(ours.__init__).parameters, inspect.signature(ref.__init__).parameters
        assert [(n, q.default, q.kind) for n, q in po.items()] == [(n, q.default, q.kind) for n, q in pr.items()], cls
    # same parameter names and the same init values under the same seed
    def build(mod):
        torch.manual_seed(13)
        return mod.Model(mod.MF(torch.nn.Embedding(50, 16, padding_idx=0), torch.nn.Embedding(37, 16, padding_idx=0),
                                item_bias=True), reg_alphas={"all": 0.1}, fuse_forward=True)
    a, b = build(bpr), build(ref_model)
    sa, sb = a.state_dict(), b.state_dict()
    assert list(sa) == list(sb)
    for k in sa:
        assert torch.equal(sa[k], sb[k]), k
    assert set(a.logits_model.get_features()) == set(b.logits_model.get_features())


@pytest.mark.skipif(not REF.exists(), reason="reference tree only exists in the build container")
def test_knn_models_init_like_reference():
    """Same parameter names, state_dict keys and initial values under the same seed as the reference's
    ItemKNN / FreeItemKNN (model.py:156-174, 202-222)."""
    _, bpr, _, _ = _ours()
    sys.modules["revisit_bpr_ref_loss"] = _load_ref("revisit_bpr/models/bpr/loss.py", "revisit_bpr_ref_loss")
    src = (REF / "revisit_bpr/models/bpr/model.py").read_text().replace(
        "from revisit_bpr.models.bpr.loss import Loss", "from revisit_bpr_ref_loss import Loss")
    ref_model = types.ModuleType("revisit_bpr_ref_model2")
    exec(compile(src, "ref_model", "exec"), ref_model.__dict__)  # noqa: S102
    for make in (lambda m: m.ItemKNN(21, 6, bias=True), lambda m: m.ItemKNN(21, 6, padding_idx=3),
                 lambda m: m.FreeItemKNN(17, bias=True), lambda m: m.FreeItemKNN(17)):
        torch.manual_seed(5)
        a = make(bpr)
        torch.manual_seed(5)
        b = make(ref_model)
        sa, sb = a.state_dict(), b.state_dict()
        assert list(sa) == list(sb)
        for k in sa:
            assert torch.equal(sa[k], sb[k]), k
        assert [n for n, _ in a.named_parameters()] == [n for n, _ in b.named_parameters()]
        fa, fb = a.get_features(), b.get_features()
        assert set(fa) == set(fb) and (fa["bias"] is None) == (fb["bias"] is None)


def test_knn_models_surface_and_cpu_failure():
    from rbpr import native
    models, bpr, _, _ = _ours()
    from revisit_bpr.models.bpr import model as model_module
    assert model_module.ItemKNN is bpr.ItemKNN and model_module.FreeItemKNN is bpr.FreeItemKNN
    sig = lambda c: list(inspect.signature(c.__init__).parameters)[1:]  # noqa: E731
    assert sig(bpr.ItemKNN) == ["num_items", "hidden_dim", "padding_idx", "bias"]
    assert sig(bpr.FreeItemKNN) == ["num_items", "padding_idx", "bias"]
    knn = bpr.ItemKNN(11, 4, bias=True)
    assert knn._weights.shape == (11, 4) and knn._weights[0].abs().sum() == 0 and knn._bias.abs().sum() == 0
    assert 0 < knn._weights[1:].min() and knn._weights.max() < 1
    free = bpr.FreeItemKNN(9)
    assert free._weights.shape == (9, 9) and free._bias is None and "_bias" not in free.state_dict()
    batch = {"user": torch.tensor([1, 2]), "item": torch.tensor([[1], [2]]), "neg": torch.tensor([[3], [4]]),
             "seen_items": torch.tensor([[1, 5], [2, 0]])}
    for lm in (knn, free):
        model = models.BPR(lm, reg_alphas={"item": 0.1})
        model.bind_optimizer(torch.optim.Adagrad(model.parameters(), lr=0.1))  # any optimizer: not fused
        with pytest.raises(native.NativeError):  # CPU tensors: no fallback
            model(batch)
    with pytest.raises(ValueError, match="seen_items should be present"):
        free(None, batch["item"], {})
    with pytest.raises(KeyError):
        knn(None, batch["item"], {})


def test_mf_init_rule_and_features():
    _, bpr, _, _ = _ours()
    torch.manual_seed(0)
    mf = bpr.MF(torch.nn.Embedding(100, 32, padding_idx=0), torch.nn.Embedding(60, 32, padding_idx=0), item_bias=True)
    f = mf.get_features()
    assert set(f) == {"user", "item", "user_bias", "item_bias"} and f["user_bias"] is None
    for t in (f["user"], f["item"]):
        assert t[0].abs().sum() == 0
        assert t.abs().max() <= 0.5 / 32 and t[1:].abs().max() > 0.4 / 32
    assert f["item_bias"].abs().sum() == 0
    with pytest.raises(ValueError):
        bpr.MF(torch.nn.Embedding(5, 8), torch.nn.Embedding(5, 4))


def test_loss_module_matches_logsigmoid():
    _, bpr, _, _ = _ours()
    x = torch.linspace(-30, 30, 101)
    np.testing.assert_allclose(bpr.Loss(size_average=False)(x).numpy(), (-torch.nn.functional.logsigmoid(x)).numpy(),
                               rtol=1e-6, atol=1e-7)
    np.testing.assert_allclose(bpr.Loss()(x).item(), (-torch.nn.functional.logsigmoid(x)).mean().item(), rtol=1e-6)


def test_cpu_calls_fail_loudly_and_optimizer_binding_rules():
    from rbpr import native
    models, bpr, modules, metrics = _ours()
    model = models.BPR(bpr.MF(torch.nn.Embedding(10, 8, padding_idx=0), torch.nn.Embedding(7, 8, padding_idx=0)))
    batch = {"user": torch.tensor([1]), "item": torch.tensor([[1]]), "neg": torch.tensor([[2]])}
    model.train()
    with pytest.raises(native.NativeError):  # CPU tensors: no fallback
        model(batch)
    model.eval()
    with pytest.raises(native.NativeError):
        model({"user": torch.tensor([1]), "item": torch.tensor([[1, 2]])})
    with pytest.raises(native.NativeError):
        metrics.Recall(3).compute(torch.zeros(2, 5), torch.zeros(2, 5))
    from revisit_bpr.metrics.metric import prepare_target
    with pytest.raises(native.NativeError):
        prepare_target(torch.zeros(2, 5), torch.zeros(2, 5))
    with pytest.raises(native.NativeError):
        modules.UniformSampler(7, torch.Generator().manual_seed(1)).sample(
            {"item": torch.zeros(1, 1, dtype=torch.long), "seen_items": torch.zeros(1, 3, dtype=torch.long)})
    # optimizers the fused step stands in for
    for good in (torch.optim.SGD(model.parameters(), lr=0.1),
                 torch.optim.SGD(model.parameters(), lr=0.1, momentum=0.9, nesterov=True),
                 torch.optim.Adam(model.parameters(), lr=0.1, betas=(0.8, 0.9)),
                 torch.optim.RMSprop(model.parameters(), lr=0.1, alpha=0.9)):
        model.bind_optimizer(good)
    for bad in (torch.optim.SGD(model.parameters(), lr=0.1, momentum=0.9, dampening=0.5),
                torch.optim.RMSprop(model.parameters(), lr=0.1, momentum=0.5),
                torch.optim.Adagrad(model.parameters(), lr=0.1),
                torch.optim.Adam(model.parameters(), lr=0.1, weight_decay=0.1)):
        with pytest.raises(NotImplementedError):
            model.bind_optimizer(bad)


def test_mini_engine_event_order_and_filters():
    from experiments._engine import Engine, Events
    log = []
    eng = Engine(lambda e, b: log.append(("step", b)) or b)
    eng.add_event_handler(Events.EPOCH_STARTED | Events.COMPLETED, lambda e: log.append(("eval", e.state.epoch)))
    eng.add_event_handler(Events.GET_BATCH_COMPLETED(every=2), lambda: log.append("every2"))
    st = eng.run([10, 11, 12], max_epochs=2)
    assert st.iteration == 6 and st.epoch == 2
    assert log == [("eval", 1), ("step", 10), "every2", ("step", 11), ("step", 12), ("eval", 2), "every2",
                   ("step", 10), ("step", 11), "every2", ("step", 12), ("eval", 2)]
