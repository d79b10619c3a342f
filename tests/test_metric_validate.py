"""Evaluation side of the path: metrics pinned to the reference's docstring known answers,
validator construction (host logic, CPU), and — on a GPU — ``Solver.eval`` against the oracle's forward."""
import json
import os

import numpy as np
import pytest
import torch

import ppsci

GOLD = json.load(open(os.path.join(os.path.dirname(__file__), "golden", "reference_known_answers.json")))["metric_docstrings"]


def _dicts():
    out = {k: torch.tensor(v, dtype=torch.float32) for k, v in GOLD["output"].items()}
    lab = {k: torch.tensor(v, dtype=torch.float32) for k, v in GOLD["label"].items()}
    return out, lab


@pytest.mark.parametrize("name,keep", [("MSE", False), ("MSE", True), ("MAE", False), ("MAE", True), ("RMSE", False),
                                       ("L2Rel", False), ("MeanL2Rel", False), ("MeanL2Rel", True)])
def test_metric_matches_reference_docstring(name, keep):
    out, lab = _dicts()
    res = getattr(ppsci.metric, name)(keep_batch=keep)(out, lab)
    want = GOLD[name + ("_keep_batch" if keep else "")]
    for key, val in want.items():
        np.testing.assert_allclose(np.asarray(res[key].tolist()), np.asarray(val), rtol=2e-7, atol=0)


def test_metric_argument_errors():
    with pytest.raises(ValueError):
        ppsci.metric.L2Rel(keep_batch=True)
    with pytest.raises(ValueError):
        ppsci.metric.RMSE(keep_batch=True)


def test_geometry_validator_builds_labels_and_loader():
    ppsci.utils.misc.set_random_seed(42)
    rect = ppsci.geometry.Rectangle((0, 0), (1, 1))
    v = ppsci.validate.GeometryValidator(
        {"u": lambda out: out["u"]}, {"u": lambda d: np.sin(d["x"]) * d["y"]}, rect,
        {"dataset": "NamedArrayDataset", "total_size": 37, "batch_size": 16,
         "sampler": {"name": "BatchSampler", "drop_last": False, "shuffle": False}},
        ppsci.loss.MSELoss("mean"), evenly=True, metric={"L2Rel": ppsci.metric.L2Rel()}, name="val")
    assert v.input_keys == ("x", "y") and v.output_keys == ("u",)
    inner = v.data_loader.loader
    assert len(inner) == 3  # 37 points in batches of 16, last batch kept
    n = 0
    for inp, lab, wt in inner:
        np.testing.assert_allclose(lab["u"].numpy(), np.sin(inp["x"].numpy()) * inp["y"].numpy(), rtol=1e-6)
        assert float(wt["u"].min()) == 1.0
        n += inp["x"].shape[0]
    assert n == 37


@pytest.mark.gpu
def test_solver_eval_matches_oracle_forward():
    from oracle import ppsci_oracle as O

    ppsci.utils.misc.set_random_seed(7)
    model = ppsci.arch.MLP(("x", "y"), ("u",), 3, 32, "tanh")
    rect = ppsci.geometry.Rectangle((0, 0), (1, 1))
    exact = lambda d: np.cos(d["x"]) * np.cosh(d["y"])  # noqa: E731
    val = ppsci.validate.GeometryValidator(
        {"u": lambda out: out["u"]}, {"u": exact}, rect,
        {"dataset": "NamedArrayDataset", "total_size": 1000, "batch_size": 300,
         "sampler": {"name": "BatchSampler", "drop_last": False, "shuffle": False}},
        ppsci.loss.MSELoss("mean"), evenly=True,
        metric={"L2Rel": ppsci.metric.L2Rel(), "MSE": ppsci.metric.MSE()}, name="val")
    solver = ppsci.solver.Solver(model, {}, None, None, validator={"val": val})
    target, metrics = solver.eval()
    # oracle forward on the same points with the same parameters
    ds = val.data_loader.loader.ds
    om = O.OracleMLP(("x", "y"), ("u",), [32, 32, 32], "tanh")
    params = model.flat.detach().cpu().double()
    x = {k: torch.as_tensor(v).double() for k, v in ds.input.items()}
    u = om(params, x)["u"]
    lab = torch.as_tensor(ds.label["u"]).double()
    l2 = float(torch.linalg.vector_norm(lab - u) / torch.linalg.vector_norm(lab))
    mse = float(((lab - u) ** 2).mean())
    assert abs(metrics["L2Rel"]["u"] - l2) <= 1e-5 * l2
    assert abs(metrics["MSE"]["u"] - mse) <= 1e-5 * mse
    assert target == pytest.approx(metrics["L2Rel"]["u"])


def test_build_metric_and_validator_from_config():
    ppsci.utils.misc.set_random_seed(1)
    m = ppsci.metric.build_metric([{"MSE": {"keep_batch": False}}, {"L2Rel": None}])
    assert list(m) == ["MSE", "L2Rel"] and isinstance(m["L2Rel"], ppsci.metric.L2Rel)
    geom = {"rect": ppsci.geometry.Rectangle((0, 0), (1, 1))}
    v = ppsci.validate.build_validator([{"GeometryValidator": {
        "output_expr": {"u": lambda out: out["u"]}, "label_dict": {"u": 0}, "geom": "rect",
        "dataloader_cfg": {"dataset": "NamedArrayDataset", "total_size": 10, "batch_size": 4},
        "loss": {"name": "MSELoss", "reduction": "mean"}, "metric": [{"MSE": None}], "name": "v0"}}], geom=geom)
    assert list(v) == ["v0"] and isinstance(v["v0"].metric["MSE"], ppsci.metric.MSE)
    assert ppsci.validate.build_validator(None) is None and ppsci.metric.build_metric(None) is None


def test_supervised_validator_defaults_to_label_keys():
    x = np.linspace(0, 1, 9, dtype=np.float32).reshape(-1, 1)
    v = ppsci.validate.SupervisedValidator(
        {"dataset": {"name": "NamedArrayDataset", "input": {"x": x}, "label": {"u": x ** 2}}, "batch_size": 4},
        ppsci.loss.MSELoss("mean"), metric={"MAE": ppsci.metric.MAE()}, name="sup")
    assert v.input_keys == ("x",) and v.output_keys == ("u",)
    assert float(v.output_expr["u"]({"u": 3.0})) == 3.0
    assert len(v.data_loader.loader) == 3


def test_lbfgs_factory_arguments():
    model = ppsci.arch.MLP(("x",), ("u",), 2, 8)
    opt = ppsci.optimizer.LBFGS(0.5, max_iter=3, history_size=7)(model)
    assert opt.is_lbfgs and opt.get_lr() == 0.5
    opt.set_lr(0.25)
    assert opt.get_lr() == 0.25
    with pytest.raises(ValueError):
        ppsci.optimizer.LBFGS(line_search_fn="armijo")
    with pytest.raises(RuntimeError):  # parameters live on the CPU here: no CPU fallback
        opt.step(lambda: torch.zeros(()))
