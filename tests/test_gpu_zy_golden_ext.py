"""The CUDA path against the committed fixtures of the widened rows (bags, Ftrl, the census input layer):
tests/golden/*.npz written by make_golden.py, reproduced by the oracle in tests/test_golden.py."""
import importlib.util
import os

import numpy as np
import pytest
import torch

from tests._util import REL, rel_err, to_dev

pytestmark = pytest.mark.gpu
GOLD = os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden")


def _load(name):
    return dict(np.load(os.path.join(GOLD, name)))


@pytest.mark.parametrize("name", ["bags_mean_k16.npz", "bags_sqrtn_k8.npz"])
def test_bags_step_matches_fixture(pkg, cuda, name):
    g = _load(name)
    B, F, K = g["e"].shape
    rows = [int(r) for r in g["rows"]]
    layer = pkg.EmbeddingBagFM(F, K, rows, combiner=str(g["combiner"]), optimizer="adagrad", lr=float(g["lr"])).train()
    layer.load_tables(g["table"], g["w1"])
    with torch.no_grad():
        layer.bias.fill_(0.125)
    first, fm, emb = layer.forward_bags(to_dev(g["bag_offsets"]), to_dev(g["bag_index"]), to_dev(g["bag_weight"]))
    assert np.array_equal(emb.detach().cpu().numpy().reshape(B, F, K), g["e32"]), "combined embeddings must be bit-exact"
    assert rel_err(fm.detach().cpu().numpy(), g["fm"], 0.5 * (g["e"] ** 2).sum((1, 2))[:, None] + 1e-30) <= REL
    assert rel_err(first.detach().cpu().numpy(), g["first"], np.abs(g["w1"]).max() * F * 4 + 0.125) <= REL
    torch.autograd.backward((first, fm, emb), (to_dev(g["g_first"])[:, None], to_dev(g["g_fm"])[:, None],
                                               to_dev(g["u"].reshape(B, -1))))
    torch.cuda.synchronize()
    got_t, got_w = layer.table.cpu().numpy(), layer.w1.cpu().numpy()
    assert int(layer.last_n_unique.item()) == len(g["touched"])
    assert rel_err(got_t, g["table_after"], np.abs(g["table"]).max()) <= REL
    assert rel_err(got_w, g["w1_after"], np.abs(g["w1"]).max() + 1e-3) <= REL
    untouched = np.ones(len(g["w1"]), bool)
    untouched[g["touched"]] = False
    assert np.array_equal(got_t[untouched], g["table"][untouched])


@pytest.mark.parametrize("sharded", [False, True])
def test_ftrl_two_steps_match_fixture(pkg, cuda, sharded):
    g = _load("ftrl_two_steps.npz")
    rows = [int(r) for r in g["rows"]]
    F, K = len(rows), g["table"].shape[1]
    B = g["idx"].shape[0]
    N = g["table"].shape[0]
    cls = pkg.ShardedEmbeddingFM if sharded else pkg.EmbeddingFM
    layer = cls(F, K, rows, optimizer="adagrad", lr=float(g["lr"]), linear_optimizer="ftrl", linear_lr=float(g["lin_lr"]),
                l1_regularization_strength=float(g["l1"]), l2_regularization_strength=float(g["l2"])).train()
    layer.load_tables(g["table"], g["w1"])
    idx, val = to_dev(g["idx"]), to_dev(g["val"])
    for s in range(g["g_first"].shape[0]):
        first, fm, emb = layer(idx, val)
        torch.autograd.backward((first, fm, emb), (to_dev(g["g_first"][s])[:, None], to_dev(g["g_fm"][s])[:, None],
                                                   to_dev(g["u"][s].reshape(B, -1))))
    torch.cuda.synchronize()
    assert rel_err(layer.table.cpu().numpy()[:N], g["table_after"], np.abs(g["table"]).max()) <= REL
    assert rel_err(layer.w1.cpu().numpy()[:N], g["w1_after"], np.abs(g["w1"]).max() + float(g["lin_lr"])) <= REL
    assert rel_err(layer.lin_z.cpu().numpy()[:N, 0], g["z_after"], np.abs(g["z_after"]).max() + 1.0) <= REL
    assert rel_err(layer.w1_accum.cpu().numpy()[:N], g["n_after"], 0.1 + np.abs(g["n_after"]).max()) <= REL


def test_input_layer_matches_fixture(pkg, cuda):
    spec = importlib.util.spec_from_file_location("make_golden", os.path.join(GOLD, "make_golden.py"))
    mg = importlib.util.module_from_spec(spec)
    spec.loader.exec_module(mg)
    g = _load("input_layer_census.npz")
    layer = pkg.InputLayer(mg.CENSUS)
    emb = to_dev(g["emb"]).requires_grad_(True)
    x0 = layer(to_dev(g["numeric"]), to_dev(g["indicator_ids"]), emb)
    assert np.array_equal(x0.detach().cpu().numpy(), g["x0"])
    dy = torch.arange(x0.numel(), device="cuda", dtype=torch.float32).reshape(x0.shape)
    x0.backward(dy)
    a, b = (int(v) for v in g["occupation_columns"])
    assert torch.equal(emb.grad, dy[:, a:b])
