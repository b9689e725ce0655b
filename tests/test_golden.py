"""The oracle against the committed fixtures (tests/golden/*.npz, written by make_golden.py): CPU."""
import os

import numpy as np
import pytest

from oracle import deepctr_oracle as O

GOLD = os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden")


def load(name):
    return dict(np.load(os.path.join(GOLD, name)))


@pytest.mark.parametrize("name", ["deepfm_cfg1_small.npz", "deepfm_k16_skew.npz"])
def test_oracle_reproduces_deepfm_fixture(name):
    g = load(name)
    N, K = g["table"].shape
    for dtype, tol in ((np.float64, 1e-12), (np.float32, 2e-5)):
        t, a = g["table"].astype(dtype), np.full((N, K), 0.1, dtype)
        l1, a1 = g["w1"].astype(dtype), np.full(N, 0.1, dtype)
        r = O.deepfm_layer_step(t, a, l1, a1, 0.0, g["field_offset"], g["idx"], g["val"].astype(dtype),
                                g["labels"], float(g["lr"]), u=g["u"].astype(dtype), dtype=dtype)
        assert np.array_equal(r["rows"], g["touched"])
        np.testing.assert_allclose(r["e"], g["e"], rtol=tol, atol=tol)
        floor = np.abs(g["e"]).max() ** 2 * g["e"].shape[1]
        np.testing.assert_allclose(r["logits"], g["logits"], rtol=0, atol=tol * max(1.0, floor))
        np.testing.assert_allclose(t, g["table_after"], rtol=0, atol=tol)
        np.testing.assert_allclose(a, g["accum_after"], rtol=tol * 10, atol=tol)
        np.testing.assert_allclose(l1, g["w1_after"], rtol=0, atol=tol)
    # sparse-update semantics frozen in the fixture: rows no lookup touched are bit-identical
    untouched = np.ones(N, bool)
    untouched[g["touched"]] = False
    assert np.array_equal(g["table_after"][untouched], g["table"][untouched].astype(np.float64))
    assert np.all(g["accum_after"][untouched] == 0.1)


@pytest.mark.parametrize("name", ["dcn_d312_l3.npz", "dcn_d51_l2.npz"])
def test_oracle_reproduces_dcn_fixture(name):
    g = load(name)
    f64 = [g[k].astype(np.float64) for k in ("x0", "cross_w", "cross_b", "dy")]
    xL, s = O.cross_forward(*f64[:3])
    dx0, dw, db = O.cross_backward(*f64)
    for got, key in ((xL, "xL"), (s, "s"), (dx0, "dx0"), (dw, "dw"), (db, "db")):
        np.testing.assert_allclose(got, g[key], rtol=1e-12, atol=1e-12)


@pytest.mark.parametrize("name", ["bags_mean_k16.npz", "bags_sqrtn_k8.npz"])
def test_oracle_reproduces_bags_fixture(name):
    g = load(name)
    B, F, K = g["e"].shape
    comb = str(g["combiner"])
    t64, w64 = g["table"].astype(np.float64), g["w1"].astype(np.float64)
    e, first, x = O.embedding_bag_lookup(t64, w64, 0.125, g["field_offset"], g["bag_offsets"], g["bag_index"],
                                         g["bag_weight"], B, F, comb, np.float64)
    np.testing.assert_allclose(e, g["e"], rtol=1e-12, atol=1e-12)
    np.testing.assert_allclose(first, g["first"], rtol=1e-12, atol=1e-12)
    np.testing.assert_allclose(O.fm_second_order(e), g["fm"], rtol=1e-12, atol=1e-12)
    e32, _, _ = O.embedding_bag_lookup(g["table"], g["w1"], 0.125, g["field_offset"], g["bag_offsets"], g["bag_index"],
                                       g["bag_weight"], B, F, comb, np.float32)
    assert np.array_equal(e32, g["e32"])                      # the fp32 evaluation order is frozen bit for bit
    rows, G, g1 = O.embedding_bag_backward(t64, g["field_offset"], g["bag_offsets"], g["bag_index"], g["bag_weight"],
                                           e, x, g["g_first"], g["g_fm"], g["u"], B, F, np.float64)
    assert np.array_equal(rows, g["touched"])
    np.testing.assert_allclose(G, g["G"], rtol=1e-12, atol=1e-12)
    acc, acc1 = np.full_like(t64, 0.1), np.full_like(w64, 0.1)
    O.sparse_adagrad(t64, acc, rows, G, float(g["lr"]))
    O.sparse_adagrad(w64, acc1, rows, g1, float(g["lr"]))
    np.testing.assert_allclose(t64, g["table_after"], rtol=0, atol=1e-12)
    np.testing.assert_allclose(w64, g["w1_after"], rtol=0, atol=1e-12)


def test_oracle_reproduces_ftrl_fixture():
    g = load("ftrl_two_steps.npz")
    t, w = g["table"].astype(np.float64), g["w1"].astype(np.float64)
    acc, n1, z1 = np.full_like(t, 0.1), np.full_like(w, 0.1), np.zeros_like(w)
    for s in range(g["g_first"].shape[0]):
        r, G, g1, _ = O.embedding_backward(t, g["field_offset"], g["idx"], g["val"], g["g_first"][s], g["g_fm"][s],
                                           g["u"][s], "sum", np.float64)
        O.sparse_adagrad(t, acc, r, G, float(g["lr"]))
        O.sparse_ftrl(w, n1, z1, r, g1, float(g["lin_lr"]), float(g["l1"]), float(g["l2"]))
    for got, key in ((t, "table_after"), (w, "w1_after"), (n1, "n_after"), (z1, "z_after")):
        np.testing.assert_allclose(got, g[key], rtol=0, atol=1e-12)
    assert (g["w1_after"] == 0).sum() > 0, "the fixture exercises the l1 proximal step (exact zeros)"


def test_oracle_reproduces_input_layer_fixture():
    import importlib.util
    spec = importlib.util.spec_from_file_location("make_golden", os.path.join(GOLD, "make_golden.py"))
    mg = importlib.util.module_from_spec(spec)
    spec.loader.exec_module(mg)
    g = load("input_layer_census.npz")
    x0, where = O.input_layer(mg.CENSUS, g["numeric"], g["indicator_ids"], g["emb"])
    assert np.array_equal(x0, g["x0"]) and x0.shape[1] == 51
    assert list(where["occupation_embedding"]) == list(g["occupation_columns"])
