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
