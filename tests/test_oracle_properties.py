"""Property tests of the oracle (hypothesis): the size-independent laws the GPU tests lean on.
  FM            permutation of fields leaves fm unchanged; fm(a e) = a^2 fm(e); fm == pairwise sum
  backward      linear in (g_first, g_fm, u) for fixed inputs; duplicates sum; pruned lookups contribute nothing
  Adagrad       splitting a batch into its duplicates first (dedupe) is what the update sees: order-free
  cross         L = 0 is the identity; w = 0 adds the biases; backward is linear in dy
"""
import numpy as np
from hypothesis import given, settings
from hypothesis import strategies as st

from oracle import deepctr_oracle as O

SET = settings(max_examples=25, deadline=None)


def _case(seed, B, F, K, n):
    rng = np.random.default_rng(seed)
    rows = rng.integers(1, n + 1, size=F)
    off = np.concatenate([[0], np.cumsum(rows)[:-1]]).astype(np.int64)
    table = rng.standard_normal((int(rows.sum()), K))
    w1 = rng.standard_normal(int(rows.sum()))
    idx = np.stack([rng.integers(-1, r, size=B) for r in rows], 1).astype(np.int64)      # -1: pruned
    val = rng.random((B, F)) - 0.1                                                       # some <= 0: pruned
    return rng, rows, off, table, w1, idx, val


@SET
@given(st.integers(0, 10_000), st.integers(1, 9), st.integers(1, 7), st.sampled_from([1, 4, 8]), st.integers(1, 12))
def test_fm_laws(seed, B, F, K, n):
    rng, rows, off, table, w1, idx, val = _case(seed, B, F, K, n)
    e, keep = O.embedding_lookup(table, off, idx, val, "sum", np.float64)
    fm = O.fm_second_order(e)
    perm = rng.permutation(F)
    assert np.allclose(O.fm_second_order(e[:, perm, :]), fm, rtol=1e-11, atol=1e-12)
    assert np.allclose(O.fm_second_order(3.0 * e), 9.0 * fm, rtol=1e-11, atol=1e-12)
    assert np.allclose(O.fm_pairwise(e), fm, rtol=1e-10, atol=1e-11)
    assert np.all(e[~keep] == 0)                                   # pruned lookups are the zero vector


@SET
@given(st.integers(0, 10_000), st.integers(1, 9), st.integers(1, 6), st.sampled_from([2, 4]), st.integers(1, 5))
def test_backward_is_linear_and_sparse(seed, B, F, K, n):
    rng, rows, off, table, w1, idx, val = _case(seed, B, F, K, n)
    ga, gb = rng.standard_normal((2, B)), rng.standard_normal((2, B))
    ua, ub = rng.standard_normal((B, F, K)), rng.standard_normal((B, F, K))
    ra, Ga, g1a, da = O.embedding_backward(table, off, idx, val, ga[0], ga[1], ua, "sum", np.float64)
    rb, Gb, g1b, db = O.embedding_backward(table, off, idx, val, gb[0], gb[1], ub, "sum", np.float64)
    rc, Gc, g1c, dc = O.embedding_backward(table, off, idx, val, 2 * ga[0] - gb[0], 2 * ga[1] - gb[1], 2 * ua - ub,
                                           "sum", np.float64)
    assert np.array_equal(ra, rb) and np.array_equal(ra, rc)       # the touched rows depend on the ids only
    assert np.allclose(Gc, 2 * Ga - Gb, rtol=1e-10, atol=1e-11)
    assert np.allclose(g1c, 2 * g1a - g1b, rtol=1e-10, atol=1e-11)
    assert np.isclose(dc, 2 * da - db)
    keep = (idx >= 0) & (val > 0)
    want_rows = np.unique((np.where(keep, idx, 0) + off[None, :])[keep])
    assert np.array_equal(ra, want_rows)


@SET
@given(st.integers(0, 10_000), st.integers(2, 30), st.sampled_from([1, 4]))
def test_adagrad_sees_the_deduplicated_gradient(seed, n, K):
    """KAT-5 as a law: a row hit m times gets ONE update with the summed gradient, whatever the order."""
    rng = np.random.default_rng(seed)
    var = rng.standard_normal((5, K))
    rows_dup = rng.integers(0, 5, size=n)
    g = rng.standard_normal((n, K))
    uniq, inv = np.unique(rows_dup, return_inverse=True)
    G = np.zeros((len(uniq), K))
    np.add.at(G, inv, g)
    a, acc = var.copy(), np.full_like(var, 0.1)
    O.sparse_adagrad(a, acc, uniq, G, 0.1)
    p = rng.permutation(n)
    G2 = np.zeros_like(G)
    np.add.at(G2, inv[p], g[p])
    b, acc2 = var.copy(), np.full_like(var, 0.1)
    O.sparse_adagrad(b, acc2, uniq, G2, 0.1)
    assert np.allclose(a, b, rtol=1e-12, atol=1e-13) and np.allclose(acc, acc2, rtol=1e-12)
    rest = np.setdiff1d(np.arange(5), uniq)
    assert np.array_equal(a[rest], var[rest]) and np.all(acc[rest] == 0.1)


@SET
@given(st.integers(0, 10_000), st.integers(1, 8), st.integers(1, 20), st.integers(0, 5))
def test_cross_laws(seed, B, d, L):
    rng = np.random.default_rng(seed)
    x0 = rng.standard_normal((B, d))
    w, b = rng.standard_normal((L, d)) * 0.3, rng.standard_normal((L, d)) * 0.3
    xL, _ = O.cross_forward(x0, w, b)
    if L == 0:
        assert np.array_equal(xL, x0)
        return
    x_w0, _ = O.cross_forward(x0, np.zeros_like(w), b)
    assert np.allclose(x_w0, x0 + b.sum(0)[None, :], rtol=1e-12, atol=1e-12)
    dy1, dy2 = rng.standard_normal((2, B, d))
    a = O.cross_backward(x0, w, b, dy1)
    c = O.cross_backward(x0, w, b, dy2)
    m = O.cross_backward(x0, w, b, dy1 - 3 * dy2)
    for x, y, z in zip(a, c, m):
        assert np.allclose(z, x - 3 * y, rtol=1e-9, atol=1e-10)
