"""The oracle against closed-form known-answer tests and against
torch autograd, an independent implementation.  CPU only."""
import numpy as np
import pytest
import torch

from oracle import deepctr_oracle as O
from oracle import tf_semantics as tfs


def _rand_case(seed, B=7, F=5, K=4, rows=(6, 3, 1, 9, 2), weighted=True, dtype=np.float64):
    rng = np.random.default_rng(seed)
    rows = np.asarray(rows)
    off = np.concatenate([[0], np.cumsum(rows)[:-1]])
    N = int(rows.sum())
    table = rng.standard_normal((N, K)).astype(dtype)
    w1 = rng.standard_normal(N).astype(dtype)
    idx = np.stack([rng.integers(0, r, size=B) for r in rows], axis=1).astype(np.int64)
    val = (rng.random((B, F)) + 0.1).astype(dtype) if weighted else None
    return table, w1, off, idx, val, rng


# ---- KAT-1: models/DeepFM/test01.py graph, duplicate-id batch ---------------------------------
def test_kat1_test01_graph():
    # embedding [5,3]; x = lookup; logit = dense(ones)(x) + sum(x^2); sigmoid CE; Adagrad(0.1)
    table = np.array([[0.1, -0.2, 0.3], [0.4, 0.5, -0.6], [0.7, 0.8, 0.9],
                      [-1.0, 1.1, 1.2], [1.3, -1.4, 1.5]], dtype=np.float64)
    ids = np.array([0, 1, 0])
    labels = np.array([1.0, 0.0, 1.0])
    x = table[ids]
    z = x.sum(-1) + (x * x).sum(-1)
    g = tfs.sigmoid(z) - labels
    # closed form: dL/dT[r] = sum_{b: id=r} g_b * (1 + 2 x_r)
    want = {0: (g[0] + g[2]) * (1 + 2 * table[0]), 1: g[1] * (1 + 2 * table[1])}
    # the same through the oracle's segment-sum + Adagrad machinery: feed d(logit)/dx as `u`
    K = 3
    # one padded column so K stays 3 but the layer sees F=1
    per_lookup_u = (g[:, None] * (1 + 2 * x))[:, None, :]
    rows, G, g1, _ = O.embedding_backward(table, [0], ids[:, None], None, np.zeros(3), np.zeros(3),
                                          u=per_lookup_u, dtype=np.float64)
    assert rows.tolist() == [0, 1]
    np.testing.assert_allclose(G[0], want[0], rtol=1e-14)
    np.testing.assert_allclose(G[1], want[1], rtol=1e-14)
    acc = np.full_like(table, 0.1)
    t2 = table.copy()
    O.sparse_adagrad(t2, acc, rows, G, 0.1)
    for r in (0, 1):
        a = 0.1 + want[r] ** 2
        np.testing.assert_allclose(acc[r], a, rtol=1e-14)
        np.testing.assert_allclose(t2[r], table[r] - 0.1 * want[r] / np.sqrt(a), rtol=1e-14)
    assert np.array_equal(t2[2:], table[2:]) and np.all(acc[2:] == 0.1)   # untouched rows


# ---- KAT-2/3: FM identity ------------------------------------------------------------------------
def test_kat2_fm_equals_pairwise_exact_on_integers():
    rng = np.random.default_rng(0)
    e = rng.integers(-3, 4, size=(11, 6, 4)).astype(np.float32)     # exact in fp32
    assert np.array_equal(O.fm_second_order(e), O.fm_pairwise(e))


def test_kat3_fm_degenerate():
    rng = np.random.default_rng(1)
    e1 = rng.standard_normal((5, 1, 8))
    np.testing.assert_allclose(O.fm_second_order(e1), 0.0, atol=1e-15)         # F = 1
    v = rng.standard_normal((5, 1, 8))
    F = 7
    eF = np.repeat(v, F, axis=1)
    np.testing.assert_allclose(O.fm_second_order(eF)[:, 0], 0.5 * F * (F - 1) * (v[:, 0] ** 2).sum(-1), rtol=1e-12)


def test_fm_permutation_invariant():
    rng = np.random.default_rng(2)
    e = rng.standard_normal((9, 6, 4))
    p = rng.permutation(6)
    np.testing.assert_allclose(O.fm_second_order(e), O.fm_second_order(e[:, p]), rtol=1e-12)


# ---- KAT-4: cross ---------------------------------------------------------------------------------
def test_kat4_cross_closed_forms():
    rng = np.random.default_rng(3)
    x0 = rng.standard_normal((6, 10))
    b = rng.standard_normal((3, 10))
    xL, s = O.cross_forward(x0, np.zeros((3, 10)), b)
    np.testing.assert_allclose(xL, x0 + b.sum(0), rtol=1e-13)                   # w = 0
    assert np.all(s == 0)
    x0 = rng.standard_normal((6, 1))
    w = rng.standard_normal((4, 1))
    xL, _ = O.cross_forward(x0, w, np.zeros((4, 1)))
    x = x0.copy()
    for l in range(4):
        x = x * (1 + x0 * w[l, 0])                                            # d = 1, b = 0
    np.testing.assert_allclose(xL, x, rtol=1e-13)


def test_cross_backward_vs_autograd():
    rng = np.random.default_rng(4)
    B, d, L = 9, 13, 4
    x0, w, b, dy = (rng.standard_normal(s) for s in ((B, d), (L, d), (L, d), (B, d)))
    tx0, tw, tb = (torch.tensor(a, requires_grad=True) for a in (x0, w, b))
    xl = tx0
    for l in range(L):
        xl = tx0 * (xl @ tw[l])[:, None] + tb[l] + xl
    xl.backward(torch.tensor(dy))
    dx0, dw, db = O.cross_backward(x0, w, b, dy)
    np.testing.assert_allclose(O.cross_forward(x0, w, b)[0], xl.detach().numpy(), rtol=1e-12)
    np.testing.assert_allclose(dx0, tx0.grad.numpy(), rtol=1e-10, atol=1e-12)
    np.testing.assert_allclose(dw, tw.grad.numpy(), rtol=1e-10, atol=1e-12)
    np.testing.assert_allclose(db, tb.grad.numpy(), rtol=1e-10, atol=1e-12)


# ---- embedding backward vs autograd -------------------------------------------------------------
@pytest.mark.parametrize("weighted", [False, True])
def test_embedding_backward_vs_autograd(weighted):
    table, w1, off, idx, val, rng = _rand_case(5, weighted=weighted)
    B, F = idx.shape
    K = table.shape[1]
    u = rng.standard_normal((B, F, K))
    g_first, g_fm = rng.standard_normal(B), rng.standard_normal(B)
    tt = torch.tensor(table, requires_grad=True)
    tw = torch.tensor(w1, requires_grad=True)
    rows = torch.tensor(idx + off[None, :])
    v = torch.ones((B, F), dtype=torch.float64) if val is None else torch.tensor(val)
    e = tt[rows] * v[:, :, None]
    fm = 0.5 * ((e.sum(1) ** 2) - (e ** 2).sum(1)).sum(-1)
    first = (tw[rows] * v).sum(1)
    loss = (fm * torch.tensor(g_fm)).sum() + (first * torch.tensor(g_first)).sum() + (e * torch.tensor(u)).sum()
    loss.backward()
    ur, G, g1, dbias = O.embedding_backward(table, off, idx, val, g_first, g_fm, u, dtype=np.float64)
    dense = np.zeros_like(table)
    dense[ur] = G
    dense1 = np.zeros_like(w1)
    dense1[ur] = g1
    np.testing.assert_allclose(dense, tt.grad.numpy(), rtol=1e-10, atol=1e-12)
    np.testing.assert_allclose(dense1, tw.grad.numpy(), rtol=1e-10, atol=1e-12)
    np.testing.assert_allclose(dbias, g_first.sum(), rtol=1e-12)
    e_o, _ = O.embedding_lookup(table, off, idx, val, dtype=np.float64)
    np.testing.assert_allclose(O.fm_second_order(e_o)[:, 0], fm.detach().numpy(), rtol=1e-10)
    np.testing.assert_allclose(O.first_order(w1, 0.0, off, idx, val, np.float64)[:, 0], first.detach().numpy(), rtol=1e-10)


# ---- KAT-5: dedupe + Adagrad == torch sparse Adagrad --------------------------------------------
def test_kat5_adagrad_matches_torch_sparse_adagrad():
    rng = np.random.default_rng(6)
    N, K, B = 12, 4, 20
    table = rng.standard_normal((N, K)).astype(np.float32)
    ids = rng.integers(0, 5, size=B)          # heavy duplication; rows 5.. untouched
    g = rng.standard_normal((B, K)).astype(np.float32)
    emb = torch.nn.Embedding(N, K, sparse=True)
    emb.weight.data.copy_(torch.tensor(table))
    opt = torch.optim.Adagrad(emb.parameters(), lr=0.05, initial_accumulator_value=0.1, eps=0)
    (emb(torch.tensor(ids)) * torch.tensor(g)).sum().backward()
    opt.step()
    t2, acc = table.copy(), np.full_like(table, 0.1)
    uniq, inv = np.unique(ids, return_inverse=True)
    G = np.zeros((len(uniq), K), np.float32)
    np.add.at(G, inv, g)
    O.sparse_adagrad(t2, acc, uniq, G, 0.05)
    np.testing.assert_allclose(t2, emb.weight.detach().numpy(), rtol=2e-6, atol=1e-7)
    assert np.array_equal(t2[5:], table[5:])


# ---- KAT-6: weighted semantics --------------------------------------------------------------------
def test_kat6_weighted_semantics_and_pruning():
    table, w1, off, idx, val, _ = _rand_case(7)
    val[0, 0] = 0.0
    val[1, 2] = -1.5
    idx[2, 1] = -1
    e, keep = O.embedding_lookup(table, off, idx, val, "sum", np.float64)
    assert not keep[0, 0] and not keep[1, 2] and not keep[2, 1]
    assert np.all(e[0, 0] == 0) and np.all(e[1, 2] == 0) and np.all(e[2, 1] == 0)
    rows = idx + off[None, :]
    np.testing.assert_allclose(e[3, 1], val[3, 1] * table[rows[3, 1]])
    em, _ = O.embedding_lookup(table, off, idx, val, "mean", np.float64)
    np.testing.assert_allclose(em[3, 1], table[rows[3, 1]])
    assert np.all(em[0, 0] == 0)
    # pruned lookups do not touch their row
    g = np.ones(idx.shape[0])
    ur, G, g1, _ = O.embedding_backward(table, off, idx, val, g, g, None, dtype=np.float64)
    kept_rows = np.unique(rows[keep])
    assert np.array_equal(ur, kept_rows)


def test_step_fp32_close_to_fp64():
    table, w1, off, idx, val, rng = _rand_case(8, B=64, dtype=np.float32)
    labels = (rng.random(64) < 0.25).astype(np.float32)
    outs = {}
    for dt in (np.float32, np.float64):
        t, w = table.astype(dt), w1.astype(dt)
        acc, acc1 = np.full_like(t, 0.1), np.full_like(w, 0.1)
        outs[dt] = (O.deepfm_layer_step(t, acc, w, acc1, 0.0, off, idx, val.astype(dt), labels, 0.05, dtype=dt), t)
    np.testing.assert_allclose(outs[np.float32][0]["logits"], outs[np.float64][0]["logits"], rtol=1e-4, atol=1e-5)
    np.testing.assert_allclose(outs[np.float32][1], outs[np.float64][1], rtol=1e-4, atol=1e-5)


def test_kat7_ftrl_closed_forms():
    """KAT-7: sparse Ftrl (the linear scope's default optimizer, deepFM.py:58).
    (a) first step from (n, z) = (0.1, 0), l1 = l2 = 0:  w' = w (1 - sqrt(0.1)/sqrt(n')) - lr g / sqrt(n')
    (b) with l1 = l2 = 0 Ftrl-proximal keeps z = sum_t (g_t - sigma_t w_t) and w = -z lr / sqrt(n)
    (c) |z| <= l1 snaps the weight to exactly 0; rows not listed are untouched."""
    rng = np.random.default_rng(3)
    w = rng.standard_normal(12)
    n, z = np.full(12, 0.1), np.zeros(12)
    rows = np.array([1, 4, 5, 9])
    g = rng.standard_normal(4)
    w0 = w.copy()
    O.sparse_ftrl(w, n, z, rows, g, 0.2)
    nn = 0.1 + g * g
    assert np.allclose(w[rows], w0[rows] * (1 - np.sqrt(0.1) / np.sqrt(nn)) - 0.2 * g / np.sqrt(nn), rtol=1e-13)
    assert np.allclose(n[rows], nn)
    untouched = np.setdiff1d(np.arange(12), rows)
    assert np.array_equal(w[untouched], w0[untouched]) and np.array_equal(n[untouched], np.full(8, 0.1))
    g2 = rng.standard_normal(4)
    O.sparse_ftrl(w, n, z, rows, g2, 0.2)
    assert np.allclose(w[rows], -z[rows] * 0.2 / np.sqrt(n[rows]), rtol=1e-13)
    # l1 large enough: exact zeros
    w3, n3, z3 = w0.copy(), np.full(12, 0.1), np.zeros(12)
    O.sparse_ftrl(w3, n3, z3, rows, g * 1e-3, 0.2, l1=10.0)
    assert np.array_equal(w3[rows], np.zeros(4))
    # l2 only shrinks: same sign, smaller magnitude than the unregularised step
    w4, n4, z4 = w0.copy(), np.full(12, 0.1), np.zeros(12)
    O.sparse_ftrl(w4, n4, z4, rows, g, 0.2, l2=0.5)
    assert np.all(np.abs(w4[rows]) < np.abs(w[rows] * 0 + (w0[rows] * (1 - np.sqrt(0.1) / np.sqrt(nn)) - 0.2 * g / np.sqrt(nn))))


def _random_bags(rng, B, rows, max_len=4, weighted=True, prune=True):
    F = len(rows)
    lens = rng.integers(0, max_len + 1, size=B * F)
    off = np.concatenate([[0], np.cumsum(lens)]).astype(np.int64)
    field = np.repeat(np.arange(B * F) % F, lens)
    idx = (rng.random(off[-1]) * np.asarray(rows)[field]).astype(np.int64)
    w = (rng.random(off[-1]) + 0.25).astype(np.float32) if weighted else None
    if prune and off[-1] >= 8:
        idx[rng.integers(0, off[-1], size=max(1, off[-1] // 16))] = -1
        if weighted:
            w[rng.integers(0, off[-1], size=max(1, off[-1] // 16))] = 0.0
    return off, idx, w


def test_bags_reduce_to_single_lookups_and_match_autograd():
    """KAT-8: (a) one-entry bags under 'sum' ARE the one-id-per-field path, bit for bit;
    (b) 'mean' of k equal ids is that row; (c) gradients of all three combiners against torch autograd (fp64)."""
    import torch
    rng = np.random.default_rng(9)
    rows = [7, 1, 30, 4]
    B, F, K = 12, 4, 4
    off_f = np.concatenate([[0], np.cumsum(rows)[:-1]]).astype(np.int64)
    N = sum(rows)
    table = rng.standard_normal((N, K)).astype(np.float32)
    w1 = rng.standard_normal(N).astype(np.float32)
    # (a)
    idx = np.stack([rng.integers(0, r, size=B) for r in rows], 1).astype(np.int64)
    val = (rng.random((B, F)) + 0.5).astype(np.float32)
    e1, _ = O.embedding_lookup(table, off_f, idx, val)
    f1 = O.first_order(w1, 0.25, off_f, idx, val)
    eb, fb, xb = O.embedding_bag_lookup(table, w1, 0.25, off_f, np.arange(B * F + 1, dtype=np.int64), idx.reshape(-1),
                                        val.reshape(-1), B, F, "sum")
    assert np.array_equal(eb, e1) and np.array_equal(fb, f1) and np.array_equal(xb, val.reshape(-1))
    # (b)
    off = np.arange(0, 3 * B * F + 1, 3, dtype=np.int64)
    rep = np.repeat(idx.reshape(-1), 3)
    em, _, xm = O.embedding_bag_lookup(table, w1, 0.0, off_f, off, rep, None, B, F, "mean")
    assert np.allclose(em, table[idx + off_f[None, :]], rtol=1e-6) and np.allclose(xm, 1 / 3)
    # (c)
    off, bidx, bw = _random_bags(rng, B, rows)
    g_first, g_fm = rng.standard_normal(B), rng.standard_normal(B)
    u = rng.standard_normal((B, F, K))
    for comb in ("sum", "mean", "sqrtn"):
        t64, w64 = table.astype(np.float64), w1.astype(np.float64)
        e, first, x = O.embedding_bag_lookup(t64, w64, 0.0, off_f, off, bidx, bw, B, F, comb, np.float64)
        rows_u, G, g1 = O.embedding_bag_backward(t64, off_f, off, bidx, bw, e, x, g_first, g_fm, u, B, F, np.float64)
        T = torch.tensor(t64, requires_grad=True)
        W = torch.tensor(w64, requires_grad=True)
        slot = np.repeat(np.arange(B * F), np.diff(off))
        keep = (bidx >= 0) & (bw > 0)
        grow = np.where(keep, bidx, 0) + off_f[slot % F]
        wt = torch.tensor(np.where(keep, bw, 0.0).astype(np.float64))
        acc = torch.zeros((B * F, K), dtype=torch.float64).index_add(0, torch.tensor(slot), wt[:, None] * T[torch.tensor(grow)])
        wsum = torch.zeros(B * F, dtype=torch.float64).index_add(0, torch.tensor(slot), wt)
        wsq = torch.zeros(B * F, dtype=torch.float64).index_add(0, torch.tensor(slot), wt * wt)
        norm = {"sum": torch.ones_like(wsum), "mean": wsum, "sqrtn": wsq.sqrt()}[comb]
        norm = torch.where(wsum > 0, norm, torch.ones_like(norm))
        et = (acc / norm[:, None]).reshape(B, F, K)
        assert np.allclose(et.detach().numpy(), e, rtol=1e-12, atol=1e-12)
        lin = torch.zeros(B * F, dtype=torch.float64).index_add(0, torch.tensor(slot), wt * W[torch.tensor(grow)])
        fm = 0.5 * ((et.sum(1)) ** 2 - (et ** 2).sum(1)).sum(-1)
        loss = (lin.reshape(B, F).sum(1) * torch.tensor(g_first)).sum() + (fm * torch.tensor(g_fm)).sum() + (et * torch.tensor(u)).sum()
        loss.backward()
        assert np.allclose(T.grad.numpy()[rows_u], G, rtol=1e-10, atol=1e-12)
        assert np.allclose(W.grad.numpy()[rows_u], g1, rtol=1e-10, atol=1e-12)
        rest = np.setdiff1d(np.arange(N), rows_u)
        assert np.all(T.grad.numpy()[rest] == 0)


def test_cross_rank1_algebra_of_the_kernels():
    """The CUDA cross kernels do not sweep layer by layer: they use x_l = x0 c_l + beta_l (cross.cu header).
    This restates that algebra in numpy (fp64) and checks it against the layer-by-layer oracle and its backward,
    for random shapes -- the identities the kernels rely on, independent of any GPU."""
    rng = np.random.default_rng(11)
    for B, d, L in [(1, 1, 1), (7, 5, 3), (33, 51, 2), (20, 64, 8), (5, 9, 32)]:
        x0 = rng.standard_normal((B, d)) * 0.5
        w = rng.standard_normal((L, d)) * 0.2
        b = rng.standard_normal((L, d)) * 0.2
        dy = rng.standard_normal((B, d))
        # forward: p_l = x0 . w_l, beta_l = sum_{j<l} b_j, q_l = beta_l . w_l, c_{l+1} = c_l + c_l p_l + q_l
        p = x0 @ w.T                                            # [B, L]
        beta = np.concatenate([np.zeros((1, d)), np.cumsum(b, axis=0)])      # beta[l], l = 0..L
        q = np.einsum("ld,ld->l", beta[:L], w)
        c = np.ones((B, L + 1))
        for l in range(L):
            c[:, l + 1] = c[:, l] + c[:, l] * p[:, l] + q[l]
        xL = x0 * c[:, L:L + 1] + beta[L][None, :]
        want, s = O.cross_forward(x0, w, b)
        assert np.allclose(xL, want, rtol=1e-11, atol=1e-12)
        assert np.allclose(c[:, :L] * p + q[None, :], s, rtol=1e-11, atol=1e-12)      # s_l = c_l p_l + q_l
        # backward: a = dy . x0, ds_l = a + sum_{j>l} ds_j p_j, alpha_l = ds_l c_l
        a = np.sum(dy * x0, axis=1)
        ds = np.zeros((B, L))
        t = np.zeros(B)
        for l in range(L - 1, -1, -1):
            ds[:, l] = a + t
            t = t + ds[:, l] * p[:, l]
        alpha = ds * c[:, :L]
        dx0 = c[:, L:L + 1] * dy + alpha @ w
        D = ds.sum(axis=0)                                      # sum_b ds_l
        dw = alpha.T @ x0 + beta[:L] * D[:, None]
        tail = np.concatenate([np.cumsum((w * D[:, None])[::-1], axis=0)[::-1][1:], np.zeros((1, d))])   # sum_{j>l} w_j D_j
        db = dy.sum(axis=0)[None, :] + tail
        wdx0, wdw, wdb = O.cross_backward(x0, w, b, dy)
        assert np.allclose(dx0, wdx0, rtol=1e-10, atol=1e-11)
        assert np.allclose(dw, wdw, rtol=1e-10, atol=1e-11)
        assert np.allclose(db, wdb, rtol=1e-10, atol=1e-11)


def test_kat9_proximal_adagrad_closed_forms():
    """KAT-9: [TF] SparseApplyProximalAdagrad (models/ESMM/train.py:137-139).  l1 = l2 = 0 is plain Adagrad; a large
    l1 shrinks the weight to exactly 0; l2 alone divides the Adagrad result by 1 + l2 * lr / sqrt(a)."""
    rows = np.array([1, 3])
    g = np.array([[0.5, -2.0], [1.0, 0.25]])
    v0 = np.array([[1.0, 1.0], [0.3, -0.4], [2.0, 2.0], [-0.7, 0.9]])
    v, a = v0.copy(), np.full_like(v0, 0.1)
    O.sparse_proximal_adagrad(v, a, rows, g, 0.05)
    v2, a2 = v0.copy(), np.full_like(v0, 0.1)
    O.sparse_adagrad(v2, a2, rows, g, 0.05)
    assert np.allclose(v, v2, rtol=1e-15) and np.allclose(a, a2) and np.array_equal(v[[0, 2]], v0[[0, 2]])
    v, a = v0.copy(), np.full_like(v0, 0.1)
    O.sparse_proximal_adagrad(v, a, rows, g, 0.05, l1=1e3)
    assert np.array_equal(v[rows], np.zeros((2, 2))) and np.array_equal(v[[0, 2]], v0[[0, 2]])
    v, a = v0.copy(), np.full_like(v0, 0.1)
    O.sparse_proximal_adagrad(v, a, rows, g, 0.05, l2=0.7)
    eta = 0.05 / np.sqrt(0.1 + g * g)
    assert np.allclose(v[rows], (v0[rows] - g * eta) / (1 + 0.7 * eta), rtol=1e-15)
    v, a = v0.copy(), np.full_like(v0, 0.1)
    O.sparse_proximal_adagrad(v, a, rows, g, 0.05, l1=0.5, l2=0.7)
    p = v0[rows] - g * eta
    assert np.allclose(v[rows], np.sign(p) * np.maximum(np.abs(p) - eta * 0.5, 0) / (1 + 0.7 * eta), rtol=1e-15)
