"""Multi-hot / weighted bags (EmbeddingBagFM, dir_embed_bag_*) against the oracle: forward of the three
combiners, backward + fused update, and the reduction to the one-id-per-field path."""
import numpy as np
import pytest
import torch

from oracle import deepctr_oracle as O
from tests._util import REL, rel_err, to_dev

pytestmark = pytest.mark.gpu


def _bags(rng, B, rows, max_len, weighted=True, prune=True, skew=None):
    F = len(rows)
    lens = rng.integers(0, max_len + 1, size=B * F)
    off = np.concatenate([[0], np.cumsum(lens)]).astype(np.int64)
    field = np.repeat(np.arange(B * F) % F, lens)
    r = rng.random(off[-1])
    if skew:
        r = r ** skew
    idx = np.minimum((r * np.asarray(rows)[field]).astype(np.int64), np.asarray(rows)[field] - 1)
    w = (rng.random(off[-1]) + 0.25).astype(np.float32) if weighted else None
    if prune and off[-1] >= 16:
        idx[rng.integers(0, off[-1], size=max(1, off[-1] // 16))] = -1
        if weighted:
            w[rng.integers(0, off[-1], size=max(1, off[-1] // 16))] = 0.0
            w[rng.integers(0, off[-1], size=max(1, off[-1] // 32))] = -1.0
    return off, idx, w


def _tables(rng, rows, K):
    N = int(sum(rows))
    return (rng.standard_normal((N, K)) / np.sqrt(K)).astype(np.float32), (rng.standard_normal(N) * 0.1).astype(np.float32)


SHAPES = [(1, [5], 4, 3), (33, [7, 1, 30, 4], 8, 4), (257, [100] * 10 + [1] * 3, 16, 5), (64, [50, 9, 1000], 32, 12),
          (300, [3, 2], 16, 40), (40, [12, 40, 7], 64, 2)]


@pytest.mark.parametrize("B,rows,K,max_len", SHAPES)
@pytest.mark.parametrize("combiner", ["sum", "mean", "sqrtn"])
@pytest.mark.parametrize("weighted", [False, True])
def test_bag_forward_backward_parity(pkg, cuda, B, rows, K, max_len, combiner, weighted):
    rng = np.random.default_rng(23)
    F = len(rows)
    off_f = np.concatenate([[0], np.cumsum(rows)[:-1]]).astype(np.int64)
    table, w1 = _tables(rng, rows, K)
    off, idx, w = _bags(rng, B, rows, max_len, weighted=weighted, skew=2.0)
    layer = pkg.EmbeddingBagFM(F, K, rows, combiner=combiner, optimizer="adagrad", lr=0.05).train()
    layer.load_tables(table, w1)
    with torch.no_grad():
        layer.bias.fill_(0.125)
    first, fm, emb = layer.forward_bags(to_dev(off), to_dev(idx), to_dev(w))
    e32, first32, x32 = O.embedding_bag_lookup(table, w1, 0.125, off_f, off, idx, w, B, F, combiner)
    assert np.array_equal(emb.detach().cpu().numpy().reshape(B, F, K), e32), "combined embeddings must be bit-exact"
    t64, w64 = table.astype(np.float64), w1.astype(np.float64)
    e64, first64, x64 = O.embedding_bag_lookup(t64, w64, 0.125, off_f, off, idx, w, B, F, combiner, np.float64)
    fm64 = O.fm_second_order(e64)
    assert rel_err(fm.detach().cpu().numpy(), fm64, 0.5 * (e64 ** 2).sum((1, 2))[:, None] + 1e-30) <= REL
    assert rel_err(first.detach().cpu().numpy(), first64, np.abs(w1).max() * F * max_len + 0.125) <= REL
    g_first = rng.standard_normal(B).astype(np.float32)
    g_fm = (rng.standard_normal(B) * 0.1).astype(np.float32)
    u = (rng.standard_normal((B, F, K)) * 0.1).astype(np.float32)
    torch.autograd.backward((first, fm, emb), (to_dev(g_first)[:, None], to_dev(g_fm)[:, None], to_dev(u.reshape(B, -1))))
    torch.cuda.synchronize()
    urows, G, g1 = O.embedding_bag_backward(t64, off_f, off, idx, w, e64, x64, g_first, g_fm, u, B, F, np.float64)
    acc, acc1 = np.full_like(t64, 0.1), np.full_like(w64, 0.1)
    O.sparse_adagrad(t64, acc, urows, G, 0.05)
    O.sparse_adagrad(w64, acc1, urows, g1, 0.05)
    got_t, got_w = layer.table.cpu().numpy(), layer.w1.cpu().numpy()
    assert int(layer.last_n_unique.item()) == len(urows)
    untouched = np.ones(len(w1), bool)
    untouched[urows] = False
    assert np.array_equal(got_t[untouched], table[untouched]) and np.array_equal(got_w[untouched], w1[untouched])
    assert rel_err(got_t[urows], t64[urows], np.abs(table).max()) <= REL
    assert rel_err(got_w[urows], w64[urows], np.abs(w1).max() + 1e-3) <= REL


def test_single_entry_bags_equal_the_dense_path(pkg, cuda):
    """One entry per bag under 'sum' is the one-id-per-field layer: same rows bit for bit, same update to 1e-6."""
    rng = np.random.default_rng(5)
    rows, K, B = [50, 1, 9, 1000, 3], 16, 200
    F = len(rows)
    table, w1 = _tables(rng, rows, K)
    idx = np.stack([rng.integers(0, r, size=B) for r in rows], 1).astype(np.int64)
    val = (rng.random((B, F)) + 0.25).astype(np.float32)
    u = (rng.standard_normal((B, F * K)) * 0.1).astype(np.float32)
    g = rng.standard_normal((B, 1)).astype(np.float32)
    outs = []
    for bag in (False, True):
        layer = pkg.EmbeddingBagFM(F, K, rows, combiner="sum", optimizer="adagrad", lr=0.05).train()
        layer.load_tables(table, w1)
        if bag:
            first, fm, emb = layer.forward_bags(to_dev(np.arange(B * F + 1, dtype=np.int64)), to_dev(idx.reshape(-1)),
                                                to_dev(val.reshape(-1)))
        else:
            first, fm, emb = layer(to_dev(idx), to_dev(val))
        torch.autograd.backward((first, fm, emb), (to_dev(g), to_dev(g), to_dev(u)))
        torch.cuda.synchronize()
        outs.append((emb.detach().cpu().numpy(), fm.detach().cpu().numpy(), layer.table.cpu().numpy(), layer.w1.cpu().numpy()))
    assert np.array_equal(outs[0][0], outs[1][0])
    assert np.allclose(outs[0][1], outs[1][1], rtol=1e-6, atol=1e-6)
    assert np.allclose(outs[0][2], outs[1][2], rtol=1e-5, atol=1e-6)
    assert np.allclose(outs[0][3], outs[1][3], rtol=1e-5, atol=1e-6)


def test_bag_edge_cases(pkg, cuda):
    layer = pkg.EmbeddingBagFM(2, 8, [4, 3], combiner="mean").train()
    # every bag empty: zero embeddings, logits = bias, nothing updated
    before = layer.table.clone()
    first, fm, emb = layer.forward_bags(torch.zeros(7, dtype=torch.int64, device="cuda"),
                                        torch.zeros(0, dtype=torch.int64, device="cuda"))
    assert float(emb.detach().abs().max()) == 0 and float(fm.detach().abs().max()) == 0
    (first.sum() + fm.sum() + emb.sum()).backward()
    torch.cuda.synchronize()
    assert torch.equal(layer.table, before)
    with pytest.raises(ValueError):
        layer.forward_bags(torch.zeros(4, dtype=torch.int64, device="cuda"), torch.zeros(1, dtype=torch.int64, device="cuda"))
    with pytest.raises(ValueError):
        layer.forward_bags(torch.zeros(5, dtype=torch.int32, device="cuda"), torch.zeros(1, dtype=torch.int64, device="cuda"))
    chk = pkg.EmbeddingBagFM(2, 8, [4, 3], combiner="sum", check_bounds=True).train()
    with pytest.raises(IndexError):
        chk.forward_bags(torch.tensor([0, 1, 2], dtype=torch.int64, device="cuda"),
                         torch.tensor([1, 99], dtype=torch.int64, device="cuda"))
