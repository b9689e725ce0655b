"""GPU parity of the embedding + first-order + FM path (forward, backward, fused update) against
the oracle, through the reference-facing layer and the C ABI.  Bars (north_star): indices and
gathered rows bit-exact; logits and gradients (here: the updated rows) within 1e-5 relative."""
import numpy as np
import pytest
import torch

from oracle import deepctr_oracle as O
from tests._util import REL, make_case, make_layer, oracle_forward, rel_err, to_dev

pytestmark = pytest.mark.gpu

SHAPES = [  # (B, rows_per_field, K)
    (1, [5], 4),
    (3, [7, 1, 2], 8),
    (64, [50, 1, 9, 1000, 3, 17, 1], 16),
    (257, [100] * 26 + [1] * 13, 16),          # Criteo-shaped: 26 sparse + 13 dense
    (1024, [10_000] * 26 + [1] * 13, 8),       # cfg1 (BASELINE.json configs[0])
    (130, [33] * 70, 32),
    (65, [12, 40, 7], 64),
]


@pytest.mark.parametrize("B,rows,K", SHAPES)
@pytest.mark.parametrize("weighted", [False, True])
def test_forward_parity(pkg, cuda, B, rows, K, weighted):
    case = make_case(11, B, rows, K, weighted=weighted, prune=weighted)
    layer = make_layer(pkg, case).eval()
    with torch.no_grad():
        layer.bias.fill_(0.125)
        first, fm, emb = layer(to_dev(case["idx"]), to_dev(case["val"]))
    e, first_o, fm_o, _ = oracle_forward(case, bias=0.125)
    F = case["F"]
    assert np.array_equal(emb.cpu().numpy().reshape(B, F, K), e), "gathered rows must be bit-exact"
    e64, first64, fm64, _ = oracle_forward(case, bias=0.125, dtype=np.float64)
    floor_fm = 0.5 * (e64 ** 2).sum((1, 2))[:, None] + 1e-30
    floor_first = np.abs(case["w1"]).max() * F + 0.125
    assert rel_err(fm.cpu().numpy(), fm64, floor_fm) <= REL
    assert rel_err(first.cpu().numpy(), first64, floor_first) <= REL
    # and the fp32 oracle itself sits inside the same band (sanity of the bar)
    assert rel_err(fm_o, fm64, floor_fm) <= REL


def _run_step(pkg, case, optimizer, g_first, g_fm, u, lr=0.05):
    layer = make_layer(pkg, case, optimizer=optimizer, lr=lr).train()
    first, fm, emb = layer(to_dev(case["idx"]), to_dev(case["val"]))
    loss = (first[:, 0] * to_dev(g_first)).sum() + (fm[:, 0] * to_dev(g_fm)).sum()
    if u is not None:
        loss = loss + (emb * to_dev(u.reshape(case["B"], -1))).sum()
    loss.backward()
    torch.cuda.synchronize()
    return layer


def _oracle_step(case, optimizer, g_first, g_fm, u, lr=0.05, dtype=np.float32):
    t, w = case["table"].astype(dtype), case["w1"].astype(dtype)
    acc, acc1 = np.full_like(t, 0.1), np.full_like(w, 0.1)
    rows, G, g1, dbias, Gabs, g1abs = O.embedding_backward(t, case["off"], case["idx"], case["val"], g_first,
                                                           g_fm, u, "sum", dtype, return_abs=True)
    case["_Gabs"], case["_g1abs"], case["_G"], case["_g1"] = Gabs, g1abs, G, g1
    if optimizer == "adagrad":
        O.sparse_adagrad(t, acc, rows, G, lr)
        O.sparse_adagrad(w, acc1, rows, g1, lr)
    else:
        O.sparse_sgd(t, rows, G, lr)
        O.sparse_sgd(w, rows, g1, lr)
    return t, acc, w, acc1, rows, G, dbias


@pytest.mark.parametrize("B,rows,K", SHAPES)
@pytest.mark.parametrize("optimizer", ["adagrad", "sgd"])
@pytest.mark.parametrize("with_u", [False, True])
def test_backward_update_parity(pkg, cuda, B, rows, K, optimizer, with_u):
    case = make_case(12, B, rows, K, weighted=True, prune=True)
    rng = case["rng"]
    F = case["F"]
    g_first = rng.standard_normal(B).astype(np.float32)
    g_fm = rng.standard_normal(B).astype(np.float32)
    u = (rng.standard_normal((B, F, K)) * 0.1).astype(np.float32) if with_u else None
    layer = _run_step(pkg, case, optimizer, g_first, g_fm, u)
    t64, acc64, w64, acc1_64, urows, G64, dbias = _oracle_step(case, optimizer, g_first, g_fm, u, dtype=np.float64)
    got_t = layer.table.cpu().numpy()
    got_w = layer.w1.cpu().numpy()
    assert int(layer.last_n_unique.item()) == len(urows)
    # rows the batch did not touch are bit-identical (sparse update semantics)
    untouched = np.ones(case["N"], bool)
    untouched[urows] = False
    assert np.array_equal(got_t[untouched], case["table"][untouched])
    assert np.array_equal(got_w[untouched], case["w1"][untouched])
    assert rel_err(got_t[urows], t64[urows], np.abs(case["table"]).max()) <= REL
    assert rel_err(got_w[urows], w64[urows], np.abs(case["w1"]).max() + 1e-3) <= REL
    if optimizer == "adagrad":
        # acc = 0.1 + G^2: an error dG on the gradient (allowed: REL x the magnitude summed, Gabs)
        # shows up as 2|G|dG on the accumulator
        floor = 0.1 + 2 * np.abs(case["_G"]) * case["_Gabs"]
        assert rel_err(layer.accum.cpu().numpy()[urows], acc64[urows], floor) <= REL
        floor1 = 0.1 + 2 * np.abs(case["_g1"]) * case["_g1abs"]
        assert rel_err(layer.w1_accum.cpu().numpy()[urows], acc1_64[urows], floor1) <= REL
        assert np.all(layer.accum.cpu().numpy()[untouched] == np.float32(0.1))
    if layer.bias.grad is not None:
        assert abs(float(layer.bias.grad.item()) - dbias) <= 1e-4 * (np.abs(g_first).sum() + 1)


def test_sgd_delta_is_gradient(pkg, cuda):
    """With SGD and lr=1 the row delta IS the de-duplicated gradient: check it at 1e-5 relative."""
    B, rows, K = 300, [40, 1, 500, 3], 16
    case = make_case(13, B, rows, K, weighted=True, prune=False, skew=3.0)
    rng = case["rng"]
    g_first = rng.standard_normal(B).astype(np.float32)
    g_fm = rng.standard_normal(B).astype(np.float32)
    u = rng.standard_normal((B, len(rows), K)).astype(np.float32)
    layer = _run_step(pkg, case, "sgd", g_first, g_fm, u, lr=1.0)
    _, _, _, _, urows, G64, _ = _oracle_step(case, "sgd", g_first, g_fm, u, lr=1.0, dtype=np.float64)
    G_got = case["table"][urows].astype(np.float64) - layer.table.cpu().numpy()[urows]
    # subtraction T - G rounds at ulp(T): allow that on top of the gradient tolerance
    floor = case["_Gabs"] + 2 ** -23 / REL * np.abs(case["table"][urows])
    assert rel_err(G_got, G64, floor) <= REL


@pytest.mark.parametrize("B,rows", [(5000, [1]), (40_000, [3, 1]), (9000, [2] * 5)])
def test_long_runs_and_determinism(pkg, cuda, B, rows):
    """Heavy hitters: every lookup of a field hits 1-3 rows -> runs of thousands of duplicates that
    cross many chunks (phase 2 and the one-CTA-per-run phase 3).  Two runs must agree bit for bit."""
    K = 16
    case = make_case(14, B, rows, K, weighted=True)
    rng = case["rng"]
    g_first = rng.standard_normal(B).astype(np.float32)
    g_fm = (rng.standard_normal(B) * 0.01).astype(np.float32)
    u = (rng.standard_normal((B, len(rows), K)) * 0.01).astype(np.float32)
    a = _run_step(pkg, case, "adagrad", g_first, g_fm, u)
    b = _run_step(pkg, case, "adagrad", g_first, g_fm, u)
    assert torch.equal(a.rows, b.rows) and torch.equal(a.lin_rows, b.lin_rows)
    t64, acc64, w64, _, urows, G64, _ = _oracle_step(case, "adagrad", g_first, g_fm, u, dtype=np.float64)
    assert rel_err(a.table.cpu().numpy(), t64, np.abs(case["table"]).max()) <= REL
    floor = np.full_like(acc64, 0.1)
    floor[urows] += 2 * np.abs(case["_G"]) * case["_Gabs"]
    assert rel_err(a.accum.cpu().numpy(), acc64, floor) <= REL
    assert rel_err(a.w1.cpu().numpy(), w64, np.abs(case["w1"]).max()) <= REL


def test_sort_keys_bit_exact_and_oob(pkg, cuda):
    from dir_b200 import _lib
    case = make_case(15, 200, [9, 1, 30], 8, weighted=True, prune=True)
    layer = make_layer(pkg, case, check_bounds=True).train()
    idx = case["idx"].copy()
    B, F = idx.shape
    keys = torch.empty(B * F, dtype=torch.int32, device="cuda")
    fm = torch.empty(B, device="cuda")
    first = torch.empty(B, device="cuda")
    d_idx, d_val = to_dev(idx), to_dev(case["val"])
    _lib.check(_lib.lib().dir_embed_fm_fwd(
        layer.table.data_ptr(), layer.row_stride, layer.w1.data_ptr(), layer.lin_stride, None,
        d_idx.data_ptr(), d_val.data_ptr(), layer.field_offset.data_ptr(), layer.field_rows.data_ptr(),
        layer.n_rows, B, F, 8, None, None, first.data_ptr(), fm.data_ptr(), keys.data_ptr(), None,
        torch.cuda.current_stream().cuda_stream), "fwd")
    keep = (idx >= 0) & (case["val"] > 0)
    want = np.where(keep, idx + case["off"][None, :], case["N"]).astype(np.uint32).reshape(-1)
    assert np.array_equal(keys.cpu().numpy().view(np.uint32), want)
    # the stand-alone key kernel (what the layer uses, so that the sort can overlap the gather) agrees
    keys2 = torch.empty(B * F, dtype=torch.int32, device="cuda")
    _lib.check(_lib.lib().dir_shard_keys(d_idx.data_ptr(), d_val.data_ptr(), layer.field_offset.data_ptr(),
                                         layer.field_rows.data_ptr(), layer.n_rows, B, F, 1, None, F, keys2.data_ptr(), None,
                                         torch.cuda.current_stream().cuda_stream), "keys")
    assert torch.equal(keys, keys2)
    idx[7, 0] = 9            # one past the end of field 0
    with pytest.raises(IndexError):
        layer(to_dev(idx), d_val)


def test_errors_and_empty_batch(pkg, cuda):
    with pytest.raises(ValueError):
        pkg.EmbeddingFM(0, 16, [])
    with pytest.raises(ValueError):
        pkg.EmbeddingFM(2, 12, [3, 3])
    layer = pkg.EmbeddingFM(2, 16, [3, 3])
    with pytest.raises(ValueError):
        layer(torch.zeros((4, 3), dtype=torch.int64, device="cuda"))
    with pytest.raises(ValueError):
        layer(torch.zeros((4, 2), dtype=torch.int64))        # CPU tensor: no CPU path
    first, fm, emb = layer(torch.zeros((0, 2), dtype=torch.int64, device="cuda"))
    assert first.shape == (0, 1) and fm.shape == (0, 1) and emb.shape == (0, 32)


def test_shared_table_global_ids(pkg, cuda):
    """Classic DeepFM layout: one feature_size-row table, feature_index holds global ids."""
    rng = np.random.default_rng(16)
    N, F, K, B = 500, 6, 16, 77
    layer = pkg.EmbeddingFM(F, K, N).eval()
    idx = rng.integers(0, N, size=(B, F)).astype(np.int64)
    with torch.no_grad():
        first, fm, emb = layer(to_dev(idx))
    t = layer.table.cpu().numpy()
    e = t[idx]
    assert np.array_equal(emb.cpu().numpy().reshape(B, F, K), e)
    fm64 = O.fm_second_order(e.astype(np.float64))
    assert rel_err(fm.cpu().numpy(), fm64, 0.5 * (e.astype(np.float64) ** 2).sum((1, 2))[:, None]) <= REL


def test_onerow_path_equals_sorted_path(pkg, cuda, monkeypatch):
    """One-row (numeric) fields reduced as a column sum must agree with pushing them through the sort."""
    B, rows, K = 700, [1, 40, 1, 1, 9, 1], 16
    case = make_case(17, B, rows, K, weighted=True, prune=True)
    rng = case["rng"]
    g_first = rng.standard_normal(B).astype(np.float32)
    g_fm = (rng.standard_normal(B) * 0.1).astype(np.float32)
    u = (rng.standard_normal((B, len(rows), K)) * 0.1).astype(np.float32)
    a = _run_step(pkg, case, "adagrad", g_first, g_fm, u)
    assert a.n_onerow_fields == 4 and a.n_sorted_fields == 2
    monkeypatch.setenv("DIR_B200_SORT_ALL_FIELDS", "1")
    b = _run_step(pkg, case, "adagrad", g_first, g_fm, u)
    assert b.n_onerow_fields == 0 and b.n_sorted_fields == 6
    assert int(a.last_n_unique.item()) == int(b.last_n_unique.item())
    t64, acc64, w64, _, urows, _, _ = _oracle_step(case, "adagrad", g_first, g_fm, u, dtype=np.float64)
    for layer in (a, b):
        assert rel_err(layer.table.cpu().numpy(), t64, np.abs(case["table"]).max()) <= REL
        assert rel_err(layer.w1.cpu().numpy(), w64, np.abs(case["w1"]).max() + 1e-3) <= REL
    # a field whose every lookup is pruned leaves its row untouched, bit for bit
    case["val"][:, 0] = 0.0
    c = _run_step(pkg, case, "adagrad", g_first, g_fm, u)
    assert np.array_equal(c.table.cpu().numpy()[0], case["table"][0])
    assert float(c.accum.cpu().numpy()[0].max()) == np.float32(0.1)


def test_presort_ahead_of_forward(pkg, cuda):
    """presort(batch) ahead of forward(batch, presorted=handle) gives the same update as sorting inline;
    a handle made for other tensors is refused."""
    B, rows, K = 513, [30, 1, 200, 5], 16
    case = make_case(18, B, rows, K, weighted=True, prune=True)
    rng = case["rng"]
    g_first = rng.standard_normal(B).astype(np.float32)
    g_fm = (rng.standard_normal(B) * 0.1).astype(np.float32)
    u = (rng.standard_normal((B, len(rows), K)) * 0.1).astype(np.float32)
    ref = _run_step(pkg, case, "adagrad", g_first, g_fm, u)
    layer = make_layer(pkg, case, optimizer="adagrad", lr=0.05).train()
    idx, val = to_dev(case["idx"]), to_dev(case["val"])
    h = layer.presort(idx, val)
    other = idx.clone()
    with pytest.raises(ValueError):
        layer(other, val, presorted=h)
    first, fm, emb = layer(idx, val, presorted=h)
    loss = (first[:, 0] * to_dev(g_first)).sum() + (fm[:, 0] * to_dev(g_fm)).sum() + (emb * to_dev(u.reshape(B, -1))).sum()
    loss.backward()
    torch.cuda.synchronize()
    assert torch.equal(layer.rows, ref.rows) and torch.equal(layer.lin_rows, ref.lin_rows)


def test_two_forwards_before_the_first_backward(pkg, cuda):
    """Two towers sharing the layer: forward(A), forward(B), backward(B), backward(A) without presorted handles.  Each
    forward must keep a sorted list of its own (round-1 advisor finding: the second forward used to overwrite the
    first one's).  A and B touch disjoint rows, so the result must equal running A and B one after the other."""
    rows, K, B = [64, 1, 200], 8, 40
    rng = np.random.default_rng(3)
    N = sum(rows)
    table = (rng.standard_normal((N, K)) * 0.3).astype(np.float32)
    w1 = (rng.standard_normal(N) * 0.1).astype(np.float32)
    idxA = np.stack([rng.integers(0, 32, B), np.zeros(B, np.int64), rng.integers(0, 100, B)], 1).astype(np.int64)
    idxB = np.stack([rng.integers(32, 64, B), np.zeros(B, np.int64), rng.integers(100, 200, B)], 1).astype(np.int64)
    val = np.ones((B, 3), np.float32)
    val[:, 1] = 0.0                                     # the shared one-row field is pruned: A and B stay disjoint

    def run(interleaved):
        layer = pkg.EmbeddingFM(3, K, rows, optimizer="adagrad", lr=0.05).train()
        layer.load_tables(table, w1)
        a, b = torch.as_tensor(idxA).cuda(), torch.as_tensor(idxB).cuda()
        v = torch.as_tensor(val).cuda()
        if interleaved:
            fa = layer(a, v)
            fb = layer(b, v)
            (fb[0].sum() + fb[1].sum() + fb[2].sum()).backward()
            (fa[0].sum() + fa[1].sum() + fa[2].sum()).backward()
        else:
            for x in (a, b):
                f = layer(x, v)
                (f[0].sum() + f[1].sum() + f[2].sum()).backward()
        torch.cuda.synchronize()
        return layer.rows.clone(), layer.lin_rows.clone()

    r1, l1 = run(True)
    r2, l2 = run(False)
    assert torch.equal(r1, r2) and torch.equal(l1, l2)


def _sorted_lists(ws, n):
    from dir_b200 import _lib
    sk, sp = _lib.c_void_p(), _lib.c_void_p()
    _lib.check(_lib.lib().dir_embed_bwd_sorted(ws.data_ptr(), n, _lib.ctypes.byref(sk), _lib.ctypes.byref(sp)), "sorted")
    off_k, off_p = sk.value - ws.data_ptr(), sp.value - ws.data_ptr()
    return (ws[off_k:off_k + 4 * n].view(torch.int32).clone(), ws[off_p:off_p + 4 * n].view(torch.int32).clone())


@pytest.mark.parametrize("B,rows", [(200, [9, 1, 30]), (5000, [50, 1, 9, 1000, 3, 17, 1]), (20000, [2000] * 28 + [1, 7])])
def test_fused_keys_sort_equals_the_two_calls(pkg, cuda, B, rows):
    """dir_shard_keys_sort = dir_shard_keys + dir_embed_bwd_sort, bit for bit (keys, sorted keys, sorted positions);
    with and without a field selection; the sorted list is the stable sort of the keys."""
    from dir_b200 import _lib
    L = _lib.lib()
    st = torch.cuda.current_stream().cuda_stream
    case = make_case(17, B, rows, 8, weighted=True, prune=True, skew=2.0)
    F, N = len(rows), case["N"]
    d_idx, d_val = to_dev(case["idx"]), to_dev(case["val"])
    fo, fr = to_dev(case["off"].astype(np.int64)), to_dev(np.asarray(rows, np.int64))
    sel_fields = [f for f, r in enumerate(rows) if r > 1]
    for sel in (None, to_dev(np.asarray(sel_fields, np.int32))):
        n_sel = F if sel is None else len(sel_fields)
        n = B * n_sel
        sp = None if sel is None else sel.data_ptr()
        nbytes = int(L.dir_embed_bwd_workspace_bytes(n, 8))
        k1 = torch.empty(n, dtype=torch.int32, device="cuda")
        ws1 = torch.zeros(nbytes, dtype=torch.uint8, device="cuda")
        _lib.check(L.dir_shard_keys(d_idx.data_ptr(), d_val.data_ptr(), fo.data_ptr(), fr.data_ptr(), N, B, F, 1, sp,
                                    n_sel, k1.data_ptr(), None, st), "keys")
        _lib.check(L.dir_embed_bwd_sort(k1.data_ptr(), n, N, ws1.data_ptr(), nbytes, st), "sort")
        k2 = torch.empty(n, dtype=torch.int32, device="cuda")
        ws2 = torch.zeros(nbytes, dtype=torch.uint8, device="cuda")
        _lib.check(L.dir_shard_keys_sort(d_idx.data_ptr(), d_val.data_ptr(), fo.data_ptr(), fr.data_ptr(), N, B, F, 1, sp,
                                         n_sel, k2.data_ptr(), None, ws2.data_ptr(), nbytes, st), "keys_sort")
        torch.cuda.synchronize()
        assert torch.equal(k1, k2)
        (sk1, sp1), (sk2, sp2) = _sorted_lists(ws1, n), _sorted_lists(ws2, n)
        assert torch.equal(sk1, sk2) and torch.equal(sp1, sp2)
        want_k, want_p = torch.sort(k1.to(torch.int64) & 0xffffffff, stable=True)
        assert torch.equal(sk2.to(torch.int64) & 0xffffffff, want_k) and torch.equal(sp2.to(torch.int64), want_p)


def test_sort_of_nine_million_pairs(pkg, cuda):
    """More tiles than one round of the per-digit row scan covers (256 x 16 tiles = 8.4 M pairs), 30-bit keys."""
    from dir_b200 import _lib
    L = _lib.lib()
    n, n_rows = 9_000_011, (1 << 30) - 5
    g = torch.Generator(device="cuda").manual_seed(3)
    keys64 = torch.randint(0, n_rows + 1, (n,), generator=g, device="cuda", dtype=torch.int64)
    keys64[::7] = keys64[3]                                   # a long run of one key across many tiles
    keys = (keys64 & 0xffffffff).to(torch.int32)
    nbytes = int(L.dir_embed_bwd_workspace_bytes(n, 8))
    ws = torch.zeros(nbytes, dtype=torch.uint8, device="cuda")
    _lib.check(L.dir_embed_bwd_sort(keys.data_ptr(), n, n_rows, ws.data_ptr(), nbytes,
                                    torch.cuda.current_stream().cuda_stream), "sort")
    torch.cuda.synchronize()
    sk, sp = _sorted_lists(ws, n)
    want_k, want_p = torch.sort(keys64, stable=True)
    assert torch.equal(sk.to(torch.int64) & 0xffffffff, want_k) and torch.equal(sp.to(torch.int64), want_p)
