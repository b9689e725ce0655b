"""The CUDA path at BASELINE.json's FULL sizes (cfg2: B = 65 536, 10 M rows; cfg5: B = 262 144, Zipf(1.1);
cfg3: 6-layer cross on d = 624), where the numpy oracle is too slow.  Checked instead against an independent
torch implementation in fp64 on the device (`tests/_util.torch_embedding_reference`, itself pinned to the oracle in
tests/test_host_cpu.py) and through size-independent properties: gathered rows bit-exact, untouched rows
bit-identical, the distinct-row count, sortedness / permutation / stability of the sorted list, and run-to-run
determinism.  Tolerances are the small-case ones (north_star: 1e-5 relative, denominators per SURVEY 7.2)."""
import ctypes

import numpy as np
import pytest
import torch

from tests._util import REL, torch_embedding_reference

pytestmark = pytest.mark.gpu


def _build(pkg, name, optimizer, lr):
    w = pkg.synth.cfg(name)
    idx, val, _ = pkg.synth.make_inputs(w)
    torch.manual_seed(pkg.synth.SEED_TABLES)
    layer = pkg.EmbeddingFM(w.field_size, w.embedding_size, list(w.rows_per_field), optimizer=optimizer, lr=lr).train()
    with torch.no_grad():
        layer.w1.normal_(0.0, 0.01)          # TF's zero init would make the first-order path trivial
        layer.bias.fill_(0.125)
    return w, layer, torch.as_tensor(idx).cuda(), torch.as_tensor(val).cuda()


def _max_rel(got, ref, floor):
    return float(((got - ref).abs() / torch.maximum(ref.abs(), floor)).max())


def test_cfg2_forward_full_size(pkg, cuda):
    w, layer, idx, val = _build(pkg, "cfg2", "adagrad", 0.05)
    B, F, K = idx.shape[0], w.field_size, w.embedding_size
    with torch.no_grad():
        first, fm, emb = layer(idx, val)
    torch.cuda.synchronize()
    table = layer.table
    rows = idx + layer.field_offset[None, :]
    keep = (idx >= 0) & (val > 0)
    want = table[rows] * torch.where(keep, val, torch.zeros_like(val))[..., None]      # fp32: the kernel's value * row
    assert torch.equal(emb.view(B, F, K), want), "gathered rows must be bit-exact at full size"
    ref = torch_embedding_reference(table, layer.w1, 0.125, layer.field_offset, idx, val)
    floor_fm = 0.5 * (ref["e"] ** 2).sum((1, 2)) + 1e-30
    assert _max_rel(fm[:, 0].double(), ref["fm"], floor_fm) <= REL
    assert _max_rel(first[:, 0].double(), ref["first"], ref["first_abs"] + 0.125) <= REL


@pytest.mark.parametrize("name,optimizer,lr", [("cfg2", "sgd", 1.0), ("cfg2", "adagrad", 0.05), ("cfg5", "adagrad", 0.05)])
def test_backward_update_full_size(pkg, cuda, name, optimizer, lr):
    w, layer, idx, val = _build(pkg, name, optimizer, lr)
    B, F, K = idx.shape[0], w.field_size, w.embedding_size
    N = w.n_rows
    gen = torch.Generator(device="cuda").manual_seed(pkg.synth.SEED_UPSTREAM)
    g_first = torch.randn(B, device="cuda", generator=gen) * 0.5
    g_fm = torch.randn(B, device="cuda", generator=gen) * 0.5
    u = torch.randn((B, F * K), device="cuda", generator=gen) * 0.01
    rows0, lin0 = layer.rows.clone(), layer.lin_rows.clone()
    acc1_0 = layer.lin_acc.clone() if layer.lin_acc is not None else None

    def step():
        first, fm, emb = layer(idx, val)
        torch.autograd.backward((first, fm, emb), (g_first[:, None], g_fm[:, None], u))
        torch.cuda.synchronize()

    step()
    after, lin_after = layer.rows.clone(), layer.lin_rows.clone()
    n_unique = int(layer.last_n_unique.item())

    T0, w0 = rows0[:, :K], lin0[:, 0]
    ref = torch_embedding_reference(T0, w0, 0.125, layer.field_offset, idx, val, g_first, g_fm, u)
    touched, G, Gfloor, g1, g1abs = ref["touched"], ref["G"], ref["Gfloor"], ref["g1"], ref["g1abs"]
    assert n_unique == int(touched.sum())
    assert torch.equal(after[~touched], rows0[~touched]), "rows the batch did not touch must stay bit-identical"
    assert torch.equal(lin_after[~touched], lin0[~touched])
    got_T, got_w = after[:, :K].double(), lin_after[:, 0].double()
    T64, w64 = T0.double(), w0.double()
    if optimizer == "sgd":          # lr = 1: the row delta IS the de-duplicated gradient
        err = (got_T - (T64 - G)).abs()
        bound = REL * Gfloor + 2.0 ** -23 * (T64.abs() + G.abs())      # + the rounding of the subtraction itself
        assert bool((err <= bound).all()), float((err - bound).max())
        err1 = (got_w - (w64 - g1)).abs()
        assert bool((err1 <= REL * g1abs + 2.0 ** -23 * (w64.abs() + g1.abs())).all())
    else:
        # Adagrad: T' = T - lr f(G), f(G) = G / sqrt(0.1 + G^2), monotone with slope 0.1 / (0.1 + G^2)^1.5 <= 1/sqrt(0.1).
        # The gradient sum is allowed the same cancellation-aware band as in the SGD branch (REL * Gfloor);
        # what reaches the row is that band pushed through f -- exactly, since f is monotone -- plus the rounding of
        # rsqrt / multiply / subtract (a few ulp of the step and of T).  A flat REL * max|T| ignores the conditioning:
        # it is too loose where |G| is large and too tight where a 65 536-term sum lands near zero.
        def f(g):
            return g / (0.1 + g * g).sqrt()

        def band(g, d, t64):
            step = lr * torch.maximum((f(g + d) - f(g)).abs(), (f(g - d) - f(g)).abs())
            return step + 2.0 ** -21 * lr * f(g).abs() + 2.0 ** -23 * (t64.abs() + lr)

        acc64 = 0.1 + G * G
        want_T = torch.where(touched[:, None], T64 - lr * f(G), T64)
        err = (got_T - want_T).abs()
        bound = band(G, REL * Gfloor, T64)
        assert bool((err <= bound).all()), (float((err - bound).max()), int((err > bound).sum()))
        # and the kernel sits far inside that band where the band is widest: one-row fields (B-term column sums, carried
        # in fp64 by the kernel; what is left is the fp32 rounding of the B terms themselves, ~2e-6 on the row at G ~ 0)
        one = [int(layer.field_offset[f]) for f in range(F) if w.rows_per_field[f] == 1]
        if one:
            assert float(err[one].max()) <= 4 * REL * float(T64.abs().max()), float(err[one].max())
        got_acc = after[:, K:].double()
        want_acc = torch.where(touched[:, None], acc64, torch.full_like(acc64, 0.1))
        assert bool(((got_acc - want_acc).abs() <= REL * (0.1 + 2 * G.abs() * Gfloor)).all())
        n64 = 0.1 + g1 * g1
        want_w = torch.where(touched, w64 - lr * f(g1), w64)
        err1 = (got_w - want_w).abs()
        assert bool((err1 <= band(g1, REL * g1abs, w64)).all()), float(err1.max())
        got_n = layer.lin_acc[:, 0].double()
        want_n = torch.where(touched, n64, torch.full_like(n64, 0.1))
        assert bool(((got_n - want_n).abs() <= REL * (0.1 + 2 * g1.abs() * g1abs)).all())
    del ref, G, Gfloor, got_T, T64
    # determinism: the same step from the same state is bit-identical
    with torch.no_grad():
        layer.rows.copy_(rows0)
        layer.lin_rows.copy_(lin0)
        if acc1_0 is not None:
            layer.lin_acc.copy_(acc1_0)
    step()
    assert torch.equal(layer.rows, after) and torch.equal(layer.lin_rows, lin_after)


def test_sorted_list_properties_full_size(pkg, cuda):
    """The (row, position) list the backward consumes: keys ascending, positions a permutation, equal keys in
    position order (stable), keys[i] == key of lookup positions[i]; sorting the same batch again gives the same list."""
    from dir_b200 import _lib
    w, layer, idx, val = _build(pkg, "cfg5", "adagrad", 0.05)
    B = idx.shape[0]
    n = B * layer.n_sorted_fields

    def sorted_views(h):
        sk, sp = ctypes.c_void_p(), ctypes.c_void_p()
        _lib.check(_lib.lib().dir_embed_bwd_sorted(h.ws.buf.data_ptr(), n, ctypes.byref(sk), ctypes.byref(sp)), "sorted")
        base = h.ws.buf.data_ptr()
        ks = h.ws.buf[sk.value - base: sk.value - base + 4 * n].view(torch.int32)
        ps = h.ws.buf[sp.value - base: sp.value - base + 4 * n].view(torch.int32)
        return ks, ps

    h = layer.presort(idx, val)
    torch.cuda.synchronize()
    ks, ps = sorted_views(h)
    assert bool((ks[1:] >= ks[:-1]).all()), "keys must be ascending"               # all keys < 2^31 here
    assert torch.equal(torch.sort(ps.long()).values, torch.arange(n, device="cuda")), "positions must be a permutation"
    assert torch.equal(h.keys[ps.long()], ks), "every entry carries the key of its lookup"
    assert bool(((ks[1:] != ks[:-1]) | (ps[1:] > ps[:-1])).all()), "equal rows must stay in lookup order (stable)"
    ks1, ps1 = ks.clone(), ps.clone()
    h2 = layer.presort(idx, val)
    torch.cuda.synchronize()
    ks2, ps2 = sorted_views(h2)
    assert torch.equal(ks1, ks2) and torch.equal(ps1, ps2)


def test_cfg3_cross_full_size(pkg, cuda):
    B, d, L = 65536, 624, 6
    gen = torch.Generator(device="cuda").manual_seed(7)
    x0 = torch.randn((B, d), device="cuda", generator=gen) * 0.5
    dy = torch.randn((B, d), device="cuda", generator=gen)
    net = pkg.CrossNetwork(d, L).train()
    outs = []
    for _ in range(2):
        net.zero_grad()
        tx0 = x0.clone().requires_grad_(True)
        xL = net(tx0)
        xL.backward(dy)
        torch.cuda.synchronize()
        outs.append((xL.detach().clone(), tx0.grad.clone(), net.cross_w.grad.clone(), net.cross_b.grad.clone()))
    for a, b in zip(*outs):
        assert torch.equal(a, b), "two runs must be bit-identical"
    x64 = x0.double().requires_grad_(True)
    w64 = net.cross_w.detach().double().requires_grad_(True)
    b64 = net.cross_b.detach().double().requires_grad_(True)
    xl = x64
    for l in range(L):                                  # DeepCrossNetwork.py:345-346, layer by layer
        xl = x64 * (xl @ w64[l])[:, None] + b64[l] + xl
    xl.backward(dy.double())
    for got, ref in zip(outs[0], (xl.detach(), x64.grad, w64.grad, b64.grad)):
        assert float((got.double() - ref).abs().max()) <= REL * float(ref.abs().max())


def test_column_feed_full_size(pkg, cuda):
    w = pkg.synth.cfg("cfg2")
    idx, val, y = pkg.synth.make_inputs(w)
    sp_f = [f for f, n in enumerate(w.rows_per_field) if n > 1]
    de_f = [f for f, n in enumerate(w.rows_per_field) if n == 1]
    slot = [torch.zeros(idx.shape, dtype=torch.int64, device="cuda"), torch.zeros(val.shape, device="cuda"),
            torch.zeros(y.shape, device="cuda")]
    feeder = pkg.ColumnFeeder(sp_f, de_f, slot)
    host = [torch.as_tensor(idx[:, sp_f]).to(torch.int32).contiguous().pin_memory(),
            torch.as_tensor(val[:, de_f]).contiguous().pin_memory(), torch.as_tensor(y).pin_memory()]
    feeder.prefetch(0, host)
    got = feeder.wait(0)
    torch.cuda.synchronize()
    assert np.array_equal(got[0].cpu().numpy(), idx) and np.array_equal(got[1].cpu().numpy(), val)
    assert np.array_equal(got[2].cpu().numpy(), y)
