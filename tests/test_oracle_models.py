"""The full-model oracle (oracle/full_models.py) against torch autograd in fp64: an independent
implementation of the same graphs (the reference's TF cannot run here)."""
import numpy as np
import torch

from oracle import full_models as FM
from oracle import tf_semantics as tfs
from tests._util import make_case


def _params(case, hidden, rng, dcn_layers=0):
    F, K = case["F"], case["K"]
    d = F * K
    p = dict(table=case["table"].astype(np.float64), w1=case["w1"].astype(np.float64), bias=np.float64(0.03))
    if dcn_layers:
        W, b = FM.make_tower(rng, d, hidden, 0, np.float64)
        p.update(cross_w=tfs.truncated_normal(rng, (dcn_layers, d), 0.1, np.float64),
                 cross_b=tfs.truncated_normal(rng, (dcn_layers, d), 0.1, np.float64), W=W, b=b,
                 Wl=FM.glorot_uniform(rng, d + hidden[-1], 1, np.float64), bl=np.asarray([0.02]))
        del p["w1"], p["bias"]
    else:
        p["W"], p["b"] = FM.make_tower(rng, d, hidden, 1, np.float64)
    return p


def _t(a):
    return torch.tensor(np.asarray(a), dtype=torch.float64, requires_grad=True)


def _torch_embed(table, case):
    idx, val = case["idx"], case["val"]
    keep = torch.as_tensor((idx >= 0) & (val > 0))
    rows = torch.as_tensor(np.where((idx >= 0) & (val > 0), idx, 0) + case["off"][None, :])
    eff = torch.as_tensor(val.astype(np.float64)) * keep
    return table[rows] * eff[:, :, None], rows, eff


def _adagrad(v, g, lr, acc0=0.1):
    a = acc0 + g * g
    return v - lr * g / np.sqrt(a)


def _sparse_grad_rows(g):
    g = g.detach().numpy()
    return np.flatnonzero(np.abs(g).reshape(g.shape[0], -1).sum(1) > 0)


def test_deepfm_step_matches_autograd():
    case = make_case(7, 48, [9, 1, 30, 5, 1], 8, weighted=True, prune=True)
    rng = case["rng"]
    labels = (rng.random(48) < 0.3).astype(np.float64)
    p = _params(case, (16, 8), rng)
    # torch
    T, w1, bias = _t(p["table"]), _t(p["w1"]), _t(p["bias"])
    Ws, bs = [_t(w) for w in p["W"]], [_t(b) for b in p["b"]]
    e, rows, eff = _torch_embed(T, case)
    first = (w1[rows] * eff).sum(1) + bias
    fm = 0.5 * ((e.sum(1)) ** 2 - (e ** 2).sum(1)).sum(-1)
    x = e.reshape(48, -1)
    for i, (W, b) in enumerate(zip(Ws, bs)):
        x = x @ W + b
        if i < len(Ws) - 1:
            x = torch.relu(x)
    logits = first + fm + x[:, 0]
    loss = torch.nn.functional.binary_cross_entropy_with_logits(logits, torch.as_tensor(labels), reduction="sum")
    loss.backward()
    # oracle
    st = FM.adagrad_state(p)
    p0 = {k: ([a.copy() for a in v] if isinstance(v, list) else np.copy(v)) for k, v in p.items()}
    out = FM.deepfm_train_step(p, st, case["off"], case["idx"], case["val"], labels, 0.05)
    assert np.allclose(out["logits"], logits.detach().numpy(), rtol=1e-12, atol=1e-12)
    assert np.isclose(out["loss"], loss.item(), rtol=1e-12)
    touched = out["rows"]
    want_T = p0["table"].copy()
    want_T[touched] = _adagrad(p0["table"][touched], T.grad.numpy()[touched], 0.05)
    assert np.allclose(p["table"], want_T, rtol=1e-10, atol=1e-12)
    want_w = p0["w1"].copy()
    want_w[touched] = _adagrad(p0["w1"][touched], w1.grad.numpy()[touched], 0.05)
    assert np.allclose(p["w1"], want_w, rtol=1e-10, atol=1e-12)
    assert np.isclose(p["bias"], _adagrad(p0["bias"], bias.grad.item(), 0.05), rtol=1e-10)
    for i in range(len(Ws)):
        assert np.allclose(p["W"][i], _adagrad(p0["W"][i], Ws[i].grad.numpy(), 0.05), rtol=1e-9, atol=1e-12)
        assert np.allclose(p["b"][i], _adagrad(p0["b"][i], bs[i].grad.numpy(), 0.05), rtol=1e-9, atol=1e-12)
    # rows nobody looked up keep their accumulator (sparse Adagrad) -- autograd's dense zero gradient
    # would have left them too, but only because g = 0
    untouched = np.setdiff1d(np.arange(case["N"]), touched)
    assert np.array_equal(st["table"][untouched], np.full_like(st["table"][untouched], 0.1))


def test_dcn_step_matches_autograd():
    case = make_case(8, 40, [9, 1, 30, 5], 8, weighted=True, prune=True)
    rng = case["rng"]
    labels = (rng.random(40) < 0.3).astype(np.float64)
    L, hidden, clip = 3, (16, 8), 0.05                         # a clip norm small enough to bite
    p = _params(case, hidden, rng, dcn_layers=L)
    T, cw, cb = _t(p["table"]), _t(p["cross_w"]), _t(p["cross_b"])
    Ws, bs, Wl, bl = [_t(w) for w in p["W"]], [_t(b) for b in p["b"]], _t(p["Wl"]), _t(p["bl"])
    e, rows, eff = _torch_embed(T, case)
    x0 = e.reshape(40, -1)
    xl = x0
    for l in range(L):
        xl = x0 * (xl @ cw[l])[:, None] + cb[l] + xl
    h = x0
    for W, b in zip(Ws, bs):
        h = torch.relu(h @ W + b)
    logits = (torch.cat([xl, h], -1) @ Wl + bl)[:, 0]
    loss = torch.nn.functional.binary_cross_entropy_with_logits(logits, torch.as_tensor(labels), reduction="mean")
    loss.backward()

    def clipped(g):
        g = g.numpy()
        return g * (clip / max(np.sqrt((g * g).sum()), clip))

    st = FM.adagrad_state(p)
    p0 = {k: ([a.copy() for a in v] if isinstance(v, list) else np.copy(v)) for k, v in p.items()}
    out = FM.dcn_train_step(p, st, case["off"], case["idx"], case["val"], labels, 0.05, clip_norm=clip)
    assert np.allclose(out["logits"], logits.detach().numpy(), rtol=1e-12, atol=1e-12)
    touched = out["rows"]
    # IndexedSlices clip, ONE VARIABLE PER COLUMN (deepFM.py:385-390): each column's norm runs over its own
    # de-duplicated row gradients = the rows of that column in autograd's dense gradient
    want_T = p0["table"].copy()
    tg = T.grad.numpy().copy()
    off = list(case["off"]) + [case["N"]]
    bit = 0
    for f in range(len(case["off"])):
        seg = tg[off[f]:off[f + 1]]
        nrm = np.sqrt((seg * seg).sum())
        bit += nrm > clip
        tg[off[f]:off[f + 1]] = seg * (clip / max(nrm, clip))
    assert bit >= 1, "the clip must bind on at least one column"
    want_T[touched] = _adagrad(p0["table"][touched], tg[touched], 0.05)
    assert np.allclose(p["table"], want_T, rtol=1e-9, atol=1e-12)
    assert np.allclose(p["cross_w"], _adagrad(p0["cross_w"], clipped(cw.grad), 0.05), rtol=1e-9, atol=1e-12)
    assert np.allclose(p["cross_b"], _adagrad(p0["cross_b"], clipped(cb.grad), 0.05), rtol=1e-9, atol=1e-12)
    for i in range(len(Ws)):
        assert np.allclose(p["W"][i], _adagrad(p0["W"][i], clipped(Ws[i].grad), 0.05), rtol=1e-9, atol=1e-12)
    assert np.allclose(p["Wl"], _adagrad(p0["Wl"], clipped(Wl.grad), 0.05), rtol=1e-9, atol=1e-12)
    assert np.allclose(p["bl"], _adagrad(p0["bl"], clipped(bl.grad), 0.05), rtol=1e-9, atol=1e-12)
