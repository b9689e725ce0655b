"""Shared helpers for the GPU parity tests: build a case on the host (numpy), run the CUDA path
through the public layer / C ABI, and hand back numpy results to compare with the oracle."""
import numpy as np
import torch

from oracle import deepctr_oracle as O

REL = 1e-5   # north_star: logits and gradients within 1e-5 relative in fp32


def rel_err(x, ref, floor):
    """max |x-ref| / max(|ref|, floor): `floor` is the magnitude being cancelled."""
    x, ref = np.asarray(x, np.float64), np.asarray(ref, np.float64)
    return float(np.max(np.abs(x - ref) / np.maximum(np.abs(ref), floor))) if x.size else 0.0


def make_case(seed, B, rows, K, weighted=True, prune=False, skew=None):
    rng = np.random.default_rng(seed)
    rows = np.asarray(rows, dtype=np.int64)
    F = len(rows)
    off = np.concatenate([[0], np.cumsum(rows)[:-1]]).astype(np.int64)
    N = int(rows.sum())
    table = (rng.standard_normal((N, K)) / np.sqrt(K)).astype(np.float32)
    w1 = (rng.standard_normal(N) * 0.1).astype(np.float32)
    if skew:
        idx = np.stack([np.minimum((r * rng.random(B) ** skew).astype(np.int64), r - 1) for r in rows], 1)
    else:
        idx = np.stack([rng.integers(0, r, size=B) for r in rows], 1).astype(np.int64)
    idx = idx.reshape(B, F)
    val = None
    if weighted:
        val = (rng.random((B, F)) + 0.25).astype(np.float32)
        if prune and B * F >= 8:
            flat = val.reshape(-1)
            flat[rng.integers(0, B * F, size=max(1, B * F // 16))] = 0.0
            flat[rng.integers(0, B * F, size=max(1, B * F // 32))] = -0.5
    if prune and B * F >= 8:
        idx.reshape(-1)[rng.integers(0, B * F, size=max(1, B * F // 16))] = -1
    return dict(rows=rows, off=off, N=N, K=K, F=F, B=B, table=table, w1=w1, idx=idx, val=val, rng=rng)


def make_layer(pkg, case, optimizer="adagrad", lr=0.05, device="cuda", **kw):
    layer = pkg.EmbeddingFM(case["F"], case["K"], [int(r) for r in case["rows"]], optimizer=optimizer,
                            lr=lr, device=device, **kw)
    layer.load_tables(case["table"], case["w1"])
    return layer


def to_dev(a, device="cuda"):
    return None if a is None else torch.as_tensor(a).to(device)


def oracle_forward(case, bias=0.0, dtype=np.float32):
    t = case["table"].astype(dtype)
    e, keep = O.embedding_lookup(t, case["off"], case["idx"], case["val"], "sum", dtype)
    first = O.first_order(case["w1"].astype(dtype), bias, case["off"], case["idx"], case["val"], dtype)
    fm = O.fm_second_order(e)
    return e, first, fm, keep


def torch_embedding_reference(T0, w1_0, bias, field_offset, idx, val, g_first=None, g_fm=None, u=None):
    """The layer's forward / backward in fp64 with torch indexing and index_add_, on whatever device the tensors live:
    an independent implementation fast enough for BASELINE.json's full sizes (the numpy oracle is not).  It is itself
    checked against the oracle on a small case (tests/test_host_cpu.py).
    -> dict(e[B,F,K] fp64 from fp32 rows, S, fm[B], first[B], and with gradients given: G[N,K], Gfloor[N,K] (the
    cancellation-aware magnitude sum |x| (|g_fm| sum_f |e_f| + |u|)), g1[N], g1abs[N], touched[N] bool)."""
    B, F = idx.shape
    N, K = T0.shape
    rows = idx + field_offset[None, :]
    keep = (idx >= 0) & (val > 0)
    rows = torch.where(keep, rows, torch.zeros_like(rows))
    x = torch.where(keep, val, torch.zeros_like(val)).double()
    e = T0[rows].double() * x[..., None]
    S = e.sum(1)
    out = dict(e=e, S=S, fm=0.5 * ((S * S) - (e * e).sum(1)).sum(-1),
               first=(w1_0[rows].double() * x).sum(1) + float(bias), keep=keep, rows=rows,
               first_abs=(w1_0[rows].double().abs() * x).sum(1))
    if g_fm is None:
        return out
    gfm = g_fm.double()
    uu = u.reshape(B, F, K).double() if u is not None else torch.zeros_like(e)
    per = x[..., None] * (gfm[:, None, None] * (S[:, None, :] - e) + uu)
    A = e.abs().sum(1)
    per_floor = x[..., None] * (gfm.abs()[:, None, None] * A[:, None, :] + uu.abs())
    flat = rows.reshape(-1)
    G = torch.zeros((N, K), dtype=torch.float64, device=T0.device).index_add_(0, flat, per.reshape(-1, K))
    Gfloor = torch.zeros((N, K), dtype=torch.float64, device=T0.device).index_add_(0, flat, per_floor.reshape(-1, K))
    t1 = (x * g_first.double()[:, None]).reshape(-1)
    g1 = torch.zeros(N, dtype=torch.float64, device=T0.device).index_add_(0, flat, t1)
    g1abs = torch.zeros(N, dtype=torch.float64, device=T0.device).index_add_(0, flat, t1.abs())
    touched = torch.zeros(N, dtype=torch.bool, device=T0.device)
    touched[flat[keep.reshape(-1)]] = True
    out.update(G=G, Gfloor=Gfloor, g1=g1, g1abs=g1abs, touched=touched)
    return out
