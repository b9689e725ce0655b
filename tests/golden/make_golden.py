#!/usr/bin/env python
"""Writes tests/golden/*.npz: seeded inputs (fp32 / int64) and the oracle's fp64 outputs for them.

PARITY UNPINNED: the reference's TF 1.x graph cannot run here and the reference holds no golden
vectors for this path, so these fixtures come from the CPU oracle
(oracle/deepctr_oracle.py), which tests/test_oracle.py pins against closed-form known answers and
torch autograd.  They freeze the oracle (a later edit that changes its results fails
tests/test_golden.py) and give the GPU tests fixed vectors that do not depend on numpy's RNG stream.

    python tests/golden/make_golden.py
"""
import os
import sys

import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, os.path.dirname(os.path.dirname(HERE)))
from oracle import deepctr_oracle as O  # noqa: E402
from oracle import tf_semantics as tfs  # noqa: E402


def deepfm_case(seed, B, rows, K, lr=0.05):
    rng = np.random.Generator(np.random.PCG64(seed))
    rows = np.asarray(rows, dtype=np.int64)
    F, N = len(rows), int(rows.sum())
    off = np.concatenate([[0], np.cumsum(rows)[:-1]]).astype(np.int64)
    sd = 1.0 / np.sqrt(K)
    table = np.clip(rng.standard_normal((N, K)) * sd, -2 * sd, 2 * sd).astype(np.float32)
    w1 = (rng.standard_normal(N) * 0.01).astype(np.float32)
    idx = np.stack([rng.integers(0, r, size=B) for r in rows], 1).astype(np.int64)
    val = np.ones((B, F), dtype=np.float32)
    dense = rows == 1
    val[:, dense] = rng.random((B, int(dense.sum()))).astype(np.float32)
    val[rng.integers(0, B, 3), rng.integers(0, F, 3)] = 0.0          # pruned: weight <= 0
    idx[rng.integers(0, B, 2), rng.integers(0, F, 2)] = -1           # pruned: id < 0
    labels = (rng.random(B) < 0.25).astype(np.float32)
    u = (rng.standard_normal((B, F, K)) * 1e-2).astype(np.float32)
    t, a = table.astype(np.float64), np.full((N, K), 0.1)
    l1, a1 = w1.astype(np.float64), np.full(N, 0.1)
    r = O.deepfm_layer_step(t, a, l1, a1, 0.0, off, idx, val.astype(np.float64), labels, lr, u=u.astype(np.float64),
                            dtype=np.float64)
    return dict(rows=rows, field_offset=off, table=table, w1=w1, idx=idx, val=val, labels=labels, u=u,
                lr=np.float64(lr), first=r["first"], fm=r["fm"], logits=r["logits"], e=r["e"], g=r["g"],
                touched=r["rows"], G=r["G"], g1=r["g1"], table_after=t, accum_after=a, w1_after=l1,
                w1_accum_after=a1)


def dcn_case(seed, B, d, L):
    rng = np.random.Generator(np.random.PCG64(seed))
    x0 = (rng.standard_normal((B, d)) * 0.5).astype(np.float32)
    w = tfs.truncated_normal(rng, (L, d), 0.1)
    b = tfs.truncated_normal(rng, (L, d), 0.1)
    dy = rng.standard_normal((B, d)).astype(np.float32)
    f64 = [a.astype(np.float64) for a in (x0, w, b, dy)]
    xL, s = O.cross_forward(*f64[:3])
    dx0, dw, db = O.cross_backward(*f64)
    return dict(x0=x0, cross_w=w, cross_b=b, dy=dy, xL=xL, s=s, dx0=dx0, dw=dw, db=db)


def bags_case(seed, B, rows, K, combiner, lr=0.05, max_len=4):
    """Multi-hot / weighted bags (EmbeddingBagFM): CSR inputs, one forward + backward + Adagrad step."""
    rng = np.random.Generator(np.random.PCG64(seed))
    rows = np.asarray(rows, dtype=np.int64)
    F, N = len(rows), int(rows.sum())
    off_f = np.concatenate([[0], np.cumsum(rows)[:-1]]).astype(np.int64)
    table = (rng.standard_normal((N, K)) / np.sqrt(K)).astype(np.float32)
    w1 = (rng.standard_normal(N) * 0.1).astype(np.float32)
    lens = rng.integers(0, max_len + 1, size=B * F)
    off = np.concatenate([[0], np.cumsum(lens)]).astype(np.int64)
    field = np.repeat(np.arange(B * F) % F, lens)
    idx = np.minimum((rng.random(off[-1]) ** 2 * rows[field]).astype(np.int64), rows[field] - 1)
    w = (rng.random(off[-1]) + 0.25).astype(np.float32)
    idx[rng.integers(0, off[-1], 4)] = -1
    w[rng.integers(0, off[-1], 4)] = 0.0
    g_first = rng.standard_normal(B).astype(np.float32)
    g_fm = (rng.standard_normal(B) * 0.1).astype(np.float32)
    u = (rng.standard_normal((B, F, K)) * 0.1).astype(np.float32)
    t64, w64 = table.astype(np.float64), w1.astype(np.float64)
    e, first, x = O.embedding_bag_lookup(t64, w64, 0.125, off_f, off, idx, w, B, F, combiner, np.float64)
    e32, _, _ = O.embedding_bag_lookup(table, w1, 0.125, off_f, off, idx, w, B, F, combiner, np.float32)
    touched, G, g1 = O.embedding_bag_backward(t64, off_f, off, idx, w, e, x, g_first, g_fm, u, B, F, np.float64)
    acc, acc1 = np.full_like(t64, 0.1), np.full_like(w64, 0.1)
    O.sparse_adagrad(t64, acc, touched, G, lr)
    O.sparse_adagrad(w64, acc1, touched, g1, lr)
    return dict(rows=rows, field_offset=off_f, table=table, w1=w1, bag_offsets=off, bag_index=idx, bag_weight=w,
                combiner=np.array(combiner), g_first=g_first, g_fm=g_fm, u=u, lr=np.float64(lr), e=e, e32=e32,
                first=first, fm=O.fm_second_order(e), touched=touched, G=G, g1=g1, table_after=t64, w1_after=w64)


def ftrl_case(seed, B, rows, K, lr=0.05, lin_lr=0.2, l1=0.02, l2=0.1, steps=2):
    """Tables on Adagrad, the linear scope on Ftrl (deepFM.py:58-61): `steps` consecutive updates."""
    rng = np.random.Generator(np.random.PCG64(seed))
    rows = np.asarray(rows, dtype=np.int64)
    F, N = len(rows), int(rows.sum())
    off = np.concatenate([[0], np.cumsum(rows)[:-1]]).astype(np.int64)
    table = (rng.standard_normal((N, K)) / np.sqrt(K)).astype(np.float32)
    w1 = (rng.standard_normal(N) * 0.1).astype(np.float32)
    idx = np.stack([np.minimum((r * rng.random(B) ** 2).astype(np.int64), r - 1) for r in rows], 1)
    val = (rng.random((B, F)) + 0.25).astype(np.float32)
    g_first = rng.standard_normal((steps, B)).astype(np.float32)
    g_fm = (rng.standard_normal((steps, B)) * 0.1).astype(np.float32)
    u = (rng.standard_normal((steps, B, F, K)) * 0.1).astype(np.float32)
    t, w = table.astype(np.float64), w1.astype(np.float64)
    acc, n1, z1 = np.full_like(t, 0.1), np.full_like(w, 0.1), np.zeros_like(w)
    for s in range(steps):
        r, G, g1, _ = O.embedding_backward(t, off, idx, val, g_first[s], g_fm[s], u[s], "sum", np.float64)
        O.sparse_adagrad(t, acc, r, G, lr)
        O.sparse_ftrl(w, n1, z1, r, g1, lin_lr, l1, l2)
    return dict(rows=rows, field_offset=off, table=table, w1=w1, idx=idx, val=val, g_first=g_first, g_fm=g_fm, u=u,
                lr=np.float64(lr), lin_lr=np.float64(lin_lr), l1=np.float64(l1), l2=np.float64(l2),
                table_after=t, w1_after=w, n_after=n1, z_after=z1)


CENSUS = [("age", "numeric", 1), ("education_num", "numeric", 1), ("capital_gain", "numeric", 1),
          ("capital_loss", "numeric", 1), ("hours_per_week", "numeric", 1), ("workclass_indicator", "indicator", 9),
          ("education_indicator", "indicator", 16), ("marital_status_indicator", "indicator", 7),
          ("relationship_indicator", "indicator", 6), ("occupation_embedding", "embedding", 8)]


def input_layer_case(seed, B):
    """tf.feature_column.input_layer over the census columns of DeepCrossNetwork/train.py:88-100 (d = 51)."""
    rng = np.random.Generator(np.random.PCG64(seed))
    numeric = rng.standard_normal((B, 5)).astype(np.float32)
    ids = np.stack([rng.integers(-1, s + 1, size=B) for s in (9, 16, 7, 6)], 1).astype(np.int64)
    emb = rng.standard_normal((B, 8)).astype(np.float32)
    x0, where = O.input_layer(CENSUS, numeric, ids, emb)
    return dict(numeric=numeric, indicator_ids=ids, emb=emb, x0=x0,
                occupation_columns=np.asarray(where["occupation_embedding"], dtype=np.int64))


def main():
    np.savez_compressed(os.path.join(HERE, "deepfm_cfg1_small.npz"),
                        **deepfm_case(20261017, 64, [50] * 26 + [1] * 13, 8))     # cfg1's 39 fields, K = 8
    np.savez_compressed(os.path.join(HERE, "deepfm_k16_skew.npz"),
                        **deepfm_case(20261018, 96, [7, 1, 300, 2, 1, 33], 16))   # heavy duplicates, K = 16
    np.savez_compressed(os.path.join(HERE, "dcn_d312_l3.npz"), **dcn_case(20261019, 48, 312, 3))
    np.savez_compressed(os.path.join(HERE, "dcn_d51_l2.npz"), **dcn_case(20261020, 33, 51, 2))  # census width, default L
    np.savez_compressed(os.path.join(HERE, "bags_mean_k16.npz"), **bags_case(20261021, 48, [40, 1, 300, 5], 16, "mean"))
    np.savez_compressed(os.path.join(HERE, "bags_sqrtn_k8.npz"), **bags_case(20261022, 32, [9, 60], 8, "sqrtn"))
    np.savez_compressed(os.path.join(HERE, "ftrl_two_steps.npz"), **ftrl_case(20261023, 80, [30, 1, 200, 4], 16))
    np.savez_compressed(os.path.join(HERE, "input_layer_census.npz"), **input_layer_case(20261024, 40))
    for f in sorted(os.listdir(HERE)):
        if f.endswith(".npz"):
            print(f, os.path.getsize(os.path.join(HERE, f)), "bytes")


if __name__ == "__main__":
    main()
