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


def main():
    np.savez_compressed(os.path.join(HERE, "deepfm_cfg1_small.npz"),
                        **deepfm_case(20261017, 64, [50] * 26 + [1] * 13, 8))     # cfg1's 39 fields, K = 8
    np.savez_compressed(os.path.join(HERE, "deepfm_k16_skew.npz"),
                        **deepfm_case(20261018, 96, [7, 1, 300, 2, 1, 33], 16))   # heavy duplicates, K = 16
    np.savez_compressed(os.path.join(HERE, "dcn_d312_l3.npz"), **dcn_case(20261019, 48, 312, 3))
    np.savez_compressed(os.path.join(HERE, "dcn_d51_l2.npz"), **dcn_case(20261020, 33, 51, 2))  # census width, default L
    for f in sorted(os.listdir(HERE)):
        if f.endswith(".npz"):
            print(f, os.path.getsize(os.path.join(HERE, f)), "bytes")


if __name__ == "__main__":
    main()
