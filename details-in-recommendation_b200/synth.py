"""Synthetic Criteo-shaped workloads (host side, numpy) for the bench and the tests.

The reference ships no dataset for this path; BASELINE.json's configs name the
shapes.  "39 fields = 26 sparse + 13 dense": every field is embedded to K; a
dense field is a one-row table whose feature_value carries the number (the
weighted-column / 'sum' semantics of dataset/SequenceTensorFlowDataset/test4.py:50-55).
Seeds follow SURVEY.md section 8d: ids 1234, values 1235, tables 1236, labels 1237,
upstream grads 1238.
"""
from dataclasses import dataclass
from typing import Optional

import numpy as np

SEED_IDS, SEED_VALUES, SEED_TABLES, SEED_LABELS, SEED_UPSTREAM = 1234, 1235, 1236, 1237, 1238


@dataclass(frozen=True)
class Workload:
    name: str
    batch: int
    embedding_size: int
    rows_per_field: tuple            # N_f per field, sparse first then dense (1-row) fields
    n_dense: int = 13
    ids: str = "uniform"             # "uniform" | "zipf"
    zipf_a: float = 1.1
    cross_layer_num: int = 0

    @property
    def field_size(self) -> int:
        return len(self.rows_per_field)

    @property
    def n_rows(self) -> int:
        return int(sum(self.rows_per_field))

    @property
    def field_offset(self) -> np.ndarray:
        off = np.zeros(self.field_size, dtype=np.int64)
        off[1:] = np.cumsum(np.asarray(self.rows_per_field, dtype=np.int64))[:-1]
        return off


def _criteo(n_sparse, rows_each, n_dense):
    return tuple([rows_each] * n_sparse + [1] * n_dense)


def terabyte_rows_per_field(total=880_000_000):
    """cfg4: a fixed synthetic cardinality list DEFINED BY THIS REPO (not Criteo's):
    three fields > 1e8, ten fields < 100, the rest in between, 26 sparse fields."""
    big = [292_000_000, 227_000_000, 187_000_000]
    mid = [40_790_948, 39_979_771, 25_641_295, 20_265_000, 12_972_000, 9_758_201,
           7_267_859, 5_461_306, 3_067_956, 1_333_352, 590_152, 405_282, 142_572]
    small = [97, 63, 36, 27, 14, 10, 7, 4, 3, 3]
    rows = big + mid + small
    assert len(rows) == 26
    rows[3] += total - sum(rows) if sum(rows) < total else 0
    return tuple(rows)


def cfg(name: str, batch: Optional[int] = None) -> Workload:
    """BASELINE.json configs by number (cfg1..cfg5) plus reduced test shapes."""
    if name == "cfg1":      # reference path: B=1024, F=39, K=8, 26 x 10 000 + 13 x 1
        return Workload("cfg1", batch or 1024, 8, _criteo(26, 10_000, 13))
    if name == "cfg2":      # DeepFM embedding+FM, 10 M rows, K=16, B=65 536
        return Workload("cfg2", batch or 65_536, 16, _criteo(26, 384_615, 13))
    if name == "cfg3":      # DCN: same lookup + 6 cross layers on d = F*K = 624
        return Workload("cfg3", batch or 65_536, 16, _criteo(26, 384_615, 13), cross_layer_num=6)
    if name == "cfg4":      # Terabyte-sized tables, row-sharded
        return Workload("cfg4", batch or 65_536, 16, terabyte_rows_per_field() + (1,) * 13)
    if name == "cfg5":      # Zipf(1.1) skew stress, B=262 144
        return Workload("cfg5", batch or 262_144, 16, _criteo(26, 384_615, 13), ids="zipf")
    raise ValueError("unknown workload %r" % name)


def zipf_ids(rng, n_rows, size, a=1.1):
    """Zipf(a) bounded to n_rows by inverse CDF on cumulative 1/r^a; rank 1 = row 0."""
    if n_rows == 1:
        return np.zeros(size, dtype=np.int64)
    cdf = np.cumsum(1.0 / np.arange(1, n_rows + 1, dtype=np.float64) ** a)
    cdf /= cdf[-1]
    return np.minimum(np.searchsorted(cdf, rng.random(size), side="left"), n_rows - 1).astype(np.int64)


def make_inputs(w: Workload, batch: Optional[int] = None, seed_shift: int = 0):
    """feature_index[B,F] int64 (per-field local ids), feature_value[B,F] fp32, labels[B] fp32."""
    B = batch or w.batch
    F = w.field_size
    r_ids = np.random.Generator(np.random.PCG64(SEED_IDS + seed_shift))
    r_val = np.random.Generator(np.random.PCG64(SEED_VALUES + seed_shift))
    r_lab = np.random.Generator(np.random.PCG64(SEED_LABELS + seed_shift))
    idx = np.empty((B, F), dtype=np.int64)
    for f, n in enumerate(w.rows_per_field):
        if n == 1:
            idx[:, f] = 0
        elif w.ids == "zipf":
            idx[:, f] = zipf_ids(r_ids, n, B, w.zipf_a)
        else:
            idx[:, f] = r_ids.integers(0, n, size=B, dtype=np.int64)
    val = np.ones((B, F), dtype=np.float32)
    nd = w.n_dense
    if nd:
        val[:, F - nd:] = r_val.random((B, nd), dtype=np.float32)
    labels = (r_lab.random(B) < 0.25).astype(np.float32)
    return idx, val, labels


def make_tables(w: Workload, dtype=np.float32):
    """table[N,K] ~ N(0, 1/sqrt(K)) clipped to 2 sigma (embedding_column init [TF]);
    w1[N] ~ N(0, 0.01) (TF's zero init would make parity trivial)."""
    rng = np.random.Generator(np.random.PCG64(SEED_TABLES))
    K = w.embedding_size
    sd = 1.0 / np.sqrt(K)
    table = np.clip(rng.standard_normal((w.n_rows, K), dtype=np.float32) * sd, -2 * sd, 2 * sd)
    w1 = rng.standard_normal(w.n_rows, dtype=np.float32) * np.float32(0.01)
    return table.astype(dtype), w1.astype(dtype)


def make_upstream(w: Workload, batch: Optional[int] = None, width: Optional[int] = None):
    """u ~ N(0, 1e-2): stand-in for dL/d(embeddings) from the DNN / cross consumer."""
    B = batch or w.batch
    rng = np.random.Generator(np.random.PCG64(SEED_UPSTREAM))
    width = width or w.field_size * w.embedding_size
    return rng.standard_normal((B, width), dtype=np.float32) * np.float32(1e-2)
