"""BASELINE.json configs[3] at FULL size on one B200: 880 M rows, K = 16, fp32 rows + Adagrad accumulators = 113 GB
of the 180 GB.  The tables are filled on the device from a counter hash (`dir_table_init_counter`), so the oracle can
reproduce any row without the table (`oracle.deepctr_oracle.counter_rows`): gathered rows must be bit-exact, and after
one step the touched rows must follow the oracle's Adagrad rule.  World size 1 runs every kernel of the sharded path
against local buffers; the multi-rank layout of the same tables is exercised by tests/test_gpu_sharded.py and by
`bench.py --workload cfg4 --gpus N`."""
import numpy as np
import pytest
import torch

from oracle import deepctr_oracle as O
from tests._util import REL

pytestmark = pytest.mark.gpu


def test_counter_init_matches_oracle_small(pkg, cuda):
    rows, K = [1000, 1, 37, 5000, 1], 8
    layer = pkg.ShardedEmbeddingFM(len(rows), K, rows, init="counter")
    want = O.counter_rows(np.arange(sum(rows)), K)
    assert np.array_equal(layer.table.cpu().numpy(), want)
    # the replicated one-row fields carry the owner's rows
    off = np.concatenate([[0], np.cumsum(rows)[:-1]])
    assert np.array_equal(layer.dense_table.cpu().numpy(), want[off[[1, 4]]])


def test_cfg4_full_size_spot_check(pkg, cuda):
    torch.cuda.empty_cache()
    free, _ = torch.cuda.mem_get_info()
    if free < 150e9:
        pytest.skip("needs ~130 GB of free HBM (%.0f GB free)" % (free / 1e9))
    w = pkg.synth.cfg("cfg4", batch=8192)
    F, K, B = w.field_size, w.embedding_size, 8192
    layer = pkg.ShardedEmbeddingFM(F, K, list(w.rows_per_field), optimizer="adagrad", lr=0.05, max_batch=B,
                                   init="counter").train()
    assert layer.plan.n_rows == 880_000_013 and layer.rows.shape == (880_000_013, 2 * K)
    idx, val, _ = pkg.synth.make_inputs(w, batch=B)
    off = w.field_offset
    grow = idx + off[None, :]
    first, fm, emb = layer(torch.as_tensor(idx).cuda(), torch.as_tensor(val).cuda())
    t0 = O.counter_rows(grow.reshape(-1), K).reshape(B, F, K)
    e_want = t0 * val[:, :, None]
    assert np.array_equal(emb.detach().cpu().numpy().reshape(B, F, K), e_want), "rows of the 880 M-row table"
    gen = np.random.default_rng(3)
    g_first = gen.standard_normal(B).astype(np.float32)
    g_fm = (gen.standard_normal(B) * 0.1).astype(np.float32)
    u = (gen.standard_normal((B, F, K)) * 0.1).astype(np.float32)
    torch.autograd.backward((first, fm, emb), (torch.as_tensor(g_first).cuda()[:, None],
                                                torch.as_tensor(g_fm).cuda()[:, None],
                                                torch.as_tensor(u.reshape(B, -1)).cuda()))
    torch.cuda.synchronize()
    layer.check_errors()
    # oracle on the compact table of the rows this batch touches
    urows, inv = np.unique(grow.reshape(-1), return_inverse=True)
    compact = O.counter_rows(urows, K).astype(np.float64)
    cidx = inv.reshape(B, F).astype(np.int64)
    rows_c, G, g1, _, Gabs, _ = O.embedding_backward(compact, np.zeros(F, np.int64), cidx, val, g_first, g_fm, u,
                                                     "sum", np.float64, return_abs=True)
    acc = np.full_like(compact, 0.1)
    O.sparse_adagrad(compact, acc, rows_c, G, 0.05)
    assert int(layer.last_n_unique.item()) == len(urows)
    got = layer.rows[torch.as_tensor(urows).cuda()].cpu().numpy().astype(np.float64)
    f = lambda g: g / np.sqrt(0.1 + g * g)
    d = REL * Gabs
    band = 0.05 * np.maximum(np.abs(f(G + d) - f(G)), np.abs(f(G - d) - f(G))) + 2.0 ** -21 * (np.abs(compact[rows_c]) + 0.05)
    assert (np.abs(got[rows_c, :K] - compact[rows_c]) <= band).all()
    assert (np.abs(got[rows_c, K:] - acc[rows_c]) <= REL * (0.1 + 2 * np.abs(G) * Gabs)).all()
    # neighbours of touched rows are untouched (bit-identical to the initialiser)
    nb = np.setdiff1d(np.minimum(urows + 1, 880_000_012), urows)[:4096]
    assert np.array_equal(layer.table[torch.as_tensor(nb).cuda()].cpu().numpy(), O.counter_rows(nb, K))
    del layer
    torch.cuda.empty_cache()
