"""Multi-hot / weighted bags through the row-sharded tables (ShardedEmbeddingBagFM: dir_shard_bag_keys,
dir_embed_bag_fm_fwd over the exchanged rows, dir_embed_bag_bwd_reduce_emit_to) against the oracle: one rank on
one GPU, and 2 / 4 ranks with the batch split over them (models/DeepFM/deepFM.py:53, 77 x :163-175)."""
import os
import socket

import numpy as np
import pytest
import torch

from oracle import deepctr_oracle as O
from tests._util import REL, rel_err, to_dev
from tests.test_gpu_zbags import _bags, _tables

pytestmark = pytest.mark.gpu


def _expected(table, w1, off_f, off, idx, w, B, F, combiner, g_first, g_fm, u, lr=0.05):
    t64, w64 = table.astype(np.float64), w1.astype(np.float64)
    e64, first64, x64 = O.embedding_bag_lookup(t64, w64, 0.0, off_f, off, idx, w, B, F, combiner, np.float64)
    urows, G, g1 = O.embedding_bag_backward(t64, off_f, off, idx, w, e64, x64, g_first, g_fm, u, B, F, np.float64)
    acc, acc1 = np.full_like(t64, 0.1), np.full_like(w64, 0.1)
    O.sparse_adagrad(t64, acc, urows, G, lr)
    O.sparse_adagrad(w64, acc1, urows, g1, lr)
    return e64, first64, t64, w64, urows


@pytest.mark.parametrize("B,rows,K,max_len", [(33, [7, 1, 30, 4], 8, 4), (257, [100] * 10 + [1] * 3, 16, 5),
                                              (64, [50, 9, 1000], 32, 12), (300, [3, 2], 16, 40)])
@pytest.mark.parametrize("combiner", ["sum", "mean", "sqrtn"])
def test_world1_bags_match_oracle(pkg, cuda, B, rows, K, max_len, combiner):
    rng = np.random.default_rng(29)
    F = len(rows)
    off_f = np.concatenate([[0], np.cumsum(rows)[:-1]]).astype(np.int64)
    table, w1 = _tables(rng, rows, K)
    off, idx, w = _bags(rng, B, rows, max_len, weighted=True, skew=2.0)
    layer = pkg.ShardedEmbeddingBagFM(F, K, rows, combiner=combiner, optimizer="adagrad", lr=0.05, max_batch=B,
                                      max_entries=max(len(idx), 1)).train()
    assert layer.n_dense == 0, "bags send every field through the exchange"
    layer.load_tables(table, w1)
    first, fm, emb = layer.forward_bags(to_dev(off), to_dev(idx), to_dev(w))
    e32, _, _ = O.embedding_bag_lookup(table, w1, 0.0, off_f, off, idx, w, B, F, combiner)
    assert np.array_equal(emb.detach().cpu().numpy().reshape(B, F, K), e32), "combined embeddings must be bit-exact"
    g_first = rng.standard_normal(B).astype(np.float32)
    g_fm = (rng.standard_normal(B) * 0.1).astype(np.float32)
    u = (rng.standard_normal((B, F, K)) * 0.1).astype(np.float32)
    e64, first64, t64, w64, urows = _expected(table, w1, off_f, off, idx, w, B, F, combiner, g_first, g_fm, u)
    fm64 = O.fm_second_order(e64)
    assert rel_err(fm.detach().cpu().numpy(), fm64, 0.5 * (e64 ** 2).sum((1, 2))[:, None] + 1e-30) <= REL
    assert rel_err(first.detach().cpu().numpy(), first64, np.abs(w1).max() * F * max_len + 0.125) <= REL
    torch.autograd.backward((first, fm, emb), (to_dev(g_first)[:, None], to_dev(g_fm)[:, None], to_dev(u.reshape(B, -1))))
    torch.cuda.synchronize()
    layer.check_errors()
    N = len(w1)
    got_t, got_w = layer.table.cpu().numpy()[:N], layer.w1.cpu().numpy()[:N]
    assert int(layer.last_n_unique.item()) == len(urows)
    untouched = np.ones(N, bool)
    untouched[urows] = False
    assert np.array_equal(got_t[untouched], table[untouched]) and np.array_equal(got_w[untouched], w1[untouched])
    assert rel_err(got_t[urows], t64[urows], np.abs(table).max()) <= REL
    assert rel_err(got_w[urows], w64[urows], np.abs(w1).max() + 1e-3) <= REL
    # a second call from the updated state (the other pair of exchange buffers); eval mode leaves the rows alone
    layer.eval()
    before = layer.rows.clone()
    _, _, emb2 = layer.forward_bags(to_dev(off), to_dev(idx), to_dev(w))
    e2, _, _ = O.embedding_bag_lookup(got_t, got_w, 0.0, off_f, off, idx, w, B, F, combiner)
    assert np.array_equal(emb2.detach().cpu().numpy().reshape(B, F, K), e2) and torch.equal(layer.rows, before)


def test_sharded_bag_edge_cases(pkg, cuda):
    layer = pkg.ShardedEmbeddingBagFM(2, 8, [4, 3], combiner="mean", max_batch=8, max_entries=16).train()
    before = layer.rows.clone()
    first, fm, emb = layer.forward_bags(torch.zeros(7, dtype=torch.int64, device="cuda"),
                                        torch.zeros(0, dtype=torch.int64, device="cuda"))      # every bag empty
    assert float(emb.detach().abs().max()) == 0 and float(fm.detach().abs().max()) == 0
    (first.sum() + fm.sum() + emb.sum()).backward()
    torch.cuda.synchronize()
    assert torch.equal(layer.rows, before)
    with pytest.raises(ValueError):     # more entries than the exchange buffers were sized for
        layer.forward_bags(torch.tensor([0, 17, 17], dtype=torch.int64, device="cuda"),
                           torch.zeros(17, dtype=torch.int64, device="cuda"))
    with pytest.raises(RuntimeError):
        layer(torch.zeros((1, 2), dtype=torch.int64, device="cuda"))
    with pytest.raises(ValueError):
        pkg.ShardedEmbeddingBagFM(2, 8, [4, 3], combiner="max")


def _free_port():
    s = socket.socket()
    s.bind(("127.0.0.1", 0))
    p = s.getsockname()[1]
    s.close()
    return p


def _rank_main(rank, world, port, out, combiner):
    import torch.distributed as dist
    import dir_b200
    os.environ.update(MASTER_ADDR="127.0.0.1", MASTER_PORT=str(port))
    torch.cuda.set_device(rank)
    dist.init_process_group("nccl", rank=rank, world_size=world, device_id=torch.device("cuda", rank))
    try:
        rows, K, Bl, max_len = [50, 1, 9, 1000, 3, 17, 1], 16, 48, 6
        F, B = len(rows), 48 * world
        rng = np.random.default_rng(43)                                   # same global batch on every rank
        off_f = np.concatenate([[0], np.cumsum(rows)[:-1]]).astype(np.int64)
        table, w1 = _tables(rng, rows, K)
        off, idx, w = _bags(rng, B, rows, max_len, weighted=True, skew=2.0)
        g_first = rng.standard_normal(B).astype(np.float32)
        g_fm = (rng.standard_normal(B) * 0.1).astype(np.float32)
        u = (rng.standard_normal((B, F, K)) * 0.1).astype(np.float32)
        s0, s1 = rank * Bl * F, (rank + 1) * Bl * F                        # this rank's slots
        j0, j1 = int(off[s0]), int(off[s1])
        off_l, idx_l, w_l = off[s0:s1 + 1] - off[s0], idx[j0:j1], w[j0:j1]
        sl = slice(rank * Bl, (rank + 1) * Bl)
        layer = dir_b200.ShardedEmbeddingBagFM(F, K, rows, combiner=combiner, optimizer="adagrad", lr=0.05,
                                               max_batch=Bl, max_entries=Bl * F * max_len, device="cuda").train()
        layer.load_tables(table, w1)
        d = "cuda"
        first, fm, emb = layer.forward_bags(to_dev(off_l, d), to_dev(idx_l, d), to_dev(w_l, d))
        e32, _, _ = O.embedding_bag_lookup(table, w1, 0.0, off_f, off, idx, w, B, F, combiner)
        assert np.array_equal(emb.detach().cpu().numpy().reshape(Bl, F, K), e32[sl])
        torch.autograd.backward((first, fm, emb), (to_dev(g_first[sl], d)[:, None], to_dev(g_fm[sl], d)[:, None],
                                                   to_dev(u[sl].reshape(Bl, -1), d)))
        torch.cuda.synchronize()
        layer.check_errors()
        shards = [torch.empty_like(layer.rows) for _ in range(world)]
        lins = [torch.empty_like(layer.lin_rows) for _ in range(world)]
        dist.all_gather(shards, layer.rows)
        dist.all_gather(lins, layer.lin_rows)
        if rank == 0:
            N = len(w1)
            full = np.zeros((layer.plan.cap * world, 2 * K), np.float32)
            fw = np.zeros(layer.plan.cap * world, np.float32)
            for r in range(world):
                full[r::world] = shards[r].cpu().numpy()
                fw[r::world] = lins[r].cpu().numpy()[:, 0]
            _, _, t64, w64, urows = _expected(table, w1, off_f, off, idx, w, B, F, combiner, g_first, g_fm, u)
            untouched = np.ones(N, bool)
            untouched[urows] = False
            assert np.array_equal(full[:N, :K][untouched], table[untouched])
            assert rel_err(full[:N, :K][urows], t64[urows], np.abs(table).max()) <= REL
            assert rel_err(fw[:N][urows], w64[urows], np.abs(w1).max() + 1e-3) <= REL
        out.put((rank, "ok"))
    except Exception:
        import traceback
        out.put((rank, traceback.format_exc()))
    finally:
        dist.destroy_process_group()


@pytest.mark.parametrize("world", [2, 4])
@pytest.mark.parametrize("combiner", ["mean", "sum"])
def test_multi_rank_bags_match_oracle(pkg, cuda, combiner, world):
    if torch.cuda.device_count() < world:
        pytest.skip("needs %d GPUs (gpurun --gpus %d)" % (world, world))
    import torch.multiprocessing as mp
    ctx = mp.get_context("spawn")
    out = ctx.Queue()
    port = _free_port()
    procs = [ctx.Process(target=_rank_main, args=(r, world, port, out, combiner)) for r in range(world)]
    for p in procs:
        p.start()
    res = [out.get(timeout=300) for _ in procs]
    for p in procs:
        p.join(timeout=60)
    assert sorted(res) == [(r, "ok") for r in range(world)], res
