"""GPU parity of the row-sharded path.  With one process (world_size 1) every kernel of the sharded
pipeline runs -- composite keys, sort, distinct-row numbering, owner gather, forward on the exchanged
buffer, per-row gradient sums, owner-side merge + update -- and must match the oracle like the
single-GPU layer does.  With two GPUs (skipped otherwise) two NCCL ranks each feed their own samples
and the union of their shards must match the oracle run on the concatenated batch."""
import os
import socket

import numpy as np
import pytest
import torch

from oracle import deepctr_oracle as O
from tests._util import REL, make_case, oracle_forward, rel_err, to_dev

pytestmark = pytest.mark.gpu

SHAPES = [(1, [5], 4), (64, [50, 1, 9, 1000, 3, 17, 1], 16), (257, [100] * 26 + [1] * 13, 16),
          (1024, [10_000] * 26 + [1] * 13, 8), (130, [33] * 70, 32), (5000, [3, 1], 16)]


def _layer(pkg, case, optimizer="adagrad", lr=0.05):
    layer = pkg.ShardedEmbeddingFM(case["F"], case["K"], [int(r) for r in case["rows"]], optimizer=optimizer,
                                   lr=lr).train()
    layer.load_tables(case["table"], case["w1"])
    return layer


@pytest.mark.parametrize("B,rows,K", SHAPES)
@pytest.mark.parametrize("optimizer", ["adagrad", "sgd"])
def test_world1_matches_oracle(pkg, cuda, B, rows, K, optimizer):
    case = make_case(31, B, rows, K, weighted=True, prune=True)
    rng, F = case["rng"], case["F"]
    g_first = rng.standard_normal(B).astype(np.float32)
    g_fm = (rng.standard_normal(B) * 0.1).astype(np.float32)
    u = (rng.standard_normal((B, F, K)) * 0.1).astype(np.float32)
    layer = _layer(pkg, case, optimizer)
    first, fm, emb = layer(to_dev(case["idx"]), to_dev(case["val"]))
    e, first_o, fm_o, _ = oracle_forward(case)
    assert np.array_equal(emb.detach().cpu().numpy().reshape(B, F, K), e), "gathered rows must be bit-exact"
    e64, first64, fm64, _ = oracle_forward(case, dtype=np.float64)
    assert rel_err(fm.detach().cpu().numpy(), fm64, 0.5 * (e64 ** 2).sum((1, 2))[:, None] + 1e-30) <= REL
    assert rel_err(first.detach().cpu().numpy(), first64, np.abs(case["w1"]).max() * F + 1e-3) <= REL
    loss = (first[:, 0] * to_dev(g_first)).sum() + (fm[:, 0] * to_dev(g_fm)).sum() + (emb * to_dev(u.reshape(B, -1))).sum()
    loss.backward()
    torch.cuda.synchronize()
    t, w = case["table"].astype(np.float64), case["w1"].astype(np.float64)
    acc, acc1 = np.full_like(t, 0.1), np.full_like(w, 0.1)
    urows, G, g1, _, Gabs, g1abs = O.embedding_backward(t, case["off"], case["idx"], case["val"], g_first, g_fm, u,
                                                        "sum", np.float64, return_abs=True)
    if optimizer == "adagrad":
        O.sparse_adagrad(t, acc, urows, G, 0.05)
        O.sparse_adagrad(w, acc1, urows, g1, 0.05)
    else:
        O.sparse_sgd(t, urows, G, 0.05)
        O.sparse_sgd(w, urows, g1, 0.05)
    got_t, got_w = layer.table.cpu().numpy(), layer.w1.cpu().numpy()
    assert int(layer.last_n_unique.item()) == len(urows)
    untouched = np.ones(case["N"], bool)
    untouched[urows] = False
    assert np.array_equal(got_t[untouched], case["table"][untouched])
    assert rel_err(got_t[urows], t[urows], np.abs(case["table"]).max()) <= REL
    assert rel_err(got_w[urows], w[urows], np.abs(case["w1"]).max() + 1e-3) <= REL
    if optimizer == "adagrad":
        floor = 0.1 + 2 * np.abs(G) * Gabs
        assert rel_err(layer.accum.cpu().numpy()[urows], acc[urows], floor) <= REL
    ex = layer.last_exchange
    keep = (case["idx"] >= 0) & (case["val"] > 0)
    assert ex["unique_sent"] == len(urows) and ex["lookups"] == B * F and int(keep.sum()) >= len(urows)


def _free_port():
    s = socket.socket()
    s.bind(("127.0.0.1", 0))
    p = s.getsockname()[1]
    s.close()
    return p


def _rank_main(rank, world, port, out, exchange_mode):
    import torch.distributed as dist
    import dir_b200
    os.environ.update(MASTER_ADDR="127.0.0.1", MASTER_PORT=str(port), DIR_B200_EXCHANGE=exchange_mode)
    torch.cuda.set_device(rank)
    dist.init_process_group("nccl", rank=rank, world_size=world, device_id=torch.device("cuda", rank))
    try:
        rows, K, Bl = [50, 1, 9, 1000, 3, 17, 1], 16, 96
        case = make_case(41, Bl * world, rows, K, weighted=True, prune=True, skew=2.0)   # global batch
        rng, F = case["rng"], case["F"]
        g_first = rng.standard_normal(Bl * world).astype(np.float32)
        g_fm = (rng.standard_normal(Bl * world) * 0.1).astype(np.float32)
        u = (rng.standard_normal((Bl * world, F, K)) * 0.1).astype(np.float32)
        sl = slice(rank * Bl, (rank + 1) * Bl)
        layer = dir_b200.ShardedEmbeddingFM(F, K, rows, optimizer="adagrad", lr=0.05, max_batch=Bl,
                                            device="cuda").train()
        assert (layer.peer is not None) == (exchange_mode == "peer"), "exchange mode not honoured"
        layer.load_tables(case["table"], case["w1"])
        d = "cuda"
        first, fm, emb = layer(to_dev(case["idx"][sl], d), to_dev(case["val"][sl], d))
        e, first_o, fm_o, _ = oracle_forward(case)
        assert np.array_equal(emb.detach().cpu().numpy().reshape(Bl, F, K), e[sl])
        e64, first64, fm64, _ = oracle_forward(case, dtype=np.float64)
        assert rel_err(fm.detach().cpu().numpy(), fm64[sl], 0.5 * (e64[sl] ** 2).sum((1, 2))[:, None] + 1e-30) <= REL
        loss = ((first[:, 0] * to_dev(g_first[sl], d)).sum() + (fm[:, 0] * to_dev(g_fm[sl], d)).sum()
                + (emb * to_dev(u[sl].reshape(Bl, -1), d)).sum())
        loss.backward()
        torch.cuda.synchronize()
        shards = [torch.empty_like(layer.rows) for _ in range(world)]
        lins = [torch.empty_like(layer.lin_rows) for _ in range(world)]
        dist.all_gather(shards, layer.rows)
        dist.all_gather(lins, layer.lin_rows)
        if rank == 0:
            N = case["N"]
            full = np.zeros((layer.plan.cap * world, 2 * K), np.float32)
            fw = np.zeros(layer.plan.cap * world, np.float32)
            for r in range(world):
                full[r::world] = shards[r].cpu().numpy()
                fw[r::world] = lins[r].cpu().numpy()[:, 0]
            t, w = case["table"].astype(np.float64), case["w1"].astype(np.float64)
            acc, acc1 = np.full_like(t, 0.1), np.full_like(w, 0.1)
            urows, G, g1, _ = O.embedding_backward(t, case["off"], case["idx"], case["val"], g_first, g_fm, u,
                                                   "sum", np.float64)
            O.sparse_adagrad(t, acc, urows, G, 0.05)
            O.sparse_adagrad(w, acc1, urows, g1, 0.05)
            assert rel_err(full[:N, :K], t, np.abs(case["table"]).max()) <= REL
            assert rel_err(fw[:N], w, np.abs(case["w1"]).max() + 1e-3) <= REL
            untouched = np.ones(N, bool)
            untouched[urows] = False
            assert np.array_equal(full[:N, :K][untouched], case["table"][untouched])
        out.put((rank, "ok"))
    except Exception as ex:
        import traceback
        out.put((rank, traceback.format_exc()))
    finally:
        dist.destroy_process_group()


@pytest.mark.parametrize("world", [2, 4, 8])
@pytest.mark.parametrize("exchange_mode", ["peer", "nccl"])
def test_multi_rank_nccl_matches_oracle(pkg, cuda, exchange_mode, world):
    if torch.cuda.device_count() < world:
        pytest.skip("needs %d GPUs (gpurun --gpus %d)" % (world, world))
    import torch.multiprocessing as mp
    ctx = mp.get_context("spawn")
    out = ctx.Queue()
    port = _free_port()
    procs = [ctx.Process(target=_rank_main, args=(r, world, port, out, exchange_mode)) for r in range(world)]
    for p in procs:
        p.start()
    res = [out.get(timeout=300) for _ in procs]
    for p in procs:
        p.join(timeout=60)
    assert sorted(res) == [(r, "ok") for r in range(world)], res
