"""Host-side logic of the row-sharded path under gloo, world_size 2, on CPU: the shard plan's index
arithmetic and the two variable-size all-to-alls (ids out / rows back, gradients out), with torch
indexing standing in for the CUDA gather on the owner.  The result must equal a direct lookup in the
full table, and the owner-side gradient sums must equal the oracle's per-row sums."""
import os
import socket

import numpy as np
import pytest
import torch
import torch.distributed as dist
import torch.multiprocessing as mp

from oracle import deepctr_oracle as O


def _free_port():
    s = socket.socket()
    s.bind(("127.0.0.1", 0))
    p = s.getsockname()[1]
    s.close()
    return p


def _worker(rank, world, port, out):
    import dir_b200
    from dir_b200.sharded import ShardPlan, exchange, exchange_counts
    os.environ.update(MASTER_ADDR="127.0.0.1", MASTER_PORT=str(port))
    dist.init_process_group("gloo", rank=rank, world_size=world)
    try:
        rows = [7, 1, 12, 3]
        K, B = 4, 9
        rng = np.random.default_rng(5)                       # same on every rank
        N = sum(rows)
        table = rng.standard_normal((N, K)).astype(np.float32)
        plan = ShardPlan(rows, world, rank)
        assert plan.cap == (N + world - 1) // world
        local_table = torch.as_tensor(plan.shard_of(table).copy())
        assert local_table.shape[0] == plan.n_local
        # this rank's samples
        r2 = np.random.default_rng(100 + rank)
        idx = np.stack([r2.integers(0, r, size=B) for r in rows], 1).astype(np.int64)
        grow = torch.as_tensor(idx + np.asarray(plan.field_offset)[None, :]).reshape(-1)
        owner, local = plan.owner(grow), plan.local_row(grow)
        assert torch.equal(plan.global_row(local, owner), grow)
        # distinct (owner, local) keys, owner-major: what dir_shard_keys/sort/unique produce on the GPU
        key = owner * plan.cap + local
        ukeys, inv = torch.unique(key, sorted=True, return_inverse=True)
        send_counts = torch.bincount(ukeys // plan.cap, minlength=world)
        recv_counts = exchange_counts(send_counts)
        ss, rs = send_counts.tolist(), recv_counts.tolist()
        recv_ids = exchange((ukeys % plan.cap).to(torch.int32), ss, rs)
        assert recv_ids.numel() == 0 or int(recv_ids.max()) < plan.n_local
        # the fixed-address variant used by the static (CUDA-graph) mode lands the same ids
        from dir_b200.sharded import exchange_into, peer_offsets
        landing = torch.full((N,), -1, dtype=torch.int32)
        view = exchange_into(landing, (ukeys % plan.cap).to(torch.int32), ss, rs)
        assert torch.equal(view, recv_ids) and view.data_ptr() == landing.data_ptr()
        with pytest.raises(ValueError):
            exchange_into(torch.empty(0, dtype=torch.int32), (ukeys % plan.cap).to(torch.int32), ss, rs)
        # peer-memory segment offsets from everybody's counts: the segments tile each buffer exactly
        rows_m = [torch.empty_like(send_counts) for _ in range(world)]
        dist.all_gather(rows_m, send_counts)
        M = torch.stack(rows_m)                               # M[q, o]
        recv_off, fwd_dst_off, bwd_dst_off = peer_offsets(M, rank)
        assert recv_off.tolist() == [0] + torch.cumsum(recv_counts, 0).tolist()
        offs = [peer_offsets(M, r) for r in range(world)]
        for q in range(world):                                # requester q's row buffer, grouped by owner
            starts = [int(offs[o][1][q]) for o in range(world)]
            assert starts == [int(M[q, :o].sum()) for o in range(world)]
        for o in range(world):                                # owner o's gradient buffer, grouped by requester
            starts = [int(offs[q][2][o]) for q in range(world)]
            assert starts == [int(offs[o][0][q]) for q in range(world)]
        answer = local_table[recv_ids.long()]                 # stand-in for dir_rows_gather
        ubuf = exchange(answer, rs, ss)
        e = ubuf[inv].reshape(B, len(rows), K).numpy()
        want, _ = O.embedding_lookup(table, plan.field_offset, idx, None)
        assert np.array_equal(e, want), "rows routed through the exchange differ from a direct lookup"
        # backward: per-distinct-row sums out, the owner adds the ranks' contributions
        g = torch.as_tensor(r2.standard_normal((B * len(rows), K)).astype(np.float32))
        gu = torch.zeros((ukeys.numel(), K)).index_add_(0, inv, g)
        grecv = exchange(gu, ss, rs)
        G_local = torch.zeros((plan.cap, K)).index_add_(0, recv_ids.long(), grecv)
        # gather every rank's pieces on rank 0 and compare with a global scatter-add
        pieces = [torch.zeros((plan.cap, K)) for _ in range(world)]
        dist.all_gather(pieces, G_local)
        all_rows = [torch.zeros_like(grow) for _ in range(world)]
        all_g = [torch.zeros_like(g) for _ in range(world)]
        dist.all_gather(all_rows, grow)
        dist.all_gather(all_g, g)
        if rank == 0:
            full = torch.zeros((plan.cap * world, K))
            for r in range(world):
                full[r::world] = pieces[r]
            ref = torch.zeros((plan.cap * world, K)).index_add_(0, torch.cat(all_rows), torch.cat(all_g))
            assert torch.allclose(full, ref, atol=1e-5)
        out.put((rank, "ok"))
    except Exception as ex:                                   # surface the failure in the parent
        out.put((rank, repr(ex)))
    finally:
        dist.destroy_process_group()


def test_exchange_world2_gloo():
    world = 2
    ctx = mp.get_context("spawn")
    out = ctx.Queue()
    port = _free_port()
    procs = [ctx.Process(target=_worker, args=(r, world, port, out)) for r in range(world)]
    for p in procs:
        p.start()
    res = [out.get(timeout=120) for _ in procs]
    for p in procs:
        p.join(timeout=60)
    assert sorted(res) == [(0, "ok"), (1, "ok")], res


def test_shard_plan_arithmetic():
    from dir_b200.sharded import ShardPlan
    rows = [5, 1, 9]
    for world in (1, 2, 3, 8):
        seen = []
        for rank in range(world):
            p = ShardPlan(rows, world, rank)
            assert p.n_rows == 15 and p.cap == -(-15 // world)
            mine = [r for r in range(15) if r % world == rank]
            assert p.n_local == len(mine)
            assert [int(p.global_row(torch.tensor(l))) for l in range(p.n_local)] == mine
            seen += mine
        assert sorted(seen) == list(range(15))
    with pytest.raises(ValueError):
        ShardPlan(rows, 2, 2)
    full = np.arange(30).reshape(15, 2)
    assert np.array_equal(ShardPlan(rows, 4, 1).shard_of(full), full[1::4])
