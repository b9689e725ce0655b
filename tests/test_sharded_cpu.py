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


def _peer_protocol(world, rows, K, B, seed, seg_cap=None):
    """The device-driven exchange of csrc/shard_peer.cu restated with numpy, every rank simulated in one process:
    headers (count, base_u) + id segments, the owner's slot map (slot[row, q] = i + 1, no sort), rows stored at
    base_u + i of the requester, per-distinct-row sums stored at segment q of the owner, the lowest asking rank
    merging in rank order.  Returns (rows each rank's lookups see, per-global-row gradient sums, slot maps)."""
    from dir_b200.sharded import ShardPlan
    rng = np.random.default_rng(seed)
    N = sum(rows)
    F = len(rows)
    table = rng.standard_normal((N, K)).astype(np.float32)
    plans = [ShardPlan(rows, world, r) for r in range(world)]
    cap = plans[0].cap
    seg_cap = seg_cap or min(B * F, cap)
    idx = [np.stack([rng.integers(0, r, size=B) for r in rows], 1).astype(np.int64) for _ in range(world)]
    grads = [rng.standard_normal((B * F, K)).astype(np.float32) for _ in range(world)]
    off = np.asarray(plans[0].field_offset)
    # requester side: distinct (owner, local) keys, owner-major; header + ids into the owners' buffers
    hdr = np.zeros((world, world, 2), np.int64)                     # hdr[o][q] = (count, base_u)
    ids = np.full((world, world, seg_cap), -1, np.int64)            # ids[o][q][i]
    inv, uk = [], []
    for q in range(world):
        grow = (idx[q] + off[None, :]).reshape(-1)
        key = (grow % world) * cap + grow // world
        ukeys, iv = np.unique(key, return_inverse=True)
        inv.append(iv)
        uk.append(ukeys)
        owner_off = np.searchsorted(ukeys, np.arange(world + 1) * cap)
        for o in range(world):
            c = owner_off[o + 1] - owner_off[o]
            assert c <= seg_cap
            hdr[o, q] = (c, owner_off[o])
            ids[o, q, :c] = ukeys[owner_off[o]:owner_off[o + 1]] % cap
    # owner side: slot map, gather + send
    slot = np.zeros((world, cap, world), np.int64)
    rowsbuf = [np.zeros((len(uk[q]), K), np.float32) for q in range(world)]
    for o in range(world):
        local = table[o::world]
        for q in range(world):
            c, base = hdr[o, q]
            r = ids[o, q, :c]
            assert (slot[o, r, q] == 0).all()                         # a requester sends a row at most once
            slot[o, r, q] = np.arange(c) + 1
            rowsbuf[q][base:base + c] = local[r]
    seen = [rowsbuf[q][inv[q]].reshape(B, F, K) for q in range(world)]
    # backward: per-distinct-row sums into the owner's segment q; the lowest asking rank merges in rank order
    gbuf = np.zeros((world, world, seg_cap, K), np.float32)
    for q in range(world):
        gu = np.zeros((len(uk[q]), K), np.float32)
        np.add.at(gu, inv[q], grads[q])
        for o in range(world):
            c, base = hdr[o, q]
            gbuf[o, q, :c] = gu[base:base + c]
    total = np.zeros((cap * world, K), np.float64)
    merges = 0
    for o in range(world):
        for q in range(world):
            for i in range(hdr[o, q, 0]):
                r = ids[o, q, i]
                if (slot[o, r, :q] != 0).any():
                    continue                                           # an earlier rank merges this row
                g = gbuf[o, q, i].astype(np.float64)
                for p in range(q + 1, world):
                    if slot[o, r, p]:
                        g = g + gbuf[o, p, slot[o, r, p] - 1]
                total[r * world + o] = g
                merges += 1
    return table, idx, grads, off, seen, total, merges


@pytest.mark.parametrize("world", [1, 2, 3, 8])
def test_peer_protocol_restated(world):
    """Host-checkable statement of the N > 1 device-driven exchange (the kernels themselves need GPUs:
    tests/test_gpu_sharded.py).  Rows seen through the exchange equal a direct lookup; merged sums equal a global
    scatter-add; every touched row is merged exactly once.  Includes the case ceil(n_rows / G) < distinct rows of
    one requester (world 8), which round 1's buffer sizing got wrong."""
    rows, K, B = [40, 1, 90, 300, 3, 1], 4, 64
    table, idx, grads, off, seen, total, merges = _peer_protocol(world, rows, K, B, seed=11)
    F = len(rows)
    ref = np.zeros_like(total)
    touched = set()
    for q in range(world):
        want, _ = O.embedding_lookup(table, off, idx[q], None)
        assert np.array_equal(seen[q], want)
        grow = (idx[q] + off[None, :]).reshape(-1)
        np.add.at(ref, grow, grads[q].astype(np.float64))
        touched |= set(grow.tolist())
    assert merges == len(touched)
    assert np.allclose(total[:sum(rows)], ref[:sum(rows)], atol=1e-5)
    if world == 8:
        from dir_b200.sharded import ShardPlan
        cap = ShardPlan(rows, world, 0).cap
        assert max(len(set((idx[q] + off[None, :]).reshape(-1).tolist())) for q in range(world)) > cap
