"""Row-sharded embedding + FM layer: one process per GPU, tables sharded by row over the ranks of a
process group, NCCL all-to-all for the lookup and the gradient exchange, the FM / first-order /
cross interaction data-parallel on the requesting rank.

The reference's only hook for this is the partitioner around its embedding variables
(`input_layer_partitioner`, models/DeepFM/deepFM.py:163-175): with a TF parameter-server cluster
the variables are split by row and ids / IndexedSlices travel over gRPC.  Here the exchange is two
all-to-alls per direction (counts, then payload) over NVLink, and only DISTINCT rows travel:
the requester sorts its lookups by (owner, local row), numbers the distinct ones, and after the
owner answered runs the ordinary forward kernel on the received buffer; in the backward it sums
the gradients of each distinct row before shipping them, so a one-row "dense" field or a Zipf-hot
row costs one row of traffic per rank and step instead of one per sample.

`ShardPlan` (pure index arithmetic) and `exchange` / `exchange_counts` (the collectives) carry no
CUDA-only code: tests run them under gloo with world_size 2 on CPU.
"""
import math
import os
import sys
from typing import Optional, Sequence

import torch
import torch.distributed as dist

from . import _lib
from ._lib import check, ptr
from .layers import (_K_OK, _OPTIMIZERS, _Workspace, _need_cuda, _stream, linear_opt_struct,
                     resolve_linear_optimizer)


class ShardPlan:
    """owner = global row mod G, local row = global row div G (modulo, so hot low rows spread)."""

    def __init__(self, rows_per_field: Sequence[int], world_size: int, rank: int):
        if world_size <= 0 or not 0 <= rank < world_size:
            raise ValueError("need 0 <= rank < world_size")
        self.rows_per_field = [int(r) for r in rows_per_field]
        self.world_size, self.rank = world_size, rank
        self.n_rows = sum(self.rows_per_field)
        off = [0]
        for r in self.rows_per_field[:-1]:
            off.append(off[-1] + r)
        self.field_offset = off
        self.cap = (self.n_rows + world_size - 1) // world_size          # rows every rank allocates
        self.n_local = (self.n_rows - rank + world_size - 1) // world_size if self.n_rows > rank else 0
        if self.cap * world_size >= 2 ** 32 - 1:
            raise ValueError("ceil(n_rows / world_size) * world_size must be < 2^32-1")

    def owner(self, global_rows):
        return global_rows % self.world_size

    def local_row(self, global_rows):
        return global_rows // self.world_size

    def global_row(self, local_rows, rank=None):
        return local_rows * self.world_size + (self.rank if rank is None else rank)

    def shard_of(self, full):
        """This rank's rows of a full [n_rows, ...] array (numpy or torch)."""
        return full[self.rank::self.world_size]


def exchange_counts(send_counts: torch.Tensor, group=None) -> torch.Tensor:
    """all-to-all of one int64 per peer: recv[g] = what rank g will send to this rank."""
    if not dist.is_initialized():                 # a single process owns every row
        return send_counts.clone()
    recv = torch.empty_like(send_counts)
    dist.all_to_all_single(recv, send_counts.contiguous(), group=group)
    return recv


def exchange(payload: torch.Tensor, send_splits, recv_splits, group=None) -> torch.Tensor:
    """Variable-size all-to-all along dim 0: rows [sum(send_splits), ...] grouped by destination ->
    rows [sum(recv_splits), ...] grouped by source."""
    if not dist.is_initialized():
        return payload.contiguous()
    out = payload.new_empty((int(sum(recv_splits)),) + tuple(payload.shape[1:]))
    dist.all_to_all_single(out, payload.contiguous(), output_split_sizes=list(recv_splits),
                           input_split_sizes=list(send_splits), group=group)
    return out


def exchange_into(out: torch.Tensor, payload: torch.Tensor, send_splits, recv_splits, group=None) -> torch.Tensor:
    """`exchange` into the head of a caller-owned buffer (fixed address: CUDA-graph replays read it)."""
    n = int(sum(recv_splits))
    if n > out.shape[0]:
        raise ValueError("exchange_into: %d rows received, buffer holds %d" % (n, out.shape[0]))
    view = out[:n]
    if not dist.is_initialized():
        view.copy_(payload)
        return view
    dist.all_to_all_single(view, payload.contiguous(), output_split_sizes=list(recv_splits),
                           input_split_sizes=list(send_splits), group=group)
    return view


def peer_offsets(M: torch.Tensor, me: int):
    """Where the segments of the peer-memory exchange start, from everybody's counts.

    M[q, o] = distinct rows requester q wants from owner o (int64 [G, G], the same on every rank).
      recv_off[G+1]   owner `me`: its answer list is grouped by requester; group q starts at recv_off[q]
      fwd_dst_off[G]  owner `me` -> requester q: q's row buffer is grouped by owner, `me`'s group starts here
      bwd_dst_off[G]  requester `me` -> owner o: o's gradient buffer is grouped by requester (= the order of
                      its answer list), `me`'s group starts here
    """
    recv_counts = M[:, me]
    recv_off = torch.cat([torch.zeros(1, dtype=M.dtype, device=M.device), torch.cumsum(recv_counts, 0)])
    fwd_dst_off = (torch.cumsum(M, 1) - M)[:, me].contiguous()
    bwd_dst_off = (torch.cumsum(M, 0) - M)[me, :].contiguous()
    return recv_off, fwd_dst_off, bwd_dst_off


class StageTrace:
    """Optional per-stage CUDA-event timing of the sharded step (DIR_B200_TRACE=1): mark(name) closes
    the stage that just ran; report() averages over the steps seen since the last report."""

    def __init__(self):
        self.on = os.environ.get("DIR_B200_TRACE", "0") == "1"
        self.marks, self.tot, self.steps = [], {}, 0

    def mark(self, name):
        if self.on:
            ev = torch.cuda.Event(enable_timing=True)
            ev.record()
            self.marks.append((name, ev))

    def close_step(self):
        if not self.on or len(self.marks) < 2:
            return
        torch.cuda.synchronize()
        for (_, a), (name, b) in zip(self.marks[:-1], self.marks[1:]):
            self.tot[name] = self.tot.get(name, 0.0) + a.elapsed_time(b) * 1e3
        self.marks, self.steps = [], self.steps + 1

    def report(self):
        out = {k: v / max(1, self.steps) for k, v in self.tot.items()}
        self.tot, self.steps = {}, 0
        return out


class PeerBuffers:
    """Double-buffered symmetric memory for the payload exchange: every rank allocates the same buffers and
    maps its peers' copies (torch.distributed._symmetric_memory, CUDA IPC over NVLink).  `ptrs[p]` is a
    device int64[G] of peer-mapped addresses of buffer p; `barrier(p)` is a device-side cross-rank barrier
    on the current stream (no host involvement)."""

    def __init__(self, group, rows_cap, stride, device):
        import torch.distributed._symmetric_memory as symm_mem
        self.bufs, self.handles, self.ptrs = [], [], []
        name = group if group is not None else dist.group.WORLD
        for _ in range(2):
            t = symm_mem.empty((rows_cap, stride), dtype=torch.float32, device=device)
            h = symm_mem.rendezvous(t, name)
            self.bufs.append(t)
            self.handles.append(h)
            self.ptrs.append(torch.tensor(list(h.buffer_ptrs), dtype=torch.int64, device=device))
        self.rows_cap = rows_cap

    def barrier(self, p, channel=0):
        self.handles[p].barrier(channel=channel)


class ShardedLookups:
    """What `ShardedEmbeddingFM.presort` leaves behind for one batch -- everything that depends on the
    ids only: this rank's lookups sorted by (owner, local row), the distinct rows numbered, the ids
    already exchanged, and the owner's half (the ids it received, sorted for the gradient merge)."""

    def __init__(self):
        self.ws, self.ws2, self.ws3 = _Workspace(), _Workspace(), _Workspace()
        self.keys = self.uidx = self.ulocal = self.inv = self.owner_off = self.recv_ids = None
        self.send_splits = self.recv_splits = None
        self.recv_off = self.fwd_dst_off = self.bwd_dst_off = None     # device offsets of the peer exchange
        self.U = self.R = 0
        self.event = None
        self.src = None
        self.parity = None          # which half of the double-buffered peer memory (None: the layer alternates)
        self.recv_handle = self.recv_ptrs = None   # DIR_B200_IDS=peer: symmetric-memory handle / peer addresses of recv_buf
        self.inv_c = None                          # DIR_B200_SHARD_ONEROW=1: compact [B, n_sel] row indices
        self.recv_buf = None        # static mode: fixed-address landing buffer of the ids this rank answers

    @staticmethod
    def key_of(feature_index, feature_value):
        return (feature_index.data_ptr(), None if feature_value is None else feature_value.data_ptr(),
                tuple(feature_index.shape))


class _ShardedFunction(torch.autograd.Function):
    @staticmethod
    def forward(ctx, anchor, bias, layer, idx, val, train, h):
        B, F = idx.shape
        K = layer.embedding_size
        dev = idx.device
        L = _lib.lib()
        st = _stream()
        pad = layer.pad_stride
        tr = layer.trace
        tr.mark("start")
        if h.event is not None and not layer.capturing:
            torch.cuda.current_stream().wait_event(h.event)
        U, R = h.U, h.R
        static = layer.static
        # 4b: the owner answers the ids it received; rows back over NVLink
        G = layer.plan.world_size
        if h.parity is not None:
            parity = h.parity
        else:
            parity = layer._step_parity
            layer._step_parity ^= 1
        if layer.peer is not None:
            if B > layer.max_batch:
                raise ValueError("batch %d exceeds max_batch=%d the peer buffers were sized for" % (B, layer.max_batch))
            # one kernel gathers the rows and writes them straight into the requesters' buffers
            pb = layer.peer["rows"]
            # static: launch for the capacity, the kernel stops at recv_off[G] (read on the device)
            check(L.dir_rows_gather_to(ptr(layer.table), layer.row_stride,
                                       ptr(layer.w1) if layer.first_order else None, layer.lin_stride,
                                       ptr(h.recv_buf if static else h.recv_ids), layer.recv_cap if static else R,
                                       K, G, ptr(h.recv_off), ptr(pb.ptrs[parity]),
                                       ptr(h.fwd_dst_off), pad, st), "dir_rows_gather_to")
            tr.mark("fwd.gather+send")
            if layer.onerow_rep:
                # EXPERIMENT: the replicated one-row fields' rows sit behind the exchanged ones (inv points there)
                tail = pb.bufs[parity][layer.rows_cap:]
                check(L.dir_rows_gather(ptr(layer.dense_table), layer.row_stride,
                                        ptr(layer.dense_lin) if layer.first_order else None, 1,
                                        ptr(layer.dense_ids), layer.n_onerow, K, ptr(tail), pad, st), "dir_rows_gather")
            pb.barrier(parity)
            tr.mark("fwd.barrier")
            ubuf = pb.bufs[parity]
            use_peer = True
        else:
            answer = torch.empty((R, pad), dtype=torch.float32, device=dev)
            check(L.dir_rows_gather(ptr(layer.table), layer.row_stride,
                                    ptr(layer.w1) if layer.first_order else None, layer.lin_stride,
                                    ptr(h.recv_ids), R, K, ptr(answer), pad, st), "dir_rows_gather")
            tr.mark("fwd.gather")
            ubuf = exchange(answer, h.recv_splits, h.send_splits, layer.group)          # [U, K+4]
            tr.mark("fwd.a2a_rows")
            if U == 0:
                ubuf = torch.zeros((1, pad), dtype=torch.float32, device=dev)
            use_peer = False
        # 5: the ordinary forward on the received rows
        emb = torch.empty((B, F * K), dtype=torch.float32, device=dev) if layer.emit_embeddings else None
        fm = torch.empty((B, 1), dtype=torch.float32, device=dev)
        first = torch.empty((B, 1), dtype=torch.float32, device=dev)
        S = torch.empty((B, K), dtype=torch.float32, device=dev) if train else None
        lin = ubuf[:, K] if layer.first_order else None
        check(L.dir_embed_fm_fwd(ptr(ubuf), pad, ptr(lin), pad, ptr(bias) if layer.first_order else None,
                                 ptr(h.inv), ptr(val), ptr(layer.zero_offset), None,
                                 ubuf.shape[0] if static else max(U, 1), B, F, K,
                                 ptr(emb), ptr(S), ptr(first), ptr(fm), None, None, st), "dir_embed_fm_fwd")
        tr.mark("fwd.fm")
        if not layer.first_order:
            first.zero_()
        layer.last_exchange = {"unique_sent": U, "unique_received": R, "lookups": B * F}
        ctx.layer, ctx.train, ctx.shape, ctx.h = layer, train, (B, F, K), h
        ctx.use_peer, ctx.parity = use_peer, parity
        ctx.idx = idx if layer.onerow_rep else None
        ctx.set_materialize_grads(False)
        if train:
            ctx.save_for_backward(val, S, ubuf)
        if emb is None:
            emb = torch.empty((B, 0), dtype=torch.float32, device=dev)
            ctx.mark_non_differentiable(emb)
        return first, fm, emb

    @staticmethod
    def backward(ctx, g_first, g_fm, u):
        if not ctx.train:
            raise RuntimeError("ShardedEmbeddingFM.backward: forward ran without gradient tracking")
        layer, h = ctx.layer, ctx.h
        val, S, ubuf = ctx.saved_tensors
        B, F, K = ctx.shape
        U, R = h.U, h.R
        dev = S.device
        L = _lib.lib()
        st = _stream()
        pad = layer.pad_stride
        n = B * F
        g_first = (torch.zeros(B, dtype=torch.float32, device=dev) if g_first is None
                   else g_first.reshape(B).contiguous().float())
        g_fm = (torch.zeros(B, dtype=torch.float32, device=dev) if g_fm is None
                else g_fm.reshape(B).contiguous().float())
        if u is not None:
            u = u.contiguous().float()
        tr = layer.trace
        tr.mark("between")
        with torch.no_grad():
            # 6: per-distinct-row sums on the requester (the sorted list is in the handle's workspace)
            ws = h.ws.get(L.dir_embed_bwd_workspace_bytes(n, K), dev)
            n_keys = layer.plan.cap * layer.plan.world_size
            gfp = ptr(g_first) if layer.first_order else None
            if ctx.use_peer and layer.onerow_rep:
                # EXPERIMENT: the sorted list covers the multi-row fields only; the one-row fields' gradients over this
                # rank's samples go to every rank's buffer (behind the exchanged rows) from a column-sum kernel
                pb = layer.peer["grads"]
                G_ = layer.plan.world_size
                ws = h.ws.get(L.dir_embed_bwd_workspace_bytes(max(B * layer.n_sel, 1), K), dev)
                check(L.dir_embed_bwd_reduce_emit_fields_to(
                    ptr(ubuf), pad, ptr(val), gfp, ptr(g_fm), ptr(S), ptr(u), ptr(h.uidx), B, F, K, n_keys,
                    ptr(layer.sparse_fields), layer.n_sel, G_, ptr(h.owner_off), ptr(pb.ptrs[ctx.parity]),
                    ptr(h.bwd_dst_off), pad, ptr(ws), ws.numel(), st), "dir_embed_bwd_reduce_emit_fields_to")
                ows = layer._onerow_ws.get(L.dir_onerow_workspace_bytes(K), dev)
                check(L.dir_embed_bwd_onerow_emit_to(
                    ptr(layer.dense_table), layer.row_stride, ptr(ctx.idx), ptr(val), ptr(layer.dense_field_offset), gfp,
                    ptr(g_fm), ptr(S), ptr(u), ptr(layer.onerow_fields), layer.n_onerow, B, F, K, G_, layer.plan.rank,
                    ptr(pb.ptrs[ctx.parity]), layer.recv_cap, pad, ptr(ows), ows.numel(), st),
                    "dir_embed_bwd_onerow_emit_to")
                tr.mark("bwd.emit+push")
                pb.barrier(ctx.parity)
                tr.mark("bwd.barrier")
                grecv = pb.bufs[ctx.parity]
            elif ctx.use_peer and layer.fused_push:
                # 6+7 in one kernel: each distinct row's sums go straight into its owner's buffer over NVLink
                pb = layer.peer["grads"]
                check(L.dir_embed_bwd_reduce_emit_to(
                    ptr(ubuf), pad, ptr(val), gfp, ptr(g_fm), ptr(S), ptr(u), ptr(h.uidx), B, F, K, n_keys,
                    layer.plan.world_size, ptr(h.owner_off), ptr(pb.ptrs[ctx.parity]), ptr(h.bwd_dst_off), pad,
                    ptr(ws), ws.numel(), st), "dir_embed_bwd_reduce_emit_to")
                tr.mark("bwd.emit+push")
                pb.barrier(ctx.parity)
                tr.mark("bwd.barrier")
                grecv = pb.bufs[ctx.parity]
            else:
                gu = torch.empty((max(U, 1), pad), dtype=torch.float32, device=dev)
                check(L.dir_embed_bwd_reduce_emit(ptr(ubuf), pad, ptr(val), gfp, ptr(g_fm), ptr(S), ptr(u),
                                                  ptr(h.uidx), B, F, K, n_keys, ptr(gu), pad,
                                                  ptr(ws), ws.numel(), st), "dir_embed_bwd_reduce_emit")
                tr.mark("bwd.emit")
            # 7: sums to their owners; the owner merges the ranks' contributions and updates
            if ctx.use_peer and (layer.fused_push or layer.onerow_rep):
                pass
            elif ctx.use_peer:
                pb = layer.peer["grads"]
                check(L.dir_rows_push(ptr(gu), U, pad, layer.plan.world_size, ptr(h.owner_off),
                                      ptr(pb.ptrs[ctx.parity]), ptr(h.bwd_dst_off), st), "dir_rows_push")
                tr.mark("bwd.push")
                pb.barrier(ctx.parity)
                tr.mark("bwd.barrier")
                grecv = pb.bufs[ctx.parity]
            else:
                grecv = exchange(gu[:U], h.send_splits, h.recv_splits, layer.group)      # [R, K+4]
                tr.mark("bwd.a2a_grads")
            if R > 0 or layer.static:
                Rn = layer.recv_cap if layer.static else R        # static: capacity; the count stays on the device
                ws3 = h.ws3.get(L.dir_embed_bwd_workspace_bytes(Rn, K), dev)     # sorted by presort
                adagrad = layer.optimizer == "adagrad"
                n_dev = h.recv_off.data_ptr() + 8 * layer.plan.world_size if layer.static else None
                check(L.dir_rows_reduce_update(
                    ptr(layer.table), ptr(layer.accum) if adagrad else None, layer.row_stride,
                    ptr(layer.w1) if layer.first_order else None,
                    ptr(layer.w1_accum) if layer.first_order else None, layer.lin_stride,
                    ptr(grecv), pad, Rn, K, layer.plan.cap, _OPTIMIZERS[layer.optimizer], layer.lr,
                    linear_opt_struct(layer), n_dev,
                    ptr(ws3), ws3.numel(), ptr(layer.last_n_unique), st), "dir_rows_reduce_update")
                tr.mark("bwd.owner_update")
            else:
                layer.last_n_unique.zero_()
            if ctx.use_peer and layer.onerow_rep:
                adagrad = layer.optimizer == "adagrad"
                lo = layer.dense_linear_opt()
                check(L.dir_dense_rows_apply(
                    ptr(layer.dense_table), ptr(layer.dense_accum) if adagrad else None, layer.row_stride,
                    ptr(layer.dense_lin) if layer.first_order else None,
                    ptr(layer.dense_lin_acc) if layer.first_order else None, ptr(grecv), pad, layer.recv_cap,
                    layer.n_onerow, K, layer.plan.world_size, _OPTIMIZERS[layer.optimizer], layer.lr, lo,
                    ptr(layer.table), ptr(layer.accum) if adagrad else None, layer.row_stride,
                    ptr(layer.w1) if layer.first_order else None,
                    ptr(layer.w1_accum) if layer.first_order else None,
                    ptr(layer.lin_z) if (layer.first_order and layer.lin_z is not None) else None, layer.lin_stride,
                    ptr(layer.dense_shard_row), ptr(layer.last_n_unique) if layer.plan.rank == 0 else None, st),
                    "dir_dense_rows_apply")
                tr.mark("bwd.dense_apply")
            tr.close_step()
        g_bias = g_first.sum().reshape(1) if layer.first_order else None
        return None, g_bias, None, None, None, None, None


class ShardedEmbeddingFM(torch.nn.Module):
    """`EmbeddingFM` with its tables sharded by row over a process group (one process per GPU).

    Same call as EmbeddingFM: forward(feature_index[B_local, F] int64, feature_value | None) ->
    (first_order, fm_second_order, embeddings) for THIS rank's samples; `.backward()` updates the
    rows this rank owns with the gradients of every rank's samples.  `bias` is an ordinary
    replicated parameter: all-reduce its gradient like any dense parameter.
    """

    def __init__(self, field_size: int, embedding_size: int, rows_per_field: Sequence[int],
                 optimizer: str = "adagrad", lr: float = 0.01, initial_accumulator_value: float = 0.1,
                 first_order: bool = True, emit_embeddings: bool = True, check_bounds: bool = False,
                 process_group=None, max_batch: int = 65536, linear_optimizer: Optional[str] = None,
                 linear_lr: Optional[float] = None, l1_regularization_strength: float = 0.0,
                 l2_regularization_strength: float = 0.0, device="cuda"):
        super().__init__()
        if field_size <= 0:
            raise ValueError("empty columns.")                      # deepFM.py:104-105
        if embedding_size not in _K_OK:
            raise ValueError("embedding_size must be one of %r" % (_K_OK,))
        optimizer = optimizer.lower()
        if optimizer not in _OPTIMIZERS:
            raise ValueError("optimizer must be 'adagrad' or 'sgd'")
        self.l1, self.l2 = float(l1_regularization_strength), float(l2_regularization_strength)
        self.linear_optimizer, self.linear_lr, lin_needs_acc, lin_needs_z = resolve_linear_optimizer(
            optimizer, lr, linear_optimizer, linear_lr, self.l1, self.l2)
        rows = [int(r) for r in rows_per_field]
        if len(rows) != field_size:
            raise ValueError("rows_per_field must have field_size entries")
        self.group = process_group
        world = dist.get_world_size(process_group) if dist.is_initialized() else 1
        rank = dist.get_rank(process_group) if dist.is_initialized() else 0
        self.plan = ShardPlan(rows, world, rank)
        self.field_size, self.embedding_size = field_size, embedding_size
        self.optimizer, self.lr = optimizer, float(lr)
        self.first_order, self.emit_embeddings, self.check_bounds = first_order, emit_embeddings, check_bounds
        K = embedding_size
        adagrad = optimizer == "adagrad"
        self.row_stride = 2 * K if adagrad else K
        self.lin_stride = 1
        self.pad_stride = K + 4                                     # (row[K], first-order weight, 3 pad)
        self.n_rows = self.plan.cap                                 # local rows allocated
        dev = torch.device(device)
        cap = max(self.plan.cap, 1)
        self.register_buffer("field_offset", torch.tensor(self.plan.field_offset, dtype=torch.int64, device=dev))
        self.register_buffer("field_rows", torch.tensor(rows, dtype=torch.int64, device=dev))
        self.register_buffer("zero_offset", torch.zeros(field_size, dtype=torch.int64, device=dev))
        self.register_buffer("rows", torch.empty((cap, self.row_stride), dtype=torch.float32, device=dev))
        self.register_buffer("lin_rows", torch.zeros((cap, 1), dtype=torch.float32, device=dev))
        self.register_buffer("lin_acc", torch.zeros((cap, 1), dtype=torch.float32, device=dev) if lin_needs_acc else None)
        self.register_buffer("lin_z", torch.zeros((cap, 1), dtype=torch.float32, device=dev) if lin_needs_z else None)
        self.register_buffer("oob_flag", torch.zeros(1, dtype=torch.int32, device=dev))
        self.bias = torch.nn.Parameter(torch.zeros(1, dtype=torch.float32, device=dev))
        self._anchor = torch.nn.Parameter(torch.zeros(1, dtype=torch.float32, device=dev))
        self.last_n_unique = torch.zeros(1, dtype=torch.int64, device=dev)
        self.last_exchange = {}
        self._side = None
        self._inline = ShardedLookups()
        # the id exchange of the NEXT batch runs concurrently with this batch's row / gradient exchange:
        # give it a communicator of its own so the two never queue behind each other
        self.side_group = (dist.new_group(ranks=dist.get_process_group_ranks(process_group or dist.group.WORLD))
                           if dist.is_initialized() and dist.get_backend(process_group) == "nccl" else process_group)
        self.trace, self.trace_pre = StageTrace(), StageTrace()
        # Payload exchange: NVLink peer memory (symmetric buffers + a device-side barrier) when there is
        # more than one rank and torch's symmetric memory is usable, else NCCL all-to-all.
        self.peer, self._step_parity = None, 0
        self.static = self.capturing = self.ids_peer = self.onerow_rep = False
        # EXPERIMENT (DIR_B200_SHARD_ONEROW=1, needs the peer exchange): one-row fields as replicated parameters
        onerow = [f for f, r in enumerate(rows) if r == 1][:64]
        want_rep = os.environ.get("DIR_B200_SHARD_ONEROW", "0") == "1" and 0 < len(onerow) < field_size
        self.n_onerow = len(onerow) if want_rep else 0
        sparse = [f for f in range(field_size) if not (want_rep and f in set(onerow))]
        self.n_sel = len(sparse)
        self.recv_cap = 0
        self.max_batch = int(max_batch)
        want_peer = os.environ.get("DIR_B200_EXCHANGE", "peer") == "peer"
        self.fused_push = os.environ.get("DIR_B200_FUSED_PUSH", "1") == "1"   # emit + NVLink push in one kernel
        if want_peer and dist.is_initialized() and world > 1 and dev.type == "cuda" \
                and dist.get_backend(process_group) == "nccl":
            try:
                cap = min(self.max_batch * field_size, max(self.plan.cap, 1))       # distinct rows a rank can want
                n1 = self.n_onerow                                                  # replicated rows ride behind them
                self.peer = {"rows": PeerBuffers(process_group, cap + n1, self.pad_stride, dev),         # <- owners
                             "grads": PeerBuffers(process_group, cap * world + n1 * world, self.pad_stride, dev)}  # <- requesters
                self.rows_cap = cap
                # Static mode: every main-stream launch is sized for these capacities and reads the real counts
                # on the device, every buffer it touches has a fixed address -- so forward + backward of a step
                # can be captured in a CUDA graph (the id-only presort stays eager on the side stream).
                self.recv_cap = cap * world
                self.static = os.environ.get("DIR_B200_STATIC", "1") == "1"
                # experiment (not yet run on a GPU): ids through peer memory instead of the NCCL all-to-all
                self.ids_peer = self.static and os.environ.get("DIR_B200_IDS", "nccl") == "peer"
                self.onerow_rep = self.static and want_rep
            except Exception as e:                                              # no IPC / fabric support
                if rank == 0:
                    print("ShardedEmbeddingFM: symmetric memory unavailable (%s); using NCCL all-to-all" % e,
                          file=sys.stderr)
                self.peer = None
        with torch.no_grad():
            sd = 1.0 / math.sqrt(K)
            torch.nn.init.trunc_normal_(self.table, 0.0, sd, -2.0 * sd, 2.0 * sd)
            if adagrad:
                self.accum.fill_(initial_accumulator_value)
            if self.lin_acc is not None:
                self.lin_acc.fill_(initial_accumulator_value)
        if not self.onerow_rep:
            self.n_onerow, self.n_sel = 0, field_size
        else:
            self._init_onerow_replicas(onerow, sparse, dev, initial_accumulator_value)

    @property
    def table(self):
        return self.rows[:, :self.embedding_size]

    @property
    def accum(self):
        return self.rows[:, self.embedding_size:] if self.optimizer == "adagrad" else None

    @property
    def w1(self):
        return self.lin_rows[:, 0]

    @property
    def w1_accum(self):
        return self.lin_acc[:, 0] if self.lin_acc is not None else None

    # -- EXPERIMENT: replicated one-row fields ------------------------------------------------------------------
    @torch.no_grad()
    def _init_onerow_replicas(self, onerow, sparse, dev, acc0):
        K, G, rank = self.embedding_size, self.plan.world_size, self.plan.rank
        n1 = len(onerow)
        self.register_buffer("onerow_fields", torch.tensor(onerow, dtype=torch.int32, device=dev))
        self.register_buffer("sparse_fields", torch.tensor(sparse, dtype=torch.int32, device=dev))
        self.register_buffer("dense_ids", torch.arange(n1, dtype=torch.int32, device=dev))
        dfo = [0] * self.field_size
        for j, f in enumerate(onerow):
            dfo[f] = j
        self.register_buffer("dense_field_offset", torch.tensor(dfo, dtype=torch.int64, device=dev))
        grow = [self.plan.field_offset[f] for f in onerow]                     # the fields' rows in the full table
        self.dense_global_rows = grow
        self.register_buffer("dense_shard_row", torch.tensor([g // G if g % G == rank else -1 for g in grow],
                                                             dtype=torch.int64, device=dev))
        self.register_buffer("dense_rows", torch.zeros((n1, self.row_stride), dtype=torch.float32, device=dev))
        self.register_buffer("dense_lin", torch.zeros(n1, dtype=torch.float32, device=dev))
        self.register_buffer("dense_lin_acc", torch.full((n1,), float(acc0), dtype=torch.float32, device=dev)
                             if self.lin_acc is not None else None)
        self.register_buffer("dense_lin_z", torch.zeros(n1, dtype=torch.float32, device=dev)
                             if self.lin_z is not None else None)
        self._onerow_ws = _Workspace()
        self._sync_onerow_replicas()

    @torch.no_grad()
    def _sync_onerow_replicas(self):
        """Replicas <- the owners' rows of the sharded table (all-reduce of rows that are zero off their owner)."""
        K = self.embedding_size
        buf = torch.zeros((self.n_onerow, self.row_stride + 3), dtype=torch.float32, device=self.rows.device)
        mine = self.dense_shard_row >= 0
        sr = self.dense_shard_row[mine]
        buf[mine, :self.row_stride] = self.rows[sr]
        buf[mine, self.row_stride] = self.lin_rows[sr, 0]
        if self.lin_acc is not None:
            buf[mine, self.row_stride + 1] = self.lin_acc[sr, 0]
        if self.lin_z is not None:
            buf[mine, self.row_stride + 2] = self.lin_z[sr, 0]
        if dist.is_initialized() and self.plan.world_size > 1:
            dist.all_reduce(buf, group=self.group)
        self.dense_rows.copy_(buf[:, :self.row_stride])
        self.dense_lin.copy_(buf[:, self.row_stride])
        if self.dense_lin_acc is not None:
            self.dense_lin_acc.copy_(buf[:, self.row_stride + 1])
        if self.dense_lin_z is not None:
            self.dense_lin_z.copy_(buf[:, self.row_stride + 2])

    @property
    def dense_table(self):
        return self.dense_rows[:, :self.embedding_size]

    @property
    def dense_accum(self):
        return self.dense_rows[:, self.embedding_size:] if self.optimizer == "adagrad" else None

    def dense_linear_opt(self):
        """dir_linear_opt for the replicas (same rule as the sharded weights, the replicas' own Ftrl slot)."""
        if self.linear_optimizer is None:
            return None
        from .layers import _LINEAR_OPTIMIZERS
        z = self.dense_lin_z.data_ptr() if self.dense_lin_z is not None else None
        return _lib.ctypes.byref(_lib.LinearOpt(_LINEAR_OPTIMIZERS[self.linear_optimizer], self.linear_lr,
                                                self.l1, self.l2, z))

    @torch.no_grad()
    def load_tables(self, table=None, w1=None):
        """Takes the FULL [n_rows, K] table / [n_rows] first-order weights and keeps this rank's rows."""
        for dst, src in ((self.table, table), (self.w1, w1)):
            if src is not None:
                mine = torch.as_tensor(self.plan.shard_of(src), dtype=torch.float32)
                dst[:mine.shape[0]].copy_(mine.to(dst.device))
        if self.onerow_rep:
            self._sync_onerow_replicas()

    def _prepare(self, feature_index, feature_value):
        if feature_index.dim() != 2 or feature_index.shape[1] != self.field_size:
            raise ValueError("feature_index must be [B, field_size=%d]" % self.field_size)
        if feature_index.dtype != torch.int64:
            raise ValueError("feature_index must be int64")
        _need_cuda(feature_index, "feature_index")
        idx = feature_index.contiguous()
        val = None
        if feature_value is not None:
            if feature_value.shape != feature_index.shape:
                raise ValueError("feature_value must have feature_index's shape")
            _need_cuda(feature_value, "feature_value")
            val = feature_value.contiguous().float()
        return idx, val

    def side_stream(self, device):
        if self._side is None:
            self._side = torch.cuda.Stream(device=device, priority=-1)
        return self._side

    @torch.no_grad()
    def presort(self, feature_index, feature_value=None, handle=None, after=None, fork=True):
        """Everything of a step that depends on the ids only, on the side stream and on a process group of
        its own: composite keys, sort, distinct-row numbering, the counts / ids all-to-all, and the owner's
        sort of the ids it received.  Issue it for batch i+1 right after enqueueing step i: it then runs
        underneath step i, and the one host read of the split sizes waits on the side stream only.
        `fork=False`: do not order the side stream after the work already queued on the current stream (the
        ids are already on the device, or `after` marks their arrival) -- this is what lets it overlap.
        """
        idx, val = self._prepare(feature_index, feature_value)
        B, F = idx.shape
        K, G = self.embedding_size, self.plan.world_size
        dev = idx.device
        L = _lib.lib()
        n_full = B * F
        n = B * self.n_sel if self.onerow_rep else n_full      # EXPERIMENT: one-row fields stay out of the sorted list
        h = handle if handle is not None else ShardedLookups()
        main, side = torch.cuda.current_stream(), self.side_stream(dev)
        if fork:        # order after whatever the current stream has queued (it may be producing the ids)
            ev = torch.cuda.Event()
            ev.record(main)
            side.wait_event(ev)
        if after is not None:
            side.wait_event(after)
        tr = self.trace_pre
        with torch.cuda.stream(side):
            st = side.cuda_stream
            tr.mark("pre.start")
            if h.keys is None or h.keys.numel() != n or h.keys.device != dev:
                h.keys = torch.empty(n, dtype=torch.int32, device=dev)
                h.uidx = torch.empty(n, dtype=torch.int32, device=dev)
                h.ulocal = torch.empty(n, dtype=torch.int32, device=dev)
                h.inv = torch.empty((B, F), dtype=torch.int64, device=dev)
                if self.onerow_rep:
                    h.inv_c = torch.empty((B, self.n_sel), dtype=torch.int64, device=dev)
                h.owner_off = torch.empty(G + 1, dtype=torch.int64, device=dev)
                h.recv_off = torch.zeros(G + 1, dtype=torch.int64, device=dev)
                h.fwd_dst_off = torch.zeros(G, dtype=torch.int64, device=dev)
                h.bwd_dst_off = torch.zeros(G, dtype=torch.int64, device=dev)
                if self.static and self.ids_peer:
                    # the landing buffer of the ids is symmetric memory too: requesters store into it directly
                    # (collective: every rank reaches this line for the same handle, in the same order)
                    import torch.distributed._symmetric_memory as symm_mem
                    h.recv_buf = symm_mem.empty(self.recv_cap, dtype=torch.int32, device=dev)
                    h.recv_handle = symm_mem.rendezvous(h.recv_buf, self.group if self.group is not None
                                                        else dist.group.WORLD)
                    h.recv_ptrs = torch.tensor(list(h.recv_handle.buffer_ptrs), dtype=torch.int64, device=dev)
                elif self.static:
                    h.recv_buf = torch.empty(self.recv_cap, dtype=torch.int32, device=dev)
            check(L.dir_shard_keys(ptr(idx), ptr(val), ptr(self.field_offset), ptr(self.field_rows),
                                   self.plan.n_rows, B, F, G, ptr(self.sparse_fields) if self.onerow_rep else None,
                                   self.n_sel if self.onerow_rep else F, ptr(h.keys),
                                   ptr(self.oob_flag) if self.check_bounds else None, st), "dir_shard_keys")
            tr.mark("pre.keys")
            # sized for all B * F lookups whatever the list holds, so that nobody re-allocates it later
            ws = h.ws.get(L.dir_embed_bwd_workspace_bytes(max(n_full, 1), K), dev)
            check(L.dir_embed_bwd_sort(ptr(h.keys), n, self.plan.cap * G, ptr(ws), ws.numel(), st),
                  "dir_embed_bwd_sort")
            tr.mark("pre.sort")
            if n > 0:
                skeys, spos = _lib.c_void_p(), _lib.c_void_p()
                check(L.dir_embed_bwd_sorted(ptr(ws), n, _lib.ctypes.byref(skeys), _lib.ctypes.byref(spos)),
                      "dir_embed_bwd_sorted")
            else:
                skeys = spos = None
            ws2 = h.ws2.get(max(L.dir_shard_unique_workspace_bytes(n), 1), dev)
            check(L.dir_shard_unique(skeys, spos, n, self.plan.n_rows, G, ptr(h.uidx), ptr(h.ulocal),
                                     ptr(h.inv_c if self.onerow_rep else h.inv),
                                     ptr(h.owner_off), ptr(ws2), ws2.numel(), st), "dir_shard_unique")
            if self.onerow_rep:              # compact [B, n_sel] indices -> their columns of the [B, F] index
                h.inv.index_copy_(1, self.sparse_fields.long(), h.inv_c)
                # the replicated rows sit behind the exchanged ones in the row buffer; an id other than 0 is pruned
                one = self.onerow_fields.long()
                tail = self.rows_cap + torch.arange(self.n_onerow, dtype=torch.int64, device=dev)[None, :]
                h.inv.index_copy_(1, one, torch.where(idx.index_select(1, one) == 0, tail.expand(B, -1),
                                                      torch.full_like(tail, -1).expand(B, -1)))
            tr.mark("pre.unique")
            # counts: the one host read per step (NCCL needs the split sizes); it waits on the side stream only
            send_counts = h.owner_off[1:] - h.owner_off[:-1]
            if self.peer is not None:
                # everybody's counts: M[q, o] = distinct rows q wants from o.  From it, on the device, where
                # each rank's segment starts inside its peers' exchange buffers.
                M = torch.empty((G, G), dtype=torch.int64, device=dev)
                dist.all_gather_into_tensor(M, send_counts.contiguous(), group=self.side_group)
                me = self.plan.rank
                recv_counts = M[:, me].contiguous()
                recv_off, fwd_dst_off, bwd_dst_off = peer_offsets(M, me)
                # written in place: a captured step reads these at fixed addresses
                h.recv_off.copy_(recv_off)
                h.fwd_dst_off.copy_(fwd_dst_off)
                h.bwd_dst_off.copy_(bwd_dst_off)
            else:
                recv_counts = exchange_counts(send_counts, self.side_group)
            both = torch.stack([send_counts, recv_counts]).cpu()
            h.send_splits, h.recv_splits = both[0].tolist(), both[1].tolist()
            h.U, h.R = int(sum(h.send_splits)), int(sum(h.recv_splits))
            tr.mark("pre.counts+sync")
            if self.static:
                if h.R > self.recv_cap:
                    raise ValueError("%d rows requested from this rank, buffers hold %d" % (h.R, self.recv_cap))
            if self.static and self.ids_peer:
                # ids over NVLink peer memory: my segment for owner o starts where my gradient segment will
                # (bwd_dst_off[o]); a device-side barrier on this handle's own signal pad closes the exchange
                check(L.dir_ids_push(ptr(h.ulocal), n, G, ptr(h.owner_off), ptr(h.recv_ptrs), ptr(h.bwd_dst_off), st),
                      "dir_ids_push")
                h.recv_handle.barrier(channel=0)
                h.recv_ids = h.recv_buf[:h.R]
            elif self.static:
                h.recv_ids = exchange_into(h.recv_buf, h.ulocal[:h.U], h.send_splits, h.recv_splits, self.side_group)
            else:
                h.recv_ids = exchange(h.ulocal[:h.U], h.send_splits, h.recv_splits, self.side_group)
            tr.mark("pre.a2a_ids")
            if self.static:                   # laid out for the capacity: the consumer never learns R on the host
                ws3 = h.ws3.get(L.dir_embed_bwd_workspace_bytes(self.recv_cap, K), dev)
                check(L.dir_embed_bwd_sort_in(ptr(h.recv_ids), h.R, self.recv_cap, self.plan.cap, ptr(ws3),
                                              ws3.numel(), st), "dir_embed_bwd_sort_in")
            elif h.R > 0:                     # the owner's half: arrival order -> local-row order
                ws3 = h.ws3.get(L.dir_embed_bwd_workspace_bytes(h.R, K), dev)
                check(L.dir_embed_bwd_sort(ptr(h.recv_ids), h.R, self.plan.cap, ptr(ws3), ws3.numel(), st),
                      "dir_embed_bwd_sort")
            tr.mark("pre.owner_sort")
            h.event = torch.cuda.Event()
            h.event.record(side)
            tr.close_step()
        h.src = ShardedLookups.key_of(feature_index, feature_value)
        return h

    def forward(self, feature_index, feature_value=None, presorted=None):
        idx, val = self._prepare(feature_index, feature_value)
        train = self.training and torch.is_grad_enabled()
        if presorted is None:
            presorted = self.presort(feature_index, feature_value, handle=self._inline)
        elif presorted.src != ShardedLookups.key_of(feature_index, feature_value):
            raise ValueError("presorted handle was made for other feature_index / feature_value tensors")
        first, fm, emb = _ShardedFunction.apply(self._anchor, self.bias, self, idx, val, train, presorted)
        if self.check_bounds and int(self.oob_flag.item()) != 0:
            self.oob_flag.zero_()
            raise IndexError("feature_index out of range for its field")
        return first, fm, emb
