"""Host -> device input feeder for the layer: the tf.data `input_fn` side of the reference
(models/DeepCrossNetwork/train.py:127-156 hands decoded batches to the graph) reduced to what the
hot path needs -- pinned host batches of (feature_index, feature_value, labels) copied into static
device slots on a copy stream, double-buffered so the H2D of batch i+1 overlaps the kernels of
batch i.  Static slots keep device addresses fixed, so a captured CUDA graph can be replayed on them.
"""
import torch


class HostFeeder:
    def __init__(self, *slots):
        """slots: two or more lists of preallocated device tensors with identical shapes."""
        if not slots:
            raise ValueError("HostFeeder needs at least one slot of device tensors")
        for s in slots:
            for t in s:
                if not t.is_cuda:
                    raise ValueError("HostFeeder slots must be CUDA tensors")
        self.slots = [list(s) for s in slots]
        self.copy_stream = torch.cuda.Stream(device=self.slots[0][0].device)
        self.ready = [torch.cuda.Event() for _ in slots]
        self.free = [torch.cuda.Event() for _ in slots]

    def prefetch(self, slot, host_tensors):
        """Enqueue the H2D copies of one batch into `slot` (after its previous consumer finished)."""
        dst = self.slots[slot]
        if len(host_tensors) != len(dst):
            raise ValueError("expected %d tensors per batch" % len(dst))
        with torch.cuda.stream(self.copy_stream):
            self.copy_stream.wait_event(self.free[slot])
            for d, h in zip(dst, host_tensors):
                d.copy_(h, non_blocking=True)
            self.ready[slot].record(self.copy_stream)

    def wait(self, slot):
        """Make the current (compute) stream wait for `slot`'s copies; returns the device tensors."""
        torch.cuda.current_stream().wait_event(self.ready[slot])
        return self.slots[slot]

    def release(self, slot):
        """Mark `slot` consumed by everything enqueued so far on the current stream."""
        self.free[slot].record(torch.cuda.current_stream())


class ColumnFeeder(HostFeeder):
    """HostFeeder for the reference's native input form: one tensor per column.

    The reference's `input_fn` yields a dict of columns -- int ids for categorical columns, floats for
    numeric ones (models/DeepCrossNetwork/train.py:127-156; consumed at models/DeepFM/deepFM.py:159-177).
    The resolved [B,F] feature_index / feature_value pair carries a constant 1.0 for every categorical
    field and a constant id 0 for every numeric (one-row) field, so only the columns cross PCIe:
    `sparse_index[B, n_sparse]` (int32 or int64) and `dense_value[B, n_dense]`.  `dir_expand_features`
    widens them into the slot's feature_index / feature_value on the copy stream right behind the copy.

    slots: lists [feature_index[B,F] int64, feature_value[B,F] fp32, *others] of device tensors;
    sparse_fields / dense_fields: which fields the columns of sparse_index / dense_value feed, in order.
    Fields named in neither list keep (id 0, value 1.0).
    prefetch(slot, [sparse_index_host, dense_value_host, *others_host]).
    """

    def __init__(self, sparse_fields, dense_fields, *slots, index_dtype=torch.int32):
        super().__init__(*slots)
        from . import _lib
        self._lib = _lib
        if index_dtype not in (torch.int32, torch.int64):
            raise ValueError("index_dtype must be torch.int32 or torch.int64")
        idx0 = self.slots[0][0]
        if idx0.dim() != 2 or idx0.dtype != torch.int64 or self.slots[0][1].shape != idx0.shape:
            raise ValueError("each slot must start with feature_index[B,F] int64, feature_value[B,F] fp32")
        B, F = idx0.shape
        sparse_fields, dense_fields = [int(f) for f in sparse_fields], [int(f) for f in dense_fields]
        both = sparse_fields + dense_fields
        if len(set(both)) != len(both) or any(not 0 <= f < F for f in both):
            raise ValueError("sparse_fields / dense_fields must be distinct fields in [0, F)")
        src = [-(F + 1)] * F
        for j, f in enumerate(sparse_fields):
            src[f] = j
        for j, f in enumerate(dense_fields):
            src[f] = -(j + 1)
        if -(F + 1) in src:
            raise ValueError("every field must be fed by a column of sparse_index or dense_value")
        dev = idx0.device
        self.n_sparse, self.n_dense, self.index_dtype = len(sparse_fields), len(dense_fields), index_dtype
        self.field_src = torch.tensor(src, dtype=torch.int32, device=dev)
        # device landing buffers of the columns, one pair per slot
        self.columns = [(torch.empty((B, max(self.n_sparse, 1)), dtype=index_dtype, device=dev),
                         torch.empty((B, max(self.n_dense, 1)), dtype=torch.float32, device=dev))
                        for _ in self.slots]

    def bytes_per_batch(self, host_tensors):
        return sum(t.numel() * t.element_size() for t in host_tensors)

    def prefetch(self, slot, host_tensors):
        dst = self.slots[slot]
        if len(host_tensors) != len(dst):
            raise ValueError("expected [sparse_index, dense_value] + %d more tensors per batch" % (len(dst) - 2))
        sp_h, de_h = host_tensors[0], host_tensors[1]
        B, F = dst[0].shape
        if sp_h.dtype != self.index_dtype or tuple(sp_h.shape) != (B, self.n_sparse):
            raise ValueError("sparse_index must be [%d, %d] %s" % (B, self.n_sparse, self.index_dtype))
        if de_h.dtype != torch.float32 or tuple(de_h.shape) != (B, self.n_dense):
            raise ValueError("dense_value must be [%d, %d] float32" % (B, self.n_dense))
        sp_d, de_d = self.columns[slot]
        L = self._lib.lib()
        with torch.cuda.stream(self.copy_stream):
            # the columns land in buffers only this stream touches: their copies need not wait for the
            # slot's previous consumer, only the widening (and the other tensors of the slot) does
            if self.n_sparse:
                sp_d.copy_(sp_h, non_blocking=True)
            if self.n_dense:
                de_d.copy_(de_h, non_blocking=True)
            self.copy_stream.wait_event(self.free[slot])
            for d, h in zip(dst[2:], host_tensors[2:]):
                d.copy_(h, non_blocking=True)
            self._lib.check(L.dir_expand_features(
                self._lib.ptr(sp_d), 4 if self.index_dtype == torch.int32 else 8, self._lib.ptr(de_d),
                self._lib.ptr(self.field_src), B, F, self.n_sparse, self.n_dense, self._lib.ptr(dst[0]),
                self._lib.ptr(dst[1]), self.copy_stream.cuda_stream), "dir_expand_features")
            self.ready[slot].record(self.copy_stream)
