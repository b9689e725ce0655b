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
