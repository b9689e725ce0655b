"""TEST INFRASTRUCTURE / CPU BASELINE -- PyTorch-CPU op-by-op restatement of the reference graph.

PARITY UNPINNED at the TensorFlow boundary: TF 1.x cannot be imported here, so the reference's
own CPU path cannot be timed.  This module rebuilds the same graph of separate, materialising ops
on torch CPU kernels (kind "port" in bench.py's cpu_baseline), as BASELINE.md section 3 lays out:
one embedding variable per column (models/DeepFM/deepFM.py:385-390), concat, reshape, the nine
elementwise / reduce ops of fm_logit_fn (:329-334), per-column first-order gathers (:258-263),
SUM-reduced sigmoid CE (:72), and Adagrad(initial_accumulator_value=0.1, eps=0) on the sparse
gradients ([TF] dedupe + SparseApplyAdagrad); and the L-step Python loop of
x0 * (xl @ w)[:, None] + b + xl (models/DeepCrossNetwork/DeepCrossNetwork.py:345-346, 363-365).
Only tests/ and bench.py's cpu_baseline / --impl reference legs may import it.
"""
import time

import torch
import torch.nn.functional as Fn


class DeepFMLayerCPU:
    def __init__(self, rows_per_field, embedding_size, lr=0.05, table=None, w1=None, seed=0):
        g = torch.Generator().manual_seed(seed)
        self.K = embedding_size
        self.emb, self.lin = [], []
        off = 0
        for n in rows_per_field:
            if table is not None:
                t = torch.tensor(table[off:off + n])
                l = torch.tensor(w1[off:off + n]).reshape(n, 1)
            else:
                t = torch.randn((n, embedding_size), generator=g) / embedding_size ** 0.5
                l = torch.randn((n, 1), generator=g) * 0.01
            self.emb.append(t.requires_grad_(True))
            self.lin.append(l.requires_grad_(True))
            off += n
        self.bias = torch.zeros(1, requires_grad=True)
        self.opt = torch.optim.Adagrad(self.emb + self.lin, lr=lr, initial_accumulator_value=0.1, eps=0)

    def step(self, idx, val, labels, U=None):
        """idx [B,F] int64, val [B,F] fp32, labels [B]; U [B,F*K] stands in for the DNN's upstream."""
        F = idx.shape[1]
        cols, firsts = [], []
        for f in range(F):                                                  # one op chain per column
            v = val[:, f:f + 1]
            cols.append(Fn.embedding(idx[:, f], self.emb[f], sparse=True) * v)
            firsts.append(Fn.embedding(idx[:, f], self.lin[f], sparse=True) * v)
        net = torch.cat(cols, dim=1)                                        # deepFM.py:328
        e = net.reshape(-1, F, self.K)                                      # :329
        summed_squared = torch.square(torch.sum(e, -2))                     # :331
        squared_summed = torch.sum(torch.square(e), -2)                     # :332
        fm = 0.5 * torch.sum(summed_squared - squared_summed, -1, keepdim=True)   # :333-334
        first = torch.stack(firsts, 0).sum(0) + self.bias                   # linear_model AddN + bias
        logits = first + fm                                                 # :218-223
        loss = Fn.binary_cross_entropy_with_logits(logits[:, 0], labels, reduction="sum")   # :72
        if U is not None:
            loss = loss + (net * U).sum()
        self.opt.zero_grad(set_to_none=True)
        loss.backward()
        self.opt.step()
        return logits.detach()


def cross_step_cpu(x0, w, b, dy):
    """fwd + bwd of the cross stack on torch CPU ops; returns (xL, dx0, dw, db)."""
    x0 = x0.detach().requires_grad_(True)
    w = w.detach().requires_grad_(True)
    b = b.detach().requires_grad_(True)
    xl = x0
    for l in range(w.shape[0]):
        xl = x0 * (xl @ w[l])[:, None] + b[l] + xl
    xl.backward(dy)
    return xl.detach(), x0.grad, w.grad, b.grad


def time_deepfm_layer(rows_per_field, K, batches, lr=0.05, warmup=1, steps=3, threads=None, with_upstream=True):
    """samples/s of the restatement on this host; `batches` = list of (idx, val, labels) numpy."""
    if threads:
        torch.set_num_threads(threads)
    model = DeepFMLayerCPU(rows_per_field, K, lr=lr)
    tb = [(torch.as_tensor(i), torch.as_tensor(v), torch.as_tensor(y)) for i, v, y in batches]
    B, F = tb[0][0].shape
    U = torch.randn((B, F * K)) * 1e-2 if with_upstream else None
    for s in range(warmup):
        model.step(*tb[s % len(tb)], U)
    t0 = time.perf_counter()
    for s in range(steps):
        model.step(*tb[s % len(tb)], U)
    dt = time.perf_counter() - t0
    return B * steps / dt, dt / steps
