"""TEST INFRASTRUCTURE -- CPU restatement (numpy) of the reference's two full Deep-CTR graphs around the
hot path: the DNN tower, the logit sums, the heads' losses and the dense optimizers (SURVEY.md
section 8f rank 1).  PARITY UNPINNED like the rest of oracle/ (TensorFlow cannot run here); pinned
instead by torch autograd in tests/test_oracle_models.py.  Only tests/ may import this.

  DeepFM   models/DeepFM/deepFM.py:143-252   logits = linear_logits + fm_logit + dnn_logit (:218-223,
           :337-338), SUM-reduced sigmoid cross entropy (:72, :107-111), Adagrad on the 'dnn_fm' scope
           (embedding tables + tower, :230-234), the linear scope's own optimizer (:236-241)
  DCN      models/DeepCrossNetwork/DeepCrossNetwork.py:124-141  x0 = input_layer; cross stack || deep
           tower; concat; dense(1).  MEAN-reduced loss (:209-225); every gradient tensor clipped to
           L2 norm 100 (:282-289)
Dense kernels are stored TF-style, [in, out]: y = x @ W + b (tf.layers.dense).
"""
import numpy as np

from . import deepctr_oracle as O
from . import tf_semantics as tfs


# ---------------------------------------------------------------------------- DNN tower
def mlp_forward(x, weights, biases, final_linear):
    """dnn_logit_fn (deepFM.py:284-319) / _deep_architecture (DeepCrossNetwork.py:370-410) without
    dropout / batch norm: relu(x W_i + b_i) per hidden layer; when `final_linear`, the last
    (W, b) pair is the activation-free `logits` layer (deepFM.py:311-317)."""
    acts = [x]
    n = len(weights)
    for i, (W, b) in enumerate(zip(weights, biases)):
        z = acts[-1] @ W + b
        if not (final_linear and i == n - 1):
            z = np.maximum(z, 0)
        acts.append(z)
    return acts[-1], acts


def mlp_backward(acts, weights, dout, final_linear):
    """-> dx, [dW_i], [db_i] (TF autodiff of the above)."""
    n = len(weights)
    dWs, dbs = [None] * n, [None] * n
    d = dout
    for i in range(n - 1, -1, -1):
        if not (final_linear and i == n - 1):
            d = d * (acts[i + 1] > 0)
        dWs[i] = acts[i].T @ d
        dbs[i] = d.sum(axis=0)
        d = d @ weights[i].T
    return d, dWs, dbs


def dense_adagrad(var, accum, grad, lr):
    """[TF] ApplyAdagrad: accum += g*g; var -= lr*g/sqrt(accum) (no epsilon).  In place."""
    accum += grad * grad
    var -= var.dtype.type(lr) * grad / np.sqrt(accum)


def sigmoid_ce_grad(logits, labels, reduction):
    """d loss / d logits for sigmoid cross entropy with SUM (deepFM.py:72) or MEAN
    (DeepCrossNetwork.py:221-223) reduction, and the loss itself."""
    per = tfs.sigmoid_cross_entropy_with_logits(labels, logits)
    g = tfs.sigmoid(logits) - labels
    if reduction == "mean":
        return per.mean(), g / logits.shape[0]
    return per.sum(), g


# ---------------------------------------------------------------------------- DeepFM
def deepfm_forward(p, field_offset, idx, val):
    """p: dict(table[N,K], w1[N], bias, W[list], b[list]) -> logits[B], cache."""
    dt = p["table"].dtype.type
    e, _ = O.embedding_lookup(p["table"], field_offset, idx, val, "sum", dt)
    first = O.first_order(p["w1"], p["bias"], field_offset, idx, val, dt)[:, 0]
    fm = O.fm_second_order(e)[:, 0]
    B = idx.shape[0]
    dnn, acts = mlp_forward(e.reshape(B, -1), p["W"], p["b"], final_linear=True)
    logits = first + fm + dnn[:, 0]                      # add_n (:223) of linear and (fm + dnn) (:337-338)
    return logits, dict(e=e, acts=acts, first=first, fm=fm)


def deepfm_train_step(p, st, field_offset, idx, val, labels, lr, linear_lr=None):
    """One optimizer step; mutates p (parameters) and st (Adagrad accumulators: same keys) in place.
    The linear scope (w1, bias) is updated with Adagrad at linear_lr (default lr) -- see ftrl.py for the
    reference's default 'Ftrl'.  Returns dict(loss, logits, G, rows)."""
    dt = p["table"].dtype.type
    linear_lr = lr if linear_lr is None else linear_lr
    logits, c = deepfm_forward(p, field_offset, idx, val)
    loss, g = sigmoid_ce_grad(logits, labels.astype(dt), "sum")
    B = idx.shape[0]
    du, dWs, dbs = mlp_backward(c["acts"], p["W"], g[:, None], final_linear=True)
    rows, G, g1, dbias = O.embedding_backward(p["table"], field_offset, idx, val, g, g,
                                              du.reshape(c["e"].shape), "sum", dt)
    O.sparse_adagrad(p["table"], st["table"], rows, G, lr)
    O.sparse_adagrad(p["w1"], st["w1"], rows, g1, linear_lr)
    bias, bacc = np.asarray([p["bias"]], dtype=dt), np.asarray([st["bias"]], dtype=dt)
    dense_adagrad(bias, bacc, np.asarray([dbias], dtype=dt), linear_lr)
    p["bias"], st["bias"] = bias[0], bacc[0]
    for i in range(len(p["W"])):
        dense_adagrad(p["W"][i], st["W"][i], dWs[i], lr)
        dense_adagrad(p["b"][i], st["b"][i], dbs[i], lr)
    return dict(loss=loss, logits=logits, rows=rows, G=G, g1=g1)


# ---------------------------------------------------------------------------- DCN
def dcn_forward(p, field_offset, idx, val):
    """p: dict(table, cross_w[L,d], cross_b[L,d], W[list], b[list], Wl[d+h,1], bl[1]) -> logits[B], cache."""
    dt = p["table"].dtype.type
    e, _ = O.embedding_lookup(p["table"], field_offset, idx, val, "sum", dt)
    B = idx.shape[0]
    x0 = e.reshape(B, -1)
    xL, _ = O.cross_forward(x0, p["cross_w"], p["cross_b"])
    deep, acts = mlp_forward(x0, p["W"], p["b"], final_linear=False)
    m = np.concatenate([xL, deep], axis=-1)              # DeepCrossNetwork.py:136
    logits = (m @ p["Wl"] + p["bl"])[:, 0]               # :137
    return logits, dict(e=e, x0=x0, xL=xL, acts=acts, m=m)


def dcn_train_step(p, st, field_offset, idx, val, labels, lr, clip_norm=100.0, clip_tables=True):
    """One Adagrad step with per-tensor clip_by_norm (DeepCrossNetwork.py:282-289); in place.
    clip_tables=False leaves the embedding gradient unclipped (what the sharded CUDA layer does, see models.DCN)."""
    dt = p["table"].dtype.type
    logits, c = dcn_forward(p, field_offset, idx, val)
    loss, g = sigmoid_ce_grad(logits, labels.astype(dt), "mean")
    d = p["cross_w"].shape[1]
    dWl = c["m"].T @ g[:, None]
    dbl = np.asarray([g.sum()], dtype=dt)
    dm = g[:, None] @ p["Wl"].T
    dx0_c, dcw, dcb = O.cross_backward(c["x0"], p["cross_w"], p["cross_b"], dm[:, :d])
    dx0_d, dWs, dbs = mlp_backward(c["acts"], p["W"], dm[:, d:], final_linear=False)
    u = (dx0_c + dx0_d).reshape(c["e"].shape)
    zero = np.zeros(idx.shape[0], dtype=dt)
    rows, G, _, _ = O.embedding_backward(p["table"], field_offset, idx, val, zero, zero, u, "sum", dt)
    clip = (lambda t: tfs.clip_by_norm(t, clip_norm)) if clip_norm else (lambda t: t)
    G_unclipped = G
    if clip_tables and clip_norm:        # one variable per column: each column's IndexedSlices clipped on its own
        G = tfs.clip_indexed_slices_per_column(rows, G, field_offset, p["table"].shape[0], clip_norm)
    O.sparse_adagrad(p["table"], st["table"], rows, G, lr)
    dense_adagrad(p["cross_w"], st["cross_w"], clip(dcw), lr)
    dense_adagrad(p["cross_b"], st["cross_b"], clip(dcb), lr)
    for i in range(len(p["W"])):
        dense_adagrad(p["W"][i], st["W"][i], clip(dWs[i]), lr)
        dense_adagrad(p["b"][i], st["b"][i], clip(dbs[i]), lr)
    dense_adagrad(p["Wl"], st["Wl"], clip(dWl), lr)
    dense_adagrad(p["bl"], st["bl"], clip(dbl), lr)
    return dict(loss=loss, logits=logits, rows=rows, G=G, G_unclipped=G_unclipped)


# ---------------------------------------------------------------------------- parameter sets
def glorot_uniform(rng, fan_in, fan_out, dtype):
    """init_ops.glorot_uniform_initializer (deepFM.py:300, :315)."""
    lim = np.sqrt(6.0 / (fan_in + fan_out))
    return rng.uniform(-lim, lim, size=(fan_in, fan_out)).astype(dtype)


def make_tower(rng, d_in, hidden, final_units, dtype=np.float32):
    sizes = [d_in] + list(hidden) + ([final_units] if final_units else [])
    W = [glorot_uniform(rng, a, b, dtype) for a, b in zip(sizes[:-1], sizes[1:])]
    b = [(rng.standard_normal(n) * 0.01).astype(dtype) for n in sizes[1:]]     # TF: zeros; non-trivial here
    return W, b


def adagrad_state(p, init=tfs.ADAGRAD_INITIAL_ACCUMULATOR):
    out = {}
    for k, v in p.items():
        if isinstance(v, list):
            out[k] = [np.full_like(a, init) for a in v]
        elif np.ndim(v) == 0:
            out[k] = np.asarray(v).dtype.type(init)
        else:
            out[k] = np.full_like(v, init)
    return out
