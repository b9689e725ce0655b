"""TEST INFRASTRUCTURE -- CPU restatement (numpy) of the reference's Deep-CTR hot path.

PARITY UNPINNED at the TensorFlow boundary (see oracle/tf_semantics.py): the
reference's TF 1.x graph cannot run here and the reference ships no golden
vectors.  This module restates, op by op and in the reference's evaluation
order, the lines cited on each function; tests/test_oracle.py pins it against
closed-form known-answer tests and against torch autograd (an independent
implementation).  Only tests/, __graft_entry__.smoke() and bench.py's
cpu_baseline / --impl reference legs may import it; the product never does.

Layout: ONE concatenated table `table[N, K]` (the reference holds one variable
per column, deepFM.py:385-390; concatenation is a storage choice) with
`field_offset[F]`, global row = field_offset[f] + feature_index[b, f].
`dtype` selects fp32 (the reference's arithmetic) or fp64 (the adjudicator).
"""
import numpy as np

from . import tf_semantics as tfs


def global_rows(feature_index, field_offset):
    """int64 [B,F]; the index arithmetic that must be bit-exact."""
    return feature_index.astype(np.int64) + np.asarray(field_offset, dtype=np.int64)[None, :]


# ----------------------------------------------------------------------------
# forward
# ----------------------------------------------------------------------------
def embedding_lookup(table, field_offset, feature_index, feature_value=None,
                     combiner="sum", dtype=np.float32):
    """models/DeepFM/deepFM.py:363-400 (myself_input_layer): per column
    `column._get_dense_tensor` (:387) then reshape to [B, K] (:393).
    Returns e[B, F, K] (column c at [:, c, :]) and the keep mask.
    """
    dt = np.dtype(dtype).type
    eff, keep = tfs.effective_value(feature_index, feature_value, combiner, dt)
    rows = global_rows(np.where(keep, feature_index, 0), field_offset)
    e = table[rows].astype(dt) * eff[:, :, None]
    return e, keep


def first_order(w1, bias, field_offset, feature_index, feature_value=None,
                dtype=np.float32):
    """models/DeepFM/deepFM.py:255-263 linear_model(units=1, sparse_combiner='sum'):
    per column sum-combined scalar weight, summed over columns in column order
    (AddN), plus bias.  -> [B, 1]
    """
    dt = np.dtype(dtype).type
    eff, keep = tfs.effective_value(feature_index, feature_value, "sum", dt)
    rows = global_rows(np.where(keep, feature_index, 0), field_offset)
    per_col = w1[rows].astype(dt) * eff                     # [B, F]
    acc = np.zeros(per_col.shape[0], dtype=dt)
    for f in range(per_col.shape[1]):                       # AddN in column order
        acc = acc + per_col[:, f]
    return (acc + dt(bias))[:, None]


def fm_second_order(e):
    """models/DeepFM/deepFM.py:329-334, op for op:
    summed_squared = square(reduce_sum(e, -2)); squared_summed = reduce_sum(square(e), -2)
    logits = 0.5 * reduce_sum(summed_squared - squared_summed, -1); expand_dims(-1)
    """
    dt = e.dtype.type
    summed_squared = np.square(np.sum(e, axis=-2, dtype=e.dtype))
    squared_summed = np.sum(np.square(e), axis=-2, dtype=e.dtype)
    logits = dt(0.5) * np.sum(summed_squared - squared_summed, axis=-1, dtype=e.dtype)
    return logits[:, None]


def fm_pairwise(e):
    """Sum_{i<j} <v_i, v_j>: the definition fm_second_order simplifies (KAT-2)."""
    B, F, K = e.shape
    out = np.zeros(B, dtype=e.dtype)
    for i in range(F):
        for j in range(i + 1, F):
            out += np.sum(e[:, i, :] * e[:, j, :], axis=-1)
    return out[:, None]


def cross_op(x0, x, w, b):
    """models/DeepCrossNetwork/DeepCrossNetwork.py:336-347:
    x_w = tensordot(x, w, axes=1); y = x0 * expand_dims(x_w, -1) + b + x
    evaluated ((x0*s) + b) + x.
    """
    x_w = np.tensordot(x, w, axes=1).astype(x.dtype)
    return x0 * x_w[:, None] + b + x, x_w


def cross_forward(x0, cross_w, cross_b):
    """models/DeepCrossNetwork/DeepCrossNetwork.py:350-367 (_cross_architecture).
    Returns x_L[B,d] and s[B,L] (s[:, l] = x_l . w_l)."""
    xl = x0
    s = np.zeros((x0.shape[0], cross_w.shape[0]), dtype=x0.dtype)
    for l in range(cross_w.shape[0]):
        xl, s[:, l] = cross_op(x0, xl, cross_w[l], cross_b[l])
    return xl, s


# ----------------------------------------------------------------------------
# backward (TF autodiff of the ops above, restated analytically)
# ----------------------------------------------------------------------------
def cross_backward(x0, cross_w, cross_b, dy):
    """Reverse sweep of _cross_architecture.
    -> dx0[B,d], dw[L,d], db[L,d]
    """
    L = cross_w.shape[0]
    xs = [x0]
    for l in range(L):
        xs.append(cross_op(x0, xs[-1], cross_w[l], cross_b[l])[0])
    dx = dy.copy()
    dx0 = np.zeros_like(x0)
    dw = np.zeros_like(cross_w)
    db = np.zeros_like(cross_b)
    for l in range(L - 1, -1, -1):
        s_l = np.tensordot(xs[l], cross_w[l], axes=1).astype(x0.dtype)
        db[l] = dx.sum(axis=0, dtype=x0.dtype)
        ds = np.sum(dx * x0, axis=1, dtype=x0.dtype)
        dx0 = dx0 + dx * s_l[:, None]
        dw[l] = np.sum(ds[:, None] * xs[l], axis=0, dtype=x0.dtype)
        dx = dx + ds[:, None] * cross_w[l][None, :]
    dx0 = dx0 + dx
    return dx0, dw, db


def embedding_backward(table, field_offset, feature_index, feature_value,
                       g_first, g_fm, u=None, combiner="sum", dtype=np.float32, return_abs=False):
    """Backward of lookup + first order + FM (implicit in
    optimizer.minimize, models/DeepFM/deepFM.py:230-241).

    g_first[B], g_fm[B]: dL/d(first_order), dL/d(fm_second_order);
    u[B,F,K] or None:    dL/d(embeddings) from the DNN / cross consumer.
    Returns (rows[U] int64 sorted unique, G[U,K], g1[U], dbias): per-unique-row
    gradients, duplicates summed in sample order ([TF] Unique + UnsortedSegmentSum,
    then _deduplicate_indexed_slices).  Pruned lookups contribute nothing and do
    not touch their row.  return_abs=True appends Gabs[U,K] = sum |per-lookup term| per row:
    the magnitude being summed, i.e. the honest denominator for a relative error on G.
    """
    dt = np.dtype(dtype).type
    eff, keep = tfs.effective_value(feature_index, feature_value, combiner, dt)
    eff1, _ = tfs.effective_value(feature_index, feature_value, "sum", dt)
    rows = global_rows(np.where(keep, feature_index, 0), field_offset)
    e = table[rows].astype(dt) * eff[:, :, None]
    S = np.sum(e, axis=1, dtype=e.dtype)                                # [B,K]
    de = g_fm.astype(dt)[:, None, None] * (S[:, None, :] - e)
    if u is not None:
        de = de + u.astype(dt)
    per_lookup = eff[:, :, None] * de                                   # [B,F,K]
    per_lookup1 = g_first.astype(dt)[:, None] * eff1                    # [B,F]
    flat_rows = rows[keep]                                              # sample-major order
    uniq, inv = np.unique(flat_rows, return_inverse=True)
    G = np.zeros((uniq.shape[0], table.shape[1]), dtype=dt)
    g1 = np.zeros(uniq.shape[0], dtype=dt)
    np.add.at(G, inv, per_lookup[keep])                                 # sequential, sample order
    np.add.at(g1, inv, per_lookup1[keep])
    dbias = np.sum(g_first.astype(dt), dtype=dt)
    if return_abs:
        Gabs = np.zeros_like(G)
        g1abs = np.zeros_like(g1)
        np.add.at(Gabs, inv, np.abs(per_lookup[keep]))
        np.add.at(g1abs, inv, np.abs(per_lookup1[keep]))
        return uniq, G, g1, dbias, Gabs, g1abs
    return uniq, G, g1, dbias


# ----------------------------------------------------------------------------
# sparse row-wise optimizers
# ----------------------------------------------------------------------------
def sparse_adagrad(var, accum, rows, grad, lr):
    """[TF] SparseApplyAdagrad after dedupe: accum[r] += g*g; var[r] -= lr*g/sqrt(accum[r]).
    No epsilon; rows not listed are untouched.  In place; `rows` unique."""
    dt = var.dtype.type
    a = accum[rows] + grad * grad
    accum[rows] = a
    var[rows] = var[rows] - dt(lr) * grad / np.sqrt(a)


def sparse_ftrl(var, accum, linear, rows, grad, lr, l1=0.0, l2=0.0):
    """[TF] SparseApplyFtrl after dedupe, tf.train.FtrlOptimizer defaults learning_rate_power = -0.5,
    l2_shrinkage = 0 (linear_optimizer='Ftrl', models/DeepFM/deepFM.py:58, 236-241):
        n' = n + g^2;  sigma = (sqrt(n') - sqrt(n)) / lr;  z += g - sigma*w
        w  = |z| > l1 ? (sign(z)*l1 - z) / (sqrt(n')/lr + 2*l2) : 0
    accum starts at initial_accumulator_value = 0.1, linear at 0.  In place; `rows` unique."""
    dt = var.dtype.type
    n, w = accum[rows], var[rows]
    nn = n + grad * grad
    sigma = (np.sqrt(nn) - np.sqrt(n)) / dt(lr)
    z = linear[rows] + grad - sigma * w
    quad = np.sqrt(nn) / dt(lr) + dt(2.0) * dt(l2)
    var[rows] = np.where(np.abs(z) > dt(l1), (np.sign(z) * dt(l1) - z) / quad, dt(0))
    accum[rows] = nn
    linear[rows] = z


def sparse_proximal_adagrad(var, accum, rows, grad, lr, l1=0.0, l2=0.0):
    """[TF] SparseApplyProximalAdagrad behind tf.train.ProximalAdagradOptimizer(learning_rate, initial_accumulator_value
    = 0.1, l1, l2) -- the reference's dnn_optimizer in models/ESMM/train.py:137-139 -- on de-duplicated rows, in place:
        a += g^2;  eta = lr / sqrt(a);  p = v - g * eta
        v = sign(p) * max(|p| - eta * l1, 0) / (1 + l2 * eta)          (l1 = 0: v = p / (1 + l2 * eta))"""
    dt = var.dtype.type
    accum[rows] = accum[rows] + grad * grad
    eta = dt(lr) / np.sqrt(accum[rows])
    p = var[rows] - grad * eta
    den = dt(1.0) + dt(l2) * eta
    if l1 > 0:
        var[rows] = np.sign(p) * np.maximum(np.abs(p) - eta * dt(l1), dt(0.0)) / den
    else:
        var[rows] = p / den


def sparse_sgd(var, rows, grad, lr):
    """[TF] ScatterSub of the de-duplicated IndexedSlices: var[r] -= lr*g."""
    var[rows] = var[rows] - var.dtype.type(lr) * grad


# ----------------------------------------------------------------------------
# one training step of the layer, as the bench times it
# ----------------------------------------------------------------------------
def deepfm_layer_step(table, accum, w1, accum1, bias, field_offset, feature_index,
                      feature_value, labels, lr, u=None, optimizer="adagrad",
                      update_first_order=True, dtype=np.float32):
    """fwd (lookup + first order + FM) -> SUM-reduced sigmoid CE (deepFM.py:72)
    -> bwd -> fused sparse update.  Mutates table/accum/w1/accum1 in place.
    Returns dict(first, fm, logits, e, rows, G, g1, dbias, g)."""
    dt = np.dtype(dtype).type
    e, _ = embedding_lookup(table, field_offset, feature_index, feature_value, "sum", dt)
    first = first_order(w1, bias, field_offset, feature_index, feature_value, dt)
    fm = fm_second_order(e)
    logits = first + fm                                   # deepFM.py:218-223 add_n
    g = (tfs.sigmoid(logits[:, 0]) - labels.astype(dt)).astype(dt)     # SUM reduction: no 1/B
    rows, G, g1, dbias = embedding_backward(table, field_offset, feature_index,
                                            feature_value, g, g, u, "sum", dt)
    if optimizer == "adagrad":
        sparse_adagrad(table, accum, rows, G, lr)
        if update_first_order:
            sparse_adagrad(w1, accum1, rows, g1, lr)
    elif optimizer == "sgd":
        sparse_sgd(table, rows, G, lr)
        if update_first_order:
            sparse_sgd(w1, rows, g1, lr)
    else:
        raise ValueError("optimizer must be 'adagrad' or 'sgd'")
    return dict(first=first, fm=fm, logits=logits, e=e, rows=rows, G=G, g1=g1,
                dbias=dbias, g=g)


# ----------------------------------------------------------------------------
# multi-hot / weighted bags
# ----------------------------------------------------------------------------
def _bag_entries(field_offset, bag_offsets, bag_index, bag_weight, F, dt):
    """Per entry: slot s = b*F+f, field, global row, weight, keep ([TF] _safe_embedding_lookup_sparse
    prunes id < 0 and weight <= 0)."""
    nnz = bag_index.shape[0]
    n_slots = bag_offsets.shape[0] - 1
    slot = np.repeat(np.arange(n_slots, dtype=np.int64), np.diff(bag_offsets))
    field = slot % F
    w = np.ones(nnz, dtype=dt) if bag_weight is None else bag_weight.astype(dt)
    keep = (bag_index >= 0) & (w > 0)
    rows = np.where(keep, bag_index, 0) + np.asarray(field_offset, dtype=np.int64)[field]
    return slot, rows, w, keep


def embedding_bag_lookup(table, w1, bias, field_offset, bag_offsets, bag_index, bag_weight, B, F,
                         combiner="mean", dtype=np.float32):
    """myself_input_layer with multi-hot columns (models/DeepFM/deepFM.py:53, 77, 363-400) + the linear
    model on the same bags (:255-263, sparse_combiner='sum').  [TF] embedding_lookup_sparse:
    e_s = sum_i w_i T[id_i] / norm, norm = 1 ('sum'), sum w ('mean'), sqrt(sum w^2) ('sqrtn'); entries summed
    in order; empty bag -> zeros.  -> e[B,F,K], first[B,1], x[nnz] (effective scale w/norm, 0 if pruned)."""
    dt = np.dtype(dtype).type
    slot, rows, w, keep = _bag_entries(field_offset, bag_offsets, bag_index, bag_weight, F, dt)
    K = table.shape[1]
    wk = np.where(keep, w, dt(0))
    acc = np.zeros((B * F, K), dtype=dt)
    np.add.at(acc, slot[keep], wk[keep][:, None] * table[rows[keep]].astype(dt))      # sequential, entry order
    wsum = np.zeros(B * F, dtype=dt)
    wsq = np.zeros(B * F, dtype=dt)
    np.add.at(wsum, slot[keep], wk[keep])
    np.add.at(wsq, slot[keep], wk[keep] * wk[keep])
    if combiner == "sum":
        norm = np.ones(B * F, dtype=dt)
    elif combiner == "mean":
        norm = wsum
    elif combiner == "sqrtn":
        norm = np.sqrt(wsq)
    else:
        raise ValueError("combiner must be 'sum', 'mean' or 'sqrtn'")
    safe = np.where(wsum > 0, norm, dt(1))
    e = np.where((wsum > 0)[:, None], acc / safe[:, None], dt(0)).astype(dt)
    x = np.where(keep, wk / safe[slot], dt(0)).astype(dt)
    lin = np.zeros(B * F, dtype=dt)
    np.add.at(lin, slot[keep], wk[keep] * w1[rows[keep]].astype(dt))
    first = np.zeros(B, dtype=dt)
    linf = lin.reshape(B, F)
    for f in range(F):                                       # AddN in column order
        first = first + linf[:, f]
    return e.reshape(B, F, K), (first + dt(bias))[:, None], x


def embedding_bag_backward(table, field_offset, bag_offsets, bag_index, bag_weight, e, x, g_first, g_fm, u, B, F,
                           dtype=np.float32):
    """Backward of the bag lookup + first order + FM (TF autodiff; weights carry no gradient):
    entry j of slot s=(b,f):  dT[row_j] += x_j * (g_fm[b] (S_b - e_s) + u_s),  dw1[row_j] += w_j * g_first[b].
    -> (rows[U] sorted unique, G[U,K], g1[U]) with duplicates summed in entry order."""
    dt = np.dtype(dtype).type
    slot, rows, w, keep = _bag_entries(field_offset, bag_offsets, bag_index, bag_weight, F, dt)
    S = np.sum(e, axis=1, dtype=e.dtype)                                     # [B,K]
    de = g_fm.astype(dt)[:, None, None] * (S[:, None, :] - e)
    if u is not None:
        de = de + u.astype(dt)
    de = de.reshape(B * F, -1)
    per = x[:, None] * de[slot]                                              # [nnz,K]
    per1 = g_first.astype(dt)[slot // F] * np.where(keep, w, dt(0))
    uniq, inv = np.unique(rows[keep], return_inverse=True)
    G = np.zeros((uniq.shape[0], table.shape[1]), dtype=dt)
    g1 = np.zeros(uniq.shape[0], dtype=dt)
    np.add.at(G, inv, per[keep])
    np.add.at(g1, inv, per1[keep])
    return uniq, G, g1


# ----------------------------------------------------------------------------
# tf.feature_column.input_layer in the reference's DCN convention
# ----------------------------------------------------------------------------
def input_layer(columns, numeric, indicator_ids, emb):
    """models/DeepCrossNetwork/DeepCrossNetwork.py:126 over the columns of train.py:88-100.
    columns: [(name, kind, size)]; inputs in the listing order of each kind; output = the dense columns
    concatenated in NAME order ([TF] input_layer sorts by name): numeric 1-wide, indicator one-hot
    (id outside [0, size) -> zeros), embedding K-wide.  -> x0[B, d] and the list of (name, start, stop)."""
    B = next(a.shape[0] for a in (numeric, indicator_ids, emb) if a is not None)
    pieces, where = {}, {}
    i_num = i_ind = i_emb = 0
    for name, kind, size in columns:
        if kind == "numeric":
            pieces[name] = numeric[:, i_num:i_num + 1].astype(np.float32)
            i_num += 1
        elif kind == "indicator":
            ids = indicator_ids[:, i_ind]
            pieces[name] = (ids[:, None] == np.arange(size)[None, :]).astype(np.float32)
            i_ind += 1
        elif kind == "embedding":
            pieces[name] = emb[:, i_emb:i_emb + size].astype(np.float32)
            i_emb += size
        else:
            raise ValueError(kind)
    out, c = [], 0
    for name in sorted(pieces):
        out.append(pieces[name])
        where[name] = (c, c + pieces[name].shape[1])
        c += pieces[name].shape[1]
    return np.concatenate(out, axis=1) if out else np.zeros((B, 0), np.float32), where


# ----------------------------------------------------------------------------
# counter-hash table initialisation (cfg4: tables too large to come from the host)
# ----------------------------------------------------------------------------
_M64 = np.uint64(0xFFFFFFFFFFFFFFFF)


def _mix64(x):
    """splitmix64 finaliser on uint64 arrays (wrap-around arithmetic)."""
    with np.errstate(over="ignore"):
        x = (x + np.uint64(0x9E3779B97F4A7C15)) & _M64
        x = ((x ^ (x >> np.uint64(30))) * np.uint64(0xBF58476D1CE4E5B9)) & _M64
        x = ((x ^ (x >> np.uint64(27))) * np.uint64(0x94D049BB133111EB)) & _M64
        return x ^ (x >> np.uint64(31))


def counter_rows(global_rows_, K, seed=1236, sd=None):
    """Rows of a table filled by `dir_table_init_counter`: value(row, k) = (sum of the four 16-bit fields of
    mix64(seed ^ mix64(row * 64 + k)) - 131070) * (sd / 37837.2272), one fp32 multiply -- bit-exact on any host.
    This initialiser is defined by THIS repo (BASELINE cfg4: "initialised on-device by a counter-based RNG");
    the reference's own is TF's truncated_normal(0, 1/sqrt(K)), models/DeepFM/deepFM.py:385-390 [TF]."""
    sd = np.float32(1.0 / np.sqrt(K) if sd is None else sd)
    scale = np.float32(sd / np.float32(37837.2272))
    rows = np.asarray(global_rows_, dtype=np.uint64).reshape(-1, 1)
    with np.errstate(over="ignore"):
        ctr = rows * np.uint64(64) + np.arange(K, dtype=np.uint64)[None, :]
    h = _mix64(np.uint64(seed) ^ _mix64(ctr))
    m = np.uint64(0xFFFF)
    s = ((h & m) + ((h >> np.uint64(16)) & m) + ((h >> np.uint64(32)) & m) + (h >> np.uint64(48))).astype(np.int64)
    return ((s - 131070).astype(np.float32) * scale).astype(np.float32)
