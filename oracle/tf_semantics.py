"""TEST INFRASTRUCTURE -- [TF] rules the reference *calls* but does not contain.

PARITY UNPINNED: TensorFlow 1.x is not in /root/reference, is not installed in
this image and has no build for Python 3.12, and the reference holds no golden
vectors for this path.  Every rule below is the TF 1.x
behaviour as recalled from its sources; they live in this one module so a
correction lands in one place.  Only tests/, __graft_entry__.smoke() and
bench.py's cpu_baseline / --impl reference legs may import anything under
oracle/.

Each rule names the reference call site that reaches it.
"""
import numpy as np

# -- models/DeepFM/deepFM.py:387 column._get_dense_tensor on an embedding_column
# [TF] _safe_embedding_lookup_sparse: ids < 0 are pruned; when weights are
# given, entries with weight <= 0 are pruned; a row left empty yields zeros.
COMBINERS = ("sum", "mean", "sqrtn")
# default combiner of tf.feature_column.embedding_column
EMBEDDING_COLUMN_DEFAULT_COMBINER = "mean"
# models/DeepFM/deepFM.py:255 linear_model(sparse_combiner='sum')
LINEAR_DEFAULT_COMBINER = "sum"
# tf.train.AdagradOptimizer(initial_accumulator_value=0.1), no epsilon
ADAGRAD_INITIAL_ACCUMULATOR = 0.1


def keep_mask(idx, val):
    """Which (sample, field) lookups survive _safe_embedding_lookup_sparse."""
    keep = idx >= 0
    if val is not None:
        keep = keep & (val > 0)
    return keep


def effective_value(idx, val, combiner, dtype):
    """Scale applied to the gathered row for ONE id per (sample, field).

    sum:   w * e            mean: w * e / w = e        sqrtn: w * e / sqrt(w^2) = e
    (weights are > 0 after pruning, so mean and sqrtn reduce to a plain gather).
    Pruned lookups get scale 0 (zero vector, no gradient, row not touched).
    """
    if combiner not in COMBINERS:
        raise ValueError("combiner must be one of %r" % (COMBINERS,))
    keep = keep_mask(idx, val)
    if val is None or combiner != "sum":
        eff = np.ones(idx.shape, dtype=dtype)
    else:
        eff = val.astype(dtype)
    return np.where(keep, eff, dtype(0)), keep


def truncated_normal(rng, shape, stddev, dtype=np.float32):
    """tf.truncated_normal_initializer: resample beyond 2 sigma."""
    out = rng.standard_normal(shape)
    bad = np.abs(out) > 2.0
    while bad.any():
        out[bad] = rng.standard_normal(int(bad.sum()))
        bad = np.abs(out) > 2.0
    return (out * stddev).astype(dtype)


def clip_by_norm(values, clip_norm):
    """tf.clip_by_norm (models/DeepCrossNetwork/DeepCrossNetwork.py:284).

    For an IndexedSlices gradient TF clips over `.values`, duplicates not
    merged; `values` here is whatever tensor holds those values.
    t * clip_norm / max(l2norm(t), clip_norm)
    """
    dt = values.dtype.type
    l2 = np.sqrt(np.sum(values * values, dtype=values.dtype))
    return values * (dt(clip_norm) / max(l2, dt(clip_norm)))


def clip_indexed_slices_per_column(rows, values, field_offset, n_rows, clip_norm):
    """clip_by_norm as the reference applies it to the embedding gradients (DeepCrossNetwork.py:282-289): ONE
    VARIABLE PER COLUMN (deepFM.py:385-390; tf.feature_column.input_layer creates one per embedding_column), and
    [TF] embedding_lookup_sparse de-duplicates ids before the gather, so column f's gradient is an IndexedSlices
    whose `values` are the de-duplicated row sums of that column; each is clipped on its own.
    rows[U] global rows (ascending), values[U, ...] their sums -> the clipped values."""
    out = values.copy()
    off = list(field_offset) + [n_rows]
    for f in range(len(field_offset)):
        m = (rows >= off[f]) & (rows < off[f + 1])
        if m.any():
            out[m] = clip_by_norm(values[m], clip_norm)
    return out


def sigmoid_cross_entropy_with_logits(labels, logits):
    """max(x,0) - x*z + log1p(exp(-|x|))  (TF's stable form)."""
    x, z = logits, labels
    return np.maximum(x, 0) - x * z + np.log1p(np.exp(-np.abs(x)))


def sigmoid(x):
    return 1.0 / (1.0 + np.exp(-x))
