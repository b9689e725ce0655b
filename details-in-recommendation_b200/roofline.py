"""ALGORITHMIC bytes of each kernel of the path.

These are the only byte counts `bench.py`'s `roofline.achieved` may use: sort passes, `S[B,K]`,
workspaces and cache re-reads are excluded by definition.  I = 8 (int64 id), V = 4 (fp32 value, 0
when feature_value is None), R = 4K (one row), E = B*F*R when [B,F,K] embeddings / upstream
gradients cross the layer boundary (a DNN or cross consumer is attached) else 0, U = distinct
rows a batch touches, c = 4 for Adagrad (read+write row, read+write accumulator), 2 for SGD.
"""
import json
import os

I_BYTES = 8
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def embed_fwd_bytes(B, F, K, weighted=True, emit=True):
    """lookup + first order + FM forward: ids, values, K-row, first-order weight per lookup;
    write first + fm per sample; + emitted embeddings."""
    V, R = (4 if weighted else 0), 4 * K
    return B * F * (I_BYTES + V + R + 4) + 8 * B + (B * F * R if emit else 0)


def embed_bwd_bytes(B, F, K, U, weighted=True, upstream=True, optimizer="adagrad"):
    """backward + fused row update: g per sample, upstream u, per lookup id + value + one K-vector,
    per distinct row (row + first-order weight) x c."""
    V, R = (4 if weighted else 0), 4 * K
    c = 4 if optimizer == "adagrad" else 2
    return 4 * B + (B * F * R if upstream else 0) + B * F * (I_BYTES + V + R) + U * (R + 4) * c


def cross_fwd_bytes(B, d, L):
    return 2 * B * d * 4 + 2 * L * d * 4


def cross_bwd_bytes(B, d, L):
    return 3 * B * d * 4 + 4 * L * d * 4


def measured_peaks():
    """(hbm GB/s, source): MEASURED_PEAKS.json (driver-written) else the profiling recipe's fallback."""
    path = os.path.join(ROOT, "MEASURED_PEAKS.json")
    try:
        with open(path) as f:
            return float(json.load(f)["hbm_gbs"]), "measured"
    except (OSError, KeyError, ValueError):
        return 6650.0, "fallback"
