"""Front end: raw features -> (feature_index, feature_value), the step in front of the lookup.

The reference declares its inputs as TF feature columns and lets TensorFlow resolve them per batch
(models/DeepCrossNetwork/train.py:57-100 `build_model_columns`, models/ESMM/train.py:55-100; csv `input_fn`
models/DeepCrossNetwork/train.py:127-156).  This module mirrors the column constructors the reference calls --
same names, same arguments -- and resolves a decoded batch on the device with the kernels of csrc/front.cu:

    numeric_column(key)                                        value -> feature_value of a one-row field
    categorical_column_with_hash_bucket(key, hash_bucket_size) Fingerprint64(s) mod hash_bucket_size
    categorical_column_with_vocabulary_list(key, vocabulary_list, default_value=-1)
    bucketized_column(numeric_column, boundaries)              number of boundaries <= value

`FeatureFrontEnd(columns)` fixes the field order (the order given, like `column_names` in deepFM.py:325-328),
reports `field_size` / `rows_per_field` for `EmbeddingFM`, and `encode(features)` turns a dict of per-column
host arrays (what `parse_csv` yields) into the resolved pair on the device.  Strings cross PCIe once, as one
byte buffer + offsets per column.  There is no CPU path: the hashing runs in libdir_b200.so on the GPU.
"""
import csv
import ctypes
import io
from typing import Dict, List, Sequence

import numpy as np
import torch

from . import _lib
from ._lib import check, ptr


class _Column:
    kind = ""

    def __init__(self, key):
        if not isinstance(key, str) or not key:
            raise ValueError("key must be a non-empty string")
        self.key = key
        self.name = key


class NumericColumn(_Column):
    """tf.feature_column.numeric_column(key): a one-row field whose feature_value carries the number
    (weighted-column semantics, dataset/SequenceTensorFlowDataset/test4.py:50-55)."""
    kind, num_buckets = "numeric", 1


class HashBucketColumn(_Column):
    """tf.feature_column.categorical_column_with_hash_bucket(key, hash_bucket_size) (train.py:84-86)."""
    kind = "hash"

    def __init__(self, key, hash_bucket_size):
        super().__init__(key)
        if hash_bucket_size is None or int(hash_bucket_size) < 1:            # [TF] raises the same way
            raise ValueError("hash_bucket_size must be at least 1. hash_bucket_size: %r, key: %s" % (hash_bucket_size, key))
        self.num_buckets = int(hash_bucket_size)


class VocabularyListColumn(_Column):
    """tf.feature_column.categorical_column_with_vocabulary_list(key, vocabulary_list, default_value=-1)
    (train.py:63-81): the index in the list; strings outside it get default_value (-1 = pruned by the lookup)."""
    kind = "vocab"

    def __init__(self, key, vocabulary_list, default_value=-1):
        super().__init__(key)
        vocab = [v.encode("utf-8") if isinstance(v, str) else bytes(v) for v in vocabulary_list]
        if not vocab:
            raise ValueError("vocabulary_list %r must be non-empty, column_name: %s" % (vocabulary_list, key))
        if len(set(vocab)) != len(vocab):
            raise ValueError("Duplicate keys in vocabulary_list: %r, column_name: %s" % (vocabulary_list, key))
        if not -1 <= int(default_value) < len(vocab):
            raise ValueError("default_value must be -1 or an index into vocabulary_list")
        self.vocabulary, self.default_value, self.num_buckets = vocab, int(default_value), len(vocab)


class BucketizedColumn(_Column):
    """tf.feature_column.bucketized_column(source_column, boundaries) (DeepCrossNetwork.py:58)."""
    kind = "bucket"

    def __init__(self, source_column, boundaries):
        if not isinstance(source_column, NumericColumn):
            raise ValueError("source_column must be a column generated with numeric_column(). Given: %r" % (source_column,))
        super().__init__(source_column.key)
        b = [float(x) for x in boundaries]
        if not b:
            raise ValueError("boundaries must not be empty.")
        if any(b[i] >= b[i + 1] for i in range(len(b) - 1)):
            raise ValueError("boundaries must be a sorted list.")
        self.name = source_column.key + "_bucketized"
        self.boundaries, self.num_buckets = b, len(b) + 1


def numeric_column(key):
    return NumericColumn(key)


def categorical_column_with_hash_bucket(key, hash_bucket_size):
    return HashBucketColumn(key, hash_bucket_size)


def categorical_column_with_vocabulary_list(key, vocabulary_list, default_value=-1):
    return VocabularyListColumn(key, vocabulary_list, default_value)


def bucketized_column(source_column, boundaries):
    return BucketizedColumn(source_column, boundaries)


def pack_strings(values):
    """A column of strings / bytes / ints -> (uint8 bytes, int64 offsets[n+1]) host arrays.  Integers become their
    decimal string ([TF] hashes integer features through as_string)."""
    enc = []
    for v in values:
        if isinstance(v, (bytes, bytearray, np.bytes_)):
            enc.append(bytes(v))
        elif isinstance(v, (int, np.integer)):
            enc.append(str(int(v)).encode("ascii"))
        else:
            enc.append(str(v).encode("utf-8"))
    offsets = np.zeros(len(enc) + 1, dtype=np.int64)
    if enc:
        np.cumsum([len(e) for e in enc], out=offsets[1:])
    data = np.frombuffer(b"".join(enc), dtype=np.uint8).copy() if offsets[-1] else np.zeros(0, np.uint8)
    return data, offsets


def fingerprint64(value: bytes) -> int:
    """FarmHash Fingerprint64 of one host string, by the library's own host entry point."""
    buf = ctypes.create_string_buffer(bytes(value), max(len(value), 1))
    return int(_lib.lib().dir_fingerprint64_host(ctypes.cast(buf, ctypes.c_void_p), len(value)))


def decode_csv(lines: Sequence[str], column_names: Sequence[str], record_defaults: Sequence[Sequence]) -> Dict[str, list]:
    """tf.decode_csv + dict(zip(_CSV_COLUMNS, columns)) of the reference's `parse_csv` (train.py:131-137), on the host:
    one list per column; a field's type is its default's (int / float / str); an empty field takes the default; a
    column whose default is [] is required."""
    if len(column_names) != len(record_defaults):
        raise ValueError("column_names and record_defaults must have the same length")
    cols = {n: [] for n in column_names}
    rows = csv.reader(io.StringIO("\n".join(l.rstrip("\r\n") for l in lines)), skipinitialspace=False)
    for ln, rec in enumerate(rows):
        if not rec:
            continue
        if len(rec) != len(column_names):
            raise ValueError("Expect %d fields but have %d in record %d" % (len(column_names), len(rec), ln))
        for name, field, dflt in zip(column_names, rec, record_defaults):
            if field == "":
                if len(dflt) == 0:
                    raise ValueError("Field %s is required but missing in record %d!" % (name, ln))
                cols[name].append(dflt[0])
                continue
            kind = type(dflt[0]) if len(dflt) else str
            try:
                cols[name].append(kind(field) if kind in (int, float) else field)
            except ValueError:
                raise ValueError("Field %s in record %d is not a valid %s: %s" % (name, ln, kind.__name__, field))
    return cols


class FeatureFrontEnd(torch.nn.Module):
    """columns (in field order) -> resolved inputs of `EmbeddingFM` / `ShardedEmbeddingFM`.

    field_size = len(columns); rows_per_field = each column's bucket count (1 for a numeric column).
    encode(features: {key: sequence}) -> feature_index[B, F] int64, feature_value[B, F] fp32 on the device.
    """

    def __init__(self, columns: Sequence[_Column], device="cuda"):
        super().__init__()
        if not columns:
            raise ValueError("empty columns.")                       # deepFM.py:104-105
        names = [c.name for c in columns]
        if len(set(names)) != len(names):
            raise ValueError("column names must be unique")
        for c in columns:
            if not isinstance(c, _Column):
                raise ValueError("Items of feature_columns must be columns of this module. Given: %r" % (c,))
        self.columns = list(columns)
        self.field_size = len(columns)
        self.rows_per_field = [int(c.num_buckets) for c in columns]
        dev = torch.device(device)
        self._dev = dev
        self._tables = {}
        for f, c in enumerate(self.columns):
            if c.kind == "vocab":
                fps = np.asarray([fingerprint64(v) for v in c.vocabulary], dtype=np.uint64)
                if len(np.unique(fps)) != len(fps):
                    raise ValueError("vocabulary of column %s has a Fingerprint64 collision" % c.key)
                order = np.argsort(fps, kind="stable")
                self.register_buffer("vocab_fp_%d" % f, torch.as_tensor(fps[order].view(np.int64)).to(dev))
                self.register_buffer("vocab_ix_%d" % f, torch.as_tensor(order.astype(np.int64)).to(dev))
            elif c.kind == "bucket":
                self.register_buffer("bounds_%d" % f, torch.tensor(c.boundaries, dtype=torch.float32, device=dev))

    @torch.no_grad()
    def encode(self, features: Dict[str, Sequence]):
        if not isinstance(features, dict):
            raise ValueError("features should be a dictionary of `Tensor`s. Given type: {}".format(type(features)))  # deepFM.py:159-161
        if self._dev.type != "cuda":
            raise ValueError("FeatureFrontEnd needs a CUDA device: the front end has no CPU path")
        F = self.field_size
        sizes = set()
        for c in self.columns:
            if c.key not in features:
                raise ValueError("feature %r is missing" % c.key)
            sizes.add(len(features[c.key]))
        if len(sizes) != 1:
            raise ValueError("every feature must have the same number of values")
        B = sizes.pop()
        dev = self._dev
        idx = torch.zeros((B, F), dtype=torch.int64, device=dev)
        val = torch.ones((B, F), dtype=torch.float32, device=dev)
        if B == 0:
            return idx, val
        L = _lib.lib()
        st = torch.cuda.current_stream().cuda_stream
        keep = []                                        # device buffers must outlive the enqueued kernels
        for f, c in enumerate(self.columns):
            col = features[c.key]
            out = idx.data_ptr() + 8 * f
            if c.kind == "numeric":
                val[:, f] = torch.as_tensor(np.asarray(col, dtype=np.float32)).to(dev)
            elif c.kind == "bucket":
                v = torch.as_tensor(np.asarray(col, dtype=np.float32)).to(dev)
                keep.append(v)
                b = getattr(self, "bounds_%d" % f)
                check(L.dir_bucketize(ptr(v), B, 1, ptr(b), b.numel(), out, F, st), "dir_bucketize")
            else:
                data, offsets = pack_strings(col)
                d = torch.as_tensor(data).to(dev) if data.size else torch.zeros(1, dtype=torch.uint8, device=dev)
                o = torch.as_tensor(offsets).to(dev)
                keep += [d, o]
                if c.kind == "hash":
                    check(L.dir_hash_bucket(ptr(d), ptr(o), B, c.num_buckets, out, F, st), "dir_hash_bucket")
                else:
                    fp, ix = getattr(self, "vocab_fp_%d" % f), getattr(self, "vocab_ix_%d" % f)
                    check(L.dir_vocabulary_lookup(ptr(d), ptr(o), B, ptr(fp), ptr(ix), fp.numel(), c.default_value,
                                                  out, F, st), "dir_vocabulary_lookup")
        torch.cuda.current_stream().synchronize()         # `keep` may be released: the kernels have run
        return idx, val
