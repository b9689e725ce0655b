"""GPU parity of the front end (csrc/front.cu) against oracle/farmhash.py: Fingerprint64 of strings of every length
class incl. empty ones, hash buckets, vocabulary lookups with out-of-vocabulary strings, bucketize edge cases -- all
bit-exact (integer work) -- and the census columns of the reference (models/DeepCrossNetwork/train.py:57-100) end to
end into the DCN-shaped model input."""
import random

import numpy as np
import pytest
import torch

from oracle import farmhash as fh

pytestmark = pytest.mark.gpu


def _strings(seed, n):
    rnd = random.Random(seed)
    lens = list(range(0, 70)) + [127, 128, 129, 500]
    return [bytes(rnd.getrandbits(8) for _ in range(rnd.choice(lens))) for _ in range(n)]


def test_fingerprint_and_hash_bucket_bit_exact(pkg, cuda):
    from dir_b200 import _lib, frontend as fe
    L = _lib.lib()
    strs = _strings(5, 4000) + [b"", b"Hello", b"TensorFlow", b"2.x"]
    data, off = fe.pack_strings(strs)
    d, o = torch.as_tensor(data).cuda(), torch.as_tensor(off).cuda()
    n = len(strs)
    out = torch.empty(n, dtype=torch.int64, device="cuda")
    st = torch.cuda.current_stream().cuda_stream
    _lib.check(L.dir_fingerprint64(d.data_ptr(), o.data_ptr(), n, out.data_ptr(), st), "fp")
    want = np.asarray([fh.fingerprint64(s) for s in strs], dtype=np.uint64)
    assert np.array_equal(out.cpu().numpy().view(np.uint64), want)
    for buckets in (1, 3, 1000, 2 ** 31 - 1, 2 ** 40 + 7):
        _lib.check(L.dir_hash_bucket(d.data_ptr(), o.data_ptr(), n, buckets, out.data_ptr(), 1, st), "hb")
        assert out.cpu().tolist() == fh.string_to_hash_bucket_fast(strs, buckets)
    _lib.check(L.dir_hash_bucket(d.data_ptr(), o.data_ptr(), n, 3, out.data_ptr(), 1, st), "hb")
    assert out[-3:].cpu().tolist() == [0, 2, 2]                    # the published example
    assert L.dir_hash_bucket(d.data_ptr(), o.data_ptr(), n, 0, out.data_ptr(), 1, st) == -22


def test_census_columns_end_to_end(pkg, cuda):
    from dir_b200 import frontend as fe
    rnd = random.Random(9)
    workclass = ["Self-emp-not-inc", "Private", "State-gov", "Federal-gov", "Local-gov", "?", "Self-emp-inc",
                 "Without-pay", "Never-worked"]                                              # train.py:80-83
    relationship = ["Husband", "Not-in-family", "Wife", "Own-child", "Unmarried", "Other-relative"]
    cols = [fe.numeric_column("age"), fe.numeric_column("hours_per_week"),
            fe.categorical_column_with_vocabulary_list("workclass", workclass),
            fe.categorical_column_with_vocabulary_list("relationship", relationship, default_value=1),
            fe.categorical_column_with_hash_bucket("occupation", hash_bucket_size=1000),      # train.py:84-86
            fe.bucketized_column(fe.numeric_column("capital_gain"), [0.5, 1000.0, 5000.0, 99999.0])]
    front = fe.FeatureFrontEnd(cols)
    B = 777
    occ = ["Tech-support", "Craft-repair", "Other-service", "Sales", "Exec-managerial", "Prof-specialty", "?", ""]
    feats = {"age": [rnd.uniform(17, 90) for _ in range(B)], "hours_per_week": [float(rnd.randint(1, 99)) for _ in range(B)],
             "workclass": [rnd.choice(workclass + ["Martian"]) for _ in range(B)],
             "relationship": [rnd.choice(relationship + ["nobody"]) for _ in range(B)],
             "occupation": [rnd.choice(occ) for _ in range(B)],
             "capital_gain": [rnd.choice([0.0, 0.5, 999.99, 1000.0, 4999.0, 5000.0, 99999.0, 1e6, -3.0]) for _ in range(B)]}
    idx, val = front.encode(feats)
    assert idx.shape == (B, 6) and val.shape == (B, 6) and idx.dtype == torch.int64
    idx, val = idx.cpu().numpy(), val.cpu().numpy()
    assert np.array_equal(idx[:, 0], np.zeros(B, np.int64)) and np.array_equal(idx[:, 1], np.zeros(B, np.int64))
    assert np.array_equal(val[:, 0], np.asarray(feats["age"], np.float32))
    assert np.array_equal(val[:, 1], np.asarray(feats["hours_per_week"], np.float32))
    assert idx[:, 2].tolist() == fh.vocabulary_lookup(feats["workclass"], workclass)            # "Martian" -> -1
    assert idx[:, 3].tolist() == fh.vocabulary_lookup(feats["relationship"], relationship, default_value=1)
    assert idx[:, 4].tolist() == fh.string_to_hash_bucket_fast(feats["occupation"], 1000)
    assert idx[:, 5].tolist() == fh.bucketize(np.asarray(feats["capital_gain"], np.float32), [0.5, 1000.0, 5000.0, 99999.0])
    assert (val[:, 2:] == 1.0).all() and (idx[:, 2] == -1).any()
    # the resolved pair feeds the layer: out-of-vocabulary rows (-1) are pruned, like TF's default_value = -1
    layer = pkg.EmbeddingFM(front.field_size, 8, front.rows_per_field).train()
    first, fm, emb = layer(torch.as_tensor(idx).cuda(), torch.as_tensor(val).cuda())
    e = emb.detach().cpu().numpy().reshape(B, 6, 8)
    assert (e[idx[:, 2] == -1, 2] == 0).all()
    tab = layer.table.cpu().numpy()
    off = layer.field_offset.cpu().numpy()
    keep = idx[:, 4] >= 0
    assert np.array_equal(e[keep, 4], tab[off[4] + idx[keep, 4]])
    (first.sum() + fm.sum() + emb.sum()).backward()
    torch.cuda.synchronize()


def test_bucketize_and_empty_batches(pkg, cuda):
    from dir_b200 import frontend as fe
    front = fe.FeatureFrontEnd([fe.bucketized_column(fe.numeric_column("x"), [-1.0, 0.0, 2.5]),
                                fe.categorical_column_with_hash_bucket("s", 7)])
    xs = [-np.inf, -1.0, -0.5, 0.0, 2.4999, 2.5, np.inf, np.nan]
    idx, val = front.encode({"x": xs, "s": ["a"] * len(xs)})
    assert idx[:, 0].cpu().tolist() == fh.bucketize(np.asarray(xs, np.float32), [-1.0, 0.0, 2.5]) == [0, 1, 1, 2, 2, 3, 3, 3]
    assert idx[:, 1].cpu().tolist() == fh.string_to_hash_bucket_fast(["a"] * len(xs), 7)
    idx, val = front.encode({"x": [], "s": []})
    assert idx.shape == (0, 2) and val.shape == (0, 2)
    with pytest.raises(ValueError):
        front.encode({"x": [1.0], "s": []})
    with pytest.raises(ValueError):
        front.encode({"x": [1.0]})
