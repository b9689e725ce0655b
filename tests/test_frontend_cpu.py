"""The front end's restatement (oracle/farmhash.py) and the host side of details-in-recommendation_b200/frontend.py:
known answers of Fingerprint64, the published example of string_to_hash_bucket_fast, the library's host fingerprint
against the restatement over every length class, vocabulary / bucketize rules, csv decoding, and the column
constructors' error behaviour (the reference's call sites: models/DeepCrossNetwork/train.py:57-100, :127-156)."""
import random

import numpy as np
import pytest

from oracle import farmhash as fh


def test_fingerprint64_known_answers():
    assert fh.fingerprint64(b"") == 0x9AE16A3B2F90404F                     # k2: what the algorithm returns for len 0
    # the one published example of the op (TF API docs, tf.strings.to_hash_bucket_fast): three length classes
    assert fh.string_to_hash_bucket_fast(["Hello", "TensorFlow", "2.x"], 3) == [0, 2, 2]
    # integers are hashed through their decimal string
    assert fh.string_to_hash_bucket_fast([17], 1000) == fh.string_to_hash_bucket_fast(["17"], 1000)
    # frozen values of this restatement, one per branch (len 1-3, 4-7, 8-16, 17-32, 33-64, > 64)
    frozen = {b"a": 0xB3454265B6DF75E3, b"hello": 0xB48BE5A931380CE8, b"hello world": 0x588FB7478BD6B01B,
              b"x" * 17: 0x9CE745A1F3812FA5, b"y" * 33: 0x777F79E52D060EFE, b"z" * 65: 0x732F393FA3E7DF35}
    for s, want in frozen.items():
        assert fh.fingerprint64(s) == want


def test_library_host_fingerprint_matches_restatement(pkg):
    from dir_b200 import frontend as fe
    rnd = random.Random(3)
    for n in list(range(0, 140)) + [191, 192, 193, 255, 256, 257, 1000]:
        for _ in range(2):
            s = bytes(rnd.getrandbits(8) for _ in range(n))
            assert fe.fingerprint64(s) == fh.fingerprint64(s), n


def test_vocabulary_and_bucketize_rules():
    vocab = ["Husband", "Not-in-family", "Wife", "Own-child", "Unmarried", "Other-relative"]        # train.py:75-78
    assert fh.vocabulary_lookup(["Wife", "Husband", "nobody", ""], vocab) == [2, 0, -1, -1]
    assert fh.vocabulary_lookup(["nobody"], vocab, default_value=3) == [3]
    with pytest.raises(ValueError):
        fh.vocabulary_lookup(["a"], ["a", "a"])
    b = [18.0, 25.0, 30.0]
    assert fh.bucketize([17.9, 18.0, 24.99, 25.0, 30.0, 99.0, -1e9, float("nan")], b) == [0, 1, 1, 2, 3, 3, 0, 3]
    with pytest.raises(ValueError):
        fh.bucketize([1.0], [2.0, 2.0])


def test_pack_strings_and_decode_csv(pkg):
    from dir_b200 import frontend as fe
    data, off = fe.pack_strings(["ab", b"c", 17, "", "é"])
    assert off.tolist() == [0, 2, 3, 5, 5, 7] and bytes(data) == b"abc17" + "é".encode()
    data, off = fe.pack_strings([])
    assert off.tolist() == [0] and data.size == 0
    cols = ["age", "workclass", "hours", "label"]
    dfl = [[0], [""], [0.0], []]                                  # [] = required (tf.decode_csv)
    out = fe.decode_csv(['39,State-gov,40.5,>50K', '50,,,"<=50K"', ""], cols, dfl)
    assert out == {"age": [39, 50], "workclass": ["State-gov", ""], "hours": [40.5, 0.0], "label": [">50K", "<=50K"]}
    with pytest.raises(ValueError, match="required"):
        fe.decode_csv(["1,a,2.0,"], cols, dfl)
    with pytest.raises(ValueError, match="Expect 4 fields"):
        fe.decode_csv(["1,a"], cols, dfl)
    with pytest.raises(ValueError, match="not a valid int"):
        fe.decode_csv(["x,a,1.0,l"], cols, dfl)


def test_column_constructors_mirror_the_reference_errors(pkg):
    from dir_b200 import frontend as fe
    with pytest.raises(ValueError):
        fe.categorical_column_with_hash_bucket("occupation", 0)
    with pytest.raises(ValueError):
        fe.categorical_column_with_vocabulary_list("education", [])
    with pytest.raises(ValueError):
        fe.categorical_column_with_vocabulary_list("education", ["a", "a"])
    with pytest.raises(ValueError):
        fe.bucketized_column(fe.categorical_column_with_hash_bucket("x", 10), [1.0])
    with pytest.raises(ValueError):
        fe.bucketized_column(fe.numeric_column("age"), [3.0, 2.0])
    with pytest.raises(ValueError, match="empty columns"):
        fe.FeatureFrontEnd([], device="cpu")
    cols = [fe.numeric_column("age"), fe.categorical_column_with_hash_bucket("occupation", 1000),
            fe.categorical_column_with_vocabulary_list("relationship", ["Husband", "Wife"]),
            fe.bucketized_column(fe.numeric_column("hours"), [20.0, 40.0])]
    front = fe.FeatureFrontEnd(cols, device="cpu")                 # construction is host-only plumbing
    assert front.field_size == 4 and front.rows_per_field == [1, 1000, 2, 3]
    with pytest.raises(ValueError, match="no CPU path"):
        front.encode({"age": [1.0], "occupation": ["x"], "relationship": ["Wife"], "hours": [3.0]})
    with pytest.raises(ValueError, match="dictionary"):
        front.encode([1, 2])
