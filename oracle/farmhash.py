"""TEST INFRASTRUCTURE -- CPU restatement of the raw-feature -> index step in front of the hot path.

The reference builds its categorical columns with
    tf.feature_column.categorical_column_with_hash_bucket('occupation', hash_bucket_size=1000)
    tf.feature_column.categorical_column_with_vocabulary_list('education', [...])
(models/DeepCrossNetwork/train.py:57-100, models/ESMM/train.py:65-90) and names `bucketized_column` among the
columns it accepts (models/DeepCrossNetwork/DeepCrossNetwork.py:58).  The arithmetic lives in TensorFlow, which is not
in /root/reference and cannot be installed here; what those columns compute is published:

  hash bucket   [TF] string_to_hash_bucket_fast(s, n) = Fingerprint64(s) mod n, Fingerprint64 = FarmHash's
                farmhashna::Hash64 (FarmHash 1.1, Google, MIT licence) -- restated below from the published algorithm
  vocabulary    index of the string in the list, default_value (-1) when absent (num_oov_buckets = 0)
  bucketize     [TF] Bucketize: number of boundaries <= value (std::upper_bound over the sorted boundaries)

PARITY UNPINNED at the TensorFlow boundary like the rest of oracle/: pinned here by the identities the algorithm
implies (Fingerprint64(b"") == k2), by the one published example of the op (tf.strings.to_hash_bucket_fast(["Hello",
"TensorFlow", "2.x"], 3) -> [0, 2, 2], TF API documentation) and by agreement, bit for bit, with the independently
written CUDA kernel over every length class (tests/test_frontend_cpu.py, tests/test_gpu_frontend.py).
Only tests/ may import this module.
"""
import struct

M64 = (1 << 64) - 1
K0 = 0xC3A5C85C97CB3127
K1 = 0xB492B66FBE98F273
K2 = 0x9AE16A3B2F90404F


def _fetch64(s, i):
    return struct.unpack_from("<Q", s, i)[0]


def _fetch32(s, i):
    return struct.unpack_from("<I", s, i)[0]


def _rot(v, sh):
    return v if sh == 0 else ((v >> sh) | (v << (64 - sh))) & M64


def _shift_mix(v):
    return v ^ (v >> 47)


def _hash_len16(u, v, mul):
    a = ((u ^ v) * mul) & M64
    a ^= a >> 47
    b = ((v ^ a) * mul) & M64
    b ^= b >> 47
    return (b * mul) & M64


def _hash_len_0_to_16(s):
    n = len(s)
    if n >= 8:
        mul = (K2 + n * 2) & M64
        a = (_fetch64(s, 0) + K2) & M64
        b = _fetch64(s, n - 8)
        c = (_rot(b, 37) * mul + a) & M64
        d = ((_rot(a, 25) + b) * mul) & M64
        return _hash_len16(c, d, mul)
    if n >= 4:
        mul = (K2 + n * 2) & M64
        a = _fetch32(s, 0)
        return _hash_len16((n + (a << 3)) & M64, _fetch32(s, n - 4), mul)
    if n > 0:
        a, b, c = s[0], s[n >> 1], s[n - 1]
        y = (a + (b << 8)) & 0xFFFFFFFF
        z = (n + (c << 2)) & 0xFFFFFFFF
        return (_shift_mix(((y * K2) & M64) ^ ((z * K0) & M64)) * K2) & M64
    return K2


def _hash_len_17_to_32(s):
    n = len(s)
    mul = (K2 + n * 2) & M64
    a = (_fetch64(s, 0) * K1) & M64
    b = _fetch64(s, 8)
    c = (_fetch64(s, n - 8) * mul) & M64
    d = (_fetch64(s, n - 16) * K2) & M64
    return _hash_len16((_rot((a + b) & M64, 43) + _rot(c, 30) + d) & M64,
                       (a + _rot((b + K2) & M64, 18) + c) & M64, mul)


def _weak32(w, x, y, z, a, b):
    a = (a + w) & M64
    b = _rot((b + a + z) & M64, 21)
    c = a
    a = (a + x) & M64
    a = (a + y) & M64
    b = (b + _rot(a, 44)) & M64
    return (a + z) & M64, (b + c) & M64


def _weak32_at(s, i, a, b):
    return _weak32(_fetch64(s, i), _fetch64(s, i + 8), _fetch64(s, i + 16), _fetch64(s, i + 24), a, b)


def _hash_len_33_to_64(s):
    n = len(s)
    mul = (K2 + n * 2) & M64
    a = (_fetch64(s, 0) * K2) & M64
    b = _fetch64(s, 8)
    c = (_fetch64(s, n - 8) * mul) & M64
    d = (_fetch64(s, n - 16) * K2) & M64
    y = (_rot((a + b) & M64, 43) + _rot(c, 30) + d) & M64
    z = _hash_len16(y, (a + _rot((b + K2) & M64, 18) + c) & M64, mul)
    e = (_fetch64(s, 16) * mul) & M64
    f = _fetch64(s, 24)
    g = ((y + _fetch64(s, n - 32)) * mul) & M64
    h = ((z + _fetch64(s, n - 24)) * mul) & M64
    return _hash_len16((_rot((e + f) & M64, 43) + _rot(g, 30) + h) & M64,
                       (e + _rot((f + a) & M64, 18) + g) & M64, mul)


def fingerprint64(s: bytes) -> int:
    """farmhash::Fingerprint64 = farmhashna::Hash64 ([TF] Fingerprint64, core/platform/fingerprint.h)."""
    s = bytes(s)
    n = len(s)
    if n <= 16:
        return _hash_len_0_to_16(s)
    if n <= 32:
        return _hash_len_17_to_32(s)
    if n <= 64:
        return _hash_len_33_to_64(s)
    seed = 81
    x = seed
    y = (seed * K1 + 113) & M64
    z = (_shift_mix((y * K2 + 113) & M64) * K2) & M64
    v = (0, 0)
    w = (0, 0)
    x = (x * K2 + _fetch64(s, 0)) & M64
    end = ((n - 1) // 64) * 64
    last64 = end + ((n - 1) & 63) - 63
    p = 0
    while True:
        x = (_rot((x + y + v[0] + _fetch64(s, p + 8)) & M64, 37) * K1) & M64
        y = (_rot((y + v[1] + _fetch64(s, p + 48)) & M64, 42) * K1) & M64
        x ^= w[1]
        y = (y + v[0] + _fetch64(s, p + 40)) & M64
        z = (_rot((z + w[0]) & M64, 33) * K1) & M64
        v = _weak32_at(s, p, (v[1] * K1) & M64, (x + w[0]) & M64)
        w = _weak32_at(s, p + 32, (z + w[1]) & M64, (y + _fetch64(s, p + 16)) & M64)
        z, x = x, z
        p += 64
        if p == end:
            break
    mul = (K1 + ((z & 0xFF) << 1)) & M64
    p = last64
    w = ((w[0] + ((n - 1) & 63)) & M64, w[1])
    v = ((v[0] + w[0]) & M64, v[1])
    w = ((w[0] + v[0]) & M64, w[1])
    x = (_rot((x + y + v[0] + _fetch64(s, p + 8)) & M64, 37) * mul) & M64
    y = (_rot((y + v[1] + _fetch64(s, p + 48)) & M64, 42) * mul) & M64
    x ^= (w[1] * 9) & M64
    y = (y + v[0] * 9 + _fetch64(s, p + 40)) & M64
    z = (_rot((z + w[0]) & M64, 33) * mul) & M64
    v = _weak32_at(s, p, (v[1] * mul) & M64, (x + w[0]) & M64)
    w = _weak32_at(s, p + 32, (z + w[1]) & M64, (y + _fetch64(s, p + 16)) & M64)
    z, x = x, z
    return _hash_len16((_hash_len16(v[0], w[0], mul) + (_shift_mix(y) * K0) + z) & M64,
                       (_hash_len16(v[1], w[1], mul) + x) & M64, mul)


def string_to_hash_bucket_fast(strings, num_buckets):
    """[TF] categorical_column_with_hash_bucket on string input (train.py:84-86): Fingerprint64(s) mod num_buckets.
    Integer inputs are hashed through their decimal string ([TF] as_string), e.g. 17 -> b"17"."""
    out = []
    for s in strings:
        if isinstance(s, (int,)):
            s = str(s)
        if isinstance(s, str):
            s = s.encode("utf-8")
        out.append(fingerprint64(s) % num_buckets)
    return out


def vocabulary_lookup(strings, vocabulary, default_value=-1):
    """[TF] categorical_column_with_vocabulary_list (train.py:63-81): the index in the list, else default_value."""
    index = {}
    for i, v in enumerate(vocabulary):
        v = v.encode("utf-8") if isinstance(v, str) else bytes(v)
        if v in index:
            raise ValueError("duplicate vocabulary entry %r" % v)
        index[v] = i
    return [index.get(s.encode("utf-8") if isinstance(s, str) else bytes(s), default_value) for s in strings]


def bucketize(values, boundaries):
    """[TF] bucketized_column / Bucketize: number of boundaries <= value (upper_bound); boundaries ascending."""
    import bisect
    b = list(boundaries)
    if any(b[i] >= b[i + 1] for i in range(len(b) - 1)):
        raise ValueError("boundaries must be strictly increasing")
    return [bisect.bisect_right(b, float(v)) for v in values]
