// Front end: raw feature -> index, the step before the lookup.
//
// The reference builds its categorical columns with tf.feature_column.categorical_column_with_hash_bucket /
// categorical_column_with_vocabulary_list (models/DeepCrossNetwork/train.py:57-100, models/ESMM/train.py:65-90)
// and accepts bucketized_column (models/DeepCrossNetwork/DeepCrossNetwork.py:58); TensorFlow evaluates them per
// batch on the host.  Here the decoded batch's strings travel as one byte buffer + offsets and are resolved on the
// device in front of the lookup -- byte / integer work, one thread per value:
//   hash bucket   [TF] string_to_hash_bucket_fast: Fingerprint64(s) mod n, Fingerprint64 = FarmHash's
//                 farmhashna::Hash64 (FarmHash 1.1, published algorithm; Google, MIT licence)
//   vocabulary    Fingerprint64(s) looked up in the sorted fingerprints of the vocabulary (binary search); the host
//                 layer checks the vocabulary for fingerprint collisions when it is built
//   bucketize     [TF] Bucketize: number of boundaries <= value
#include "common.cuh"

namespace dir {

constexpr uint64_t kF0 = 0xc3a5c85c97cb3127ULL;
constexpr uint64_t kF1 = 0xb492b66fbe98f273ULL;
constexpr uint64_t kF2 = 0x9ae16a3b2f90404fULL;

// unaligned little-endian fetches out of the byte buffer
__host__ __device__ __forceinline__ uint64_t fetch64(const uint8_t* p) {
  uint64_t v = 0;
#pragma unroll
  for (int i = 0; i < 8; ++i) v |= (uint64_t)p[i] << (8 * i);
  return v;
}
__host__ __device__ __forceinline__ uint64_t fetch32(const uint8_t* p) {
  uint32_t v = 0;
#pragma unroll
  for (int i = 0; i < 4; ++i) v |= (uint32_t)p[i] << (8 * i);
  return v;
}
__host__ __device__ __forceinline__ uint64_t rot(uint64_t v, int s) { return s == 0 ? v : ((v >> s) | (v << (64 - s))); }
__host__ __device__ __forceinline__ uint64_t shift_mix(uint64_t v) { return v ^ (v >> 47); }
__host__ __device__ __forceinline__ uint64_t hash16(uint64_t u, uint64_t v, uint64_t mul) {
  uint64_t a = (u ^ v) * mul;
  a ^= a >> 47;
  uint64_t b = (v ^ a) * mul;
  b ^= b >> 47;
  return b * mul;
}
struct U2 {
  uint64_t first, second;
};
__host__ __device__ __forceinline__ U2 weak32(uint64_t w, uint64_t x, uint64_t y, uint64_t z, uint64_t a, uint64_t b) {
  a += w;
  b = rot(b + a + z, 21);
  const uint64_t c = a;
  a += x;
  a += y;
  b += rot(a, 44);
  return U2{a + z, b + c};
}
__host__ __device__ __forceinline__ U2 weak32_at(const uint8_t* s, uint64_t a, uint64_t b) {
  return weak32(fetch64(s), fetch64(s + 8), fetch64(s + 16), fetch64(s + 24), a, b);
}

__host__ __device__ uint64_t fingerprint64(const uint8_t* s, int64_t len) {
  if (len <= 16) {
    if (len >= 8) {
      const uint64_t mul = kF2 + (uint64_t)len * 2;
      const uint64_t a = fetch64(s) + kF2;
      const uint64_t b = fetch64(s + len - 8);
      const uint64_t c = rot(b, 37) * mul + a;
      const uint64_t d = (rot(a, 25) + b) * mul;
      return hash16(c, d, mul);
    }
    if (len >= 4) {
      const uint64_t mul = kF2 + (uint64_t)len * 2;
      const uint64_t a = fetch32(s);
      return hash16((uint64_t)len + (a << 3), fetch32(s + len - 4), mul);
    }
    if (len > 0) {
      const uint8_t a = s[0], b = s[len >> 1], c = s[len - 1];
      const uint32_t y = (uint32_t)a + ((uint32_t)b << 8);
      const uint32_t z = (uint32_t)len + ((uint32_t)c << 2);
      return shift_mix((uint64_t)y * kF2 ^ (uint64_t)z * kF0) * kF2;
    }
    return kF2;
  }
  if (len <= 32) {
    const uint64_t mul = kF2 + (uint64_t)len * 2;
    const uint64_t a = fetch64(s) * kF1;
    const uint64_t b = fetch64(s + 8);
    const uint64_t c = fetch64(s + len - 8) * mul;
    const uint64_t d = fetch64(s + len - 16) * kF2;
    return hash16(rot(a + b, 43) + rot(c, 30) + d, a + rot(b + kF2, 18) + c, mul);
  }
  if (len <= 64) {
    const uint64_t mul = kF2 + (uint64_t)len * 2;
    const uint64_t a = fetch64(s) * kF2;
    const uint64_t b = fetch64(s + 8);
    const uint64_t c = fetch64(s + len - 8) * mul;
    const uint64_t d = fetch64(s + len - 16) * kF2;
    const uint64_t y = rot(a + b, 43) + rot(c, 30) + d;
    const uint64_t z = hash16(y, a + rot(b + kF2, 18) + c, mul);
    const uint64_t e = fetch64(s + 16) * mul;
    const uint64_t f = fetch64(s + 24);
    const uint64_t g = (y + fetch64(s + len - 32)) * mul;
    const uint64_t h = (z + fetch64(s + len - 24)) * mul;
    return hash16(rot(e + f, 43) + rot(g, 30) + h, e + rot(f + a, 18) + g, mul);
  }
  const uint64_t seed = 81;
  uint64_t x = seed;
  uint64_t y = seed * kF1 + 113;
  uint64_t z = shift_mix(y * kF2 + 113) * kF2;
  U2 v{0, 0}, w{0, 0};
  x = x * kF2 + fetch64(s);
  const uint8_t* end = s + ((len - 1) / 64) * 64;
  const uint8_t* last64 = end + ((len - 1) & 63) - 63;
  do {
    x = rot(x + y + v.first + fetch64(s + 8), 37) * kF1;
    y = rot(y + v.second + fetch64(s + 48), 42) * kF1;
    x ^= w.second;
    y += v.first + fetch64(s + 40);
    z = rot(z + w.first, 33) * kF1;
    v = weak32_at(s, v.second * kF1, x + w.first);
    w = weak32_at(s + 32, z + w.second, y + fetch64(s + 16));
    const uint64_t t = z;
    z = x;
    x = t;
    s += 64;
  } while (s != end);
  const uint64_t mul = kF1 + ((z & 0xff) << 1);
  s = last64;
  w.first += (uint64_t)((len - 1) & 63);
  v.first += w.first;
  w.first += v.first;
  x = rot(x + y + v.first + fetch64(s + 8), 37) * mul;
  y = rot(y + v.second + fetch64(s + 48), 42) * mul;
  x ^= w.second * 9;
  y += v.first * 9 + fetch64(s + 40);
  z = rot(z + w.first, 33) * mul;
  v = weak32_at(s, v.second * mul, x + w.first);
  w = weak32_at(s + 32, z + w.second, y + fetch64(s + 16));
  const uint64_t t = z;
  z = x;
  x = t;
  return hash16(hash16(v.first, w.first, mul) + shift_mix(y) * kF0 + z, hash16(v.second, w.second, mul) + x, mul);
}

// mode 0: out = fingerprint (as int64 bits); mode 1: fingerprint mod num_buckets;
// mode 2: index of the fingerprint in the sorted vocabulary table, else default_value
__global__ void __launch_bounds__(256)
strings_kernel(const uint8_t* __restrict__ bytes, const int64_t* __restrict__ offsets, int64_t n, int mode,
               uint64_t num_buckets, const uint64_t* __restrict__ vocab_fp, const int64_t* __restrict__ vocab_index,
               int64_t n_vocab, int64_t default_value, int64_t* __restrict__ out, int64_t out_stride) {
  const int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= n) return;
  const int64_t b = __ldg(offsets + i), e = __ldg(offsets + i + 1);
  const uint64_t fp = fingerprint64(bytes + b, e - b);
  int64_t r;
  if (mode == 0) {
    r = (int64_t)fp;
  } else if (mode == 1) {
    r = (int64_t)(fp % num_buckets);
  } else {
    int64_t lo = 0, hi = n_vocab;
    while (lo < hi) {
      const int64_t mid = (lo + hi) >> 1;
      if (__ldg(vocab_fp + mid) < fp) lo = mid + 1; else hi = mid;
    }
    r = (lo < n_vocab && __ldg(vocab_fp + lo) == fp) ? __ldg(vocab_index + lo) : default_value;
  }
  out[i * out_stride] = r;
}

__global__ void __launch_bounds__(256)
bucketize_kernel(const float* __restrict__ values, int64_t n, int64_t value_stride,
                 const float* __restrict__ boundaries, int n_boundaries, int64_t* __restrict__ out,
                 int64_t out_stride) {
  const int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= n) return;
  const float v = __ldg(values + i * value_stride);
  int lo = 0, hi = n_boundaries;  // first boundary > v  (std::upper_bound)
  while (lo < hi) {
    const int mid = (lo + hi) >> 1;
    if (!(v < __ldg(boundaries + mid))) lo = mid + 1; else hi = mid;  // NaN sorts last, as upper_bound has it
  }
  out[i * out_stride] = lo;
}

static int strings_call(const char* what, int mode, const uint8_t* bytes, const int64_t* offsets, int64_t n,
                        uint64_t num_buckets, const uint64_t* vocab_fp, const int64_t* vocab_index, int64_t n_vocab,
                        int64_t default_value, int64_t* out, int64_t out_stride, dir_stream_t stream) {
  if (n < 0 || out_stride <= 0) return fail(DIR_EINVAL, "%s: n >= 0 and out_stride > 0 required", what);
  if (n == 0) return 0;
  if (!offsets || !out) return fail(DIR_EINVAL, "%s: offsets and out are required", what);
  if ((n + 255) / 256 > 0x7fffffffLL) return fail(DIR_EINVAL, "%s: n too large", what);
  strings_kernel<<<(unsigned)((n + 255) / 256), 256, 0, static_cast<cudaStream_t>(stream)>>>(
      bytes, offsets, n, mode, num_buckets, vocab_fp, vocab_index, n_vocab, default_value, out, out_stride);
  return launched(what);
}

}  // namespace dir

/* the same function on the host, one string: the host layer fingerprints a vocabulary with it when it builds the
 * sorted table dir_vocabulary_lookup searches (bytes is a HOST pointer here) */
extern "C" uint64_t dir_fingerprint64_host(const uint8_t* bytes_host, int64_t len) {
  return len < 0 ? 0 : dir::fingerprint64(bytes_host, len);
}

extern "C" int dir_fingerprint64(const uint8_t* bytes, const int64_t* offsets, int64_t n, int64_t* out,
                                 dir_stream_t stream) {
  return dir::strings_call("fingerprint64", 0, bytes, offsets, n, 1, nullptr, nullptr, 0, 0, out, 1, stream);
}

extern "C" int dir_hash_bucket(const uint8_t* bytes, const int64_t* offsets, int64_t n, int64_t num_buckets,
                               int64_t* out, int64_t out_stride, dir_stream_t stream) {
  if (num_buckets <= 0) return dir::fail(DIR_EINVAL, "hash_bucket: hash_bucket_size must be > 0");
  return dir::strings_call("hash_bucket", 1, bytes, offsets, n, (uint64_t)num_buckets, nullptr, nullptr, 0, 0, out,
                           out_stride, stream);
}

extern "C" int dir_vocabulary_lookup(const uint8_t* bytes, const int64_t* offsets, int64_t n,
                                     const uint64_t* vocab_fingerprints, const int64_t* vocab_index, int64_t n_vocab,
                                     int64_t default_value, int64_t* out, int64_t out_stride, dir_stream_t stream) {
  if (n_vocab < 0 || (n_vocab > 0 && (!vocab_fingerprints || !vocab_index)))
    return dir::fail(DIR_EINVAL, "vocabulary_lookup: the vocabulary table is required");
  return dir::strings_call("vocabulary_lookup", 2, bytes, offsets, n, 1, vocab_fingerprints, vocab_index, n_vocab,
                           default_value, out, out_stride, stream);
}

extern "C" int dir_bucketize(const float* values, int64_t n, int64_t value_stride, const float* boundaries,
                             int n_boundaries, int64_t* out, int64_t out_stride, dir_stream_t stream) {
  using namespace dir;
  if (n < 0 || n_boundaries < 0 || out_stride <= 0 || value_stride <= 0)
    return fail(DIR_EINVAL, "bucketize: n, n_boundaries >= 0 and positive strides required");
  if (n == 0) return 0;
  if (!values || !out || (n_boundaries > 0 && !boundaries)) return fail(DIR_EINVAL, "bucketize: null pointer");
  bucketize_kernel<<<(unsigned)((n + 255) / 256), 256, 0, static_cast<cudaStream_t>(stream)>>>(
      values, n, value_stride, boundaries, n_boundaries, out, out_stride);
  return launched("bucketize");
}
