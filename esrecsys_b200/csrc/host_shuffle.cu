// Window shuffle of the host input pipeline (HOST code; SURVEY.md 8(f) N3: "shuffle buffer; pinned-memory
// double-buffered H2D").  The reference fills a list with `shuffle_size` (i, j, count) tuples and np.random.shuffle()s it
// (get_shuffled_items, wikipedia/cooccurrence_matrix.py:80-87; 5 000 000 in train_cooccurence.py:49); in NumPy that is a
// 5M-element permutation plus three fancy-index gathers per window, 10 M pairs/s -- two orders of magnitude below the
// CUDA step.  Here the permutation is never materialised: pi = a keyed bijection of the next power of two (four rounds
// of multiply / xor-shift / add), cycle-walked into [0, n) -- fixed by `seed` -- and the batch [k0, k0 + m) of the
// shuffled window is gathered by `threads` host threads straight into its destination (the pinned (2,B) | (B,) block
// of the batch ring).
#include <stdint.h>

#include <thread>
#include <vector>

#include "esr.h"

namespace {

inline uint64_t splitmix64(uint64_t& s) {
  uint64_t z = (s += 0x9E3779B97F4A7C15ull);
  z = (z ^ (z >> 30)) * 0xBF58476D1CE4E5B9ull;
  z = (z ^ (z >> 27)) * 0x94D049BB133111EBull;
  return z ^ (z >> 31);
}

// Keyed bijection of [0, 2^bits), bits = ceil(log2 n): rounds of (multiply by an odd key, xor-shift right, add a key),
// each one invertible modulo 2^bits; cycle-walked into [0, n) (domain < 2n: fewer than two steps expected).
struct Feistel {
  int bits, shift;
  uint64_t mask, n;
  uint64_t mul[4], add[4];

  Feistel(uint64_t n_, uint64_t seed) : n(n_) {
    bits = 1;
    while (bits < 62 && ((uint64_t)1 << bits) < n_) ++bits;
    shift = bits / 2 > 0 ? bits / 2 : 1;
    mask = ((uint64_t)1 << bits) - 1;
    uint64_t s = seed;
    for (int r = 0; r < 4; ++r) {
      mul[r] = splitmix64(s) | 1;  // odd: invertible modulo 2^bits
      add[r] = splitmix64(s);
    }
  }
  inline uint64_t encrypt(uint64_t v) const {
#pragma unroll
    for (int r = 0; r < 4; ++r) {
      v = (v * mul[r]) & mask;
      v ^= v >> shift;
      v = (v + add[r]) & mask;
    }
    return v;
  }
  inline uint64_t operator()(uint64_t k) const {
    uint64_t v = encrypt(k);
    while (v >= n) v = encrypt(v);
    return v;
  }
};

void gather_range(const int32_t* i, const int32_t* j, const float* c, const Feistel& pi, int64_t k0, int64_t k1, int32_t* di,
                  int32_t* dj, float* dc, int64_t dst0) {
  constexpr int kAhead = 16;  // source indices computed (and their cache lines requested) this many elements ahead
  uint64_t idx[kAhead];
  const int64_t m = k1 - k0;
  for (int64_t a = 0; a < m && a < kAhead; ++a) {
    idx[a] = pi((uint64_t)(k0 + a));
    __builtin_prefetch(i + idx[a]);
    __builtin_prefetch(j + idx[a]);
    __builtin_prefetch(c + idx[a]);
  }
  for (int64_t a = 0; a < m; ++a) {
    const uint64_t s = idx[a % kAhead];
    if (a + kAhead < m) {
      const uint64_t nx = pi((uint64_t)(k0 + a + kAhead));
      idx[a % kAhead] = nx;
      __builtin_prefetch(i + nx);
      __builtin_prefetch(j + nx);
      __builtin_prefetch(c + nx);
    }
    di[dst0 + a] = i[s];
    dj[dst0 + a] = j[s];
    dc[dst0 + a] = c[s];
  }
}

}  // namespace

extern "C" int esr_host_shuffle_gather(const int32_t* i, const int32_t* j, const float* c, int64_t n, uint64_t seed,
                                       int64_t k0, int64_t m, int32_t* dst_i, int32_t* dst_j, float* dst_c,
                                       int32_t threads) {
  if (n < 0 || m < 0 || k0 < 0 || k0 + m > n) return ESR_EINVAL;
  if (m == 0) return ESR_OK;
  if (!i || !j || !c || !dst_i || !dst_j || !dst_c) return ESR_EINVAL;
  const Feistel pi((uint64_t)n, seed);
  int t = threads < 1 ? 1 : (threads > 64 ? 64 : threads);
  if (m < 4096 * (int64_t)t) t = (int)(m / 4096 > 0 ? m / 4096 : 1);
  if (t == 1) {
    gather_range(i, j, c, pi, k0, k0 + m, dst_i, dst_j, dst_c, 0);
    return ESR_OK;
  }
  std::vector<std::thread> pool;
  pool.reserve(t - 1);
  const int64_t per = (m + t - 1) / t;
  for (int w = 1; w < t; ++w) {
    const int64_t a = w * per, b = a + per < m ? a + per : m;
    if (a >= b) break;
    pool.emplace_back(gather_range, i, j, c, std::cref(pi), k0 + a, k0 + b, dst_i, dst_j, dst_c, a);
  }
  gather_range(i, j, c, pi, k0, k0 + (per < m ? per : m), dst_i, dst_j, dst_c, 0);
  for (auto& th : pool) th.join();
  return ESR_OK;
}
