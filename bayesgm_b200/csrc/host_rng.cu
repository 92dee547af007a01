// Host-side (no device code): NumPy's LEGACY generator restated natively, so that the index and
// prior streams of the training loops (causalbgm/base.py:406-413 `np.random.choice(n, bs,
// replace=False)` + `Gaussian_sampler.get_batch` = `np.random.normal`, prior_samplers.py:46-59;
// :489 the epoch permutation) can be produced bit-exactly off the Python thread: `choice` without
// replacement permutes all n indices per call, which at n = 1e5..1e6 costs more host time than the
// training step it feeds.
//
// Restated from NumPy's published algorithms (numpy/random: _mt19937.c, distributions.c
// `random_interval`, legacy-distributions.c `legacy_gauss`, mtrand.pyx `shuffle` / `permutation` /
// `choice`); the legacy stream is frozen by NumPy's compatibility policy (NEP 19).  The state
// round-trips through `np.random.get_state()` / `set_state()`; tests/test_host_rng.py checks
// bit-equality against NumPy itself.
#include <cmath>
#include <cstring>
#include <string>

#include "common.cuh"

namespace {

inline void mt_refill(bgm_mt19937_state* s) {
  const uint32_t UPPER = 0x80000000u, LOWER = 0x7fffffffu, A = 0x9908b0dfu;
  uint32_t* mt = s->key;
  int kk;
  for (kk = 0; kk < 624 - 397; ++kk) {
    const uint32_t y = (mt[kk] & UPPER) | (mt[kk + 1] & LOWER);
    mt[kk] = mt[kk + 397] ^ (y >> 1) ^ ((y & 1u) ? A : 0u);
  }
  for (; kk < 623; ++kk) {
    const uint32_t y = (mt[kk] & UPPER) | (mt[kk + 1] & LOWER);
    mt[kk] = mt[kk + (397 - 624)] ^ (y >> 1) ^ ((y & 1u) ? A : 0u);
  }
  const uint32_t y = (mt[623] & UPPER) | (mt[0] & LOWER);
  mt[623] = mt[396] ^ (y >> 1) ^ ((y & 1u) ? A : 0u);
  s->pos = 0;
}
inline uint32_t mt_next32(bgm_mt19937_state* s) {
  if (s->pos >= 624) mt_refill(s);
  uint32_t y = s->key[s->pos++];
  y ^= (y >> 11);
  y ^= (y << 7) & 0x9d2c5680u;
  y ^= (y << 15) & 0xefc60000u;
  y ^= (y >> 18);
  return y;
}
inline double mt_next_double(bgm_mt19937_state* s) {
  const int32_t a = (int32_t)(mt_next32(s) >> 5), b = (int32_t)(mt_next32(s) >> 6);
  return (a * 67108864.0 + b) / 9007199254740992.0;
}
// distributions.c random_interval: rejection sampling under the smallest all-ones mask >= max
inline uint32_t interval32(bgm_mt19937_state* s, uint32_t max) {
  if (max == 0) return 0;
  uint32_t mask = max;
  mask |= mask >> 1; mask |= mask >> 2; mask |= mask >> 4; mask |= mask >> 8; mask |= mask >> 16;
  uint32_t v;
  while ((v = (mt_next32(s) & mask)) > max) {}
  return v;
}
inline double legacy_gauss(bgm_mt19937_state* s) {
  if (s->has_gauss) {
    const double t = s->gauss;
    s->has_gauss = 0;
    s->gauss = 0.0;
    return t;
  }
  double f, x1, x2, r2;
  do {
    x1 = 2.0 * mt_next_double(s) - 1.0;
    x2 = 2.0 * mt_next_double(s) - 1.0;
    r2 = x1 * x1 + x2 * x2;
  } while (r2 >= 1.0 || r2 == 0.0);
  f = sqrt(-2.0 * log(r2) / r2);
  s->gauss = f * x1;
  s->has_gauss = 1;
  return f * x2;
}
// permutation(n)[:bs]: Fisher-Yates from the top over arange(n) (mtrand.pyx shuffle -> _shuffle_raw)
inline void choice_no_replace(bgm_mt19937_state* s, int n, int bs, int32_t* out, int32_t* work) {
  for (int i = 0; i < n; ++i) work[i] = i;
  for (int i = n - 1; i >= 1; --i) {
    const uint32_t j = interval32(s, (uint32_t)i);
    const int32_t t = work[i];
    work[i] = work[j];
    work[j] = t;
  }
  memcpy(out, work, sizeof(int32_t) * (size_t)bs);
}

}  // namespace

extern "C" {

int bgm_host_choice(bgm_mt19937_state* st, int n, int size, int32_t* out, int32_t* work) {
  if (!st || !out || !work || n < 1 || size < 0 || size > n) return bgm::fail(BGM_ERR_ARG, "bgm_host_choice: bad argument");
  choice_no_replace(st, n, size, out, work);
  return 0;
}

int bgm_host_normal(bgm_mt19937_state* st, double loc, double scale, long long count, float* out) {
  if (!st || !out || count < 0) return bgm::fail(BGM_ERR_ARG, "bgm_host_normal: bad argument");
  for (long long i = 0; i < count; ++i) out[i] = (float)(loc + scale * legacy_gauss(st));
  return 0;
}

int bgm_host_rand(bgm_mt19937_state* st, long long count, double* out) {
  if (!st || !out || count < 0) return bgm::fail(BGM_ERR_ARG, "bgm_host_rand: bad argument");
  for (long long i = 0; i < count; ++i) out[i] = mt_next_double(st);
  return 0;
}

int bgm_host_egm_stream(bgm_mt19937_state* st, int n, int bs, int zd, int g_d_freq, int iters, int32_t* idx_out,
                        float* z_out, int32_t* work) {
  if (!st || !idx_out || !z_out || !work || n < 1 || bs < 1 || bs > n || zd < 1 || g_d_freq < 0 || iters < 0)
    return bgm::fail(BGM_ERR_ARG, "bgm_host_egm_stream: bad argument");
  const size_t zrow = (size_t)bs * zd;
  for (int c = 0; c < iters; ++c) {
    int32_t* idx = idx_out + (size_t)c * (g_d_freq + 1) * bs;
    float* z = z_out + (size_t)c * (g_d_freq + 1) * zrow;
    for (int k = 0; k < g_d_freq; ++k) {                          // causalbgm/base.py:405-409
      choice_no_replace(st, n, bs, idx + (size_t)k * bs, work);   // :406
      for (size_t i = 0; i < zrow; ++i) z[(size_t)k * zrow + i] = (float)(0.0 + 1.0 * legacy_gauss(st));   // :407
    }
    for (size_t i = 0; i < zrow; ++i) z[(size_t)g_d_freq * zrow + i] = (float)(0.0 + 1.0 * legacy_gauss(st));  // :412
    choice_no_replace(st, n, bs, idx + (size_t)g_d_freq * bs, work);                                          // :413
  }
  return 0;
}

}  // extern "C"
