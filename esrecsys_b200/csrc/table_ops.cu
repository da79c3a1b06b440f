// Embedding-table row operations and the optimizer rules (HBM-bound, 128-bit accesses).
//
//  esr_table_gather_f32   <- jnp.take behind nn.Embed.__call__ (wikipedia/models.py:31-34,
//                            spotify/models.py:43-44)
//  esr_table_export_f32   <- reading state.params[...]['embedding'] (wikipedia/train_cooccurence.py:133)
//  esr_sparse_adagrad_f32 <- optax.adagrad applied to the touched rows only (north-star rule)
//  esr_scatter_rows_f32   <- the dense gradient pytree of jax.value_and_grad
//                            (wikipedia/train_cooccurence.py:86-87): zero except the touched rows
//  esr_dense_adam_f32     <- optax.adam via TrainState.apply_gradients (wikipedia/train_cooccurence.py:101,171)
//  esr_dense_sgdm_f32     <- optax.sgd(momentum) (spotify/train_spotify.py:238-241)
//
// Row kernels: a group of TPR lanes (TPR = min(32, pow2ceil(D/4))) owns a row and every group
// keeps kRowsPerGroup independent rows in flight, so each lane has several 16-byte loads
// outstanding before the first use (the gather is latency-bound otherwise).
#include "esr_common.cuh"

namespace esr {
namespace {

constexpr int kThreads = 256;
constexpr int kRowsPerGroup = 4;

__device__ __forceinline__ const float4* cur_row(const float* r0, const float* r1, const uint8_t* ver, int64_t row,
                                                 int D4) {
  const float* base = (ver != nullptr && ver[row]) ? r1 : r0;
  return reinterpret_cast<const float4*>(base) + row * D4;
}

// out[k,:] = table[ids[k],:]   (ids == nullptr: identity, i.e. export)
template <int TPR>
__global__ void __launch_bounds__(kThreads) k_gather(const float* __restrict__ r0, const float* __restrict__ r1,
                                                     const uint8_t* __restrict__ ver, const int32_t* __restrict__ ids,
                                                     int64_t n, int D4, float4* __restrict__ out) {
  const int lane = threadIdx.x % TPR;
  const int64_t group = (blockIdx.x * (int64_t)kThreads + threadIdx.x) / TPR;
  const int64_t first = group * kRowsPerGroup;
  const float4* src[kRowsPerGroup];
#pragma unroll
  for (int r = 0; r < kRowsPerGroup; ++r) {
    const int64_t k = first + r;
    src[r] = nullptr;
    if (k < n) {
      const int64_t row = ids ? (int64_t)ids[k] : k;
      src[r] = cur_row(r0, r1, ver, row, D4);
    }
  }
  for (int c = lane; c < D4; c += TPR) {
    float4 v[kRowsPerGroup];
#pragma unroll
    for (int r = 0; r < kRowsPerGroup; ++r)
      if (src[r]) v[r] = __ldg(src[r] + c);
#pragma unroll
    for (int r = 0; r < kRowsPerGroup; ++r)
      if (src[r]) st_stream(out + (first + r) * D4 + c, v[r]);
  }
}

// In-place Adagrad on the CURRENT buffer of each listed row (no concurrent readers).
template <int TPR>
__global__ void __launch_bounds__(kThreads) k_sparse_adagrad(float* __restrict__ r0, float* __restrict__ r1,
                                                             const uint8_t* __restrict__ ver, float* __restrict__ acc,
                                                             const int32_t* __restrict__ uniq,
                                                             const int32_t* __restrict__ n_uniq, int D4,
                                                             const float4* __restrict__ g, float lr, float eps) {
  const int lane = threadIdx.x % TPR;
  const int64_t group = (blockIdx.x * (int64_t)kThreads + threadIdx.x) / TPR;
  const int64_t first = group * kRowsPerGroup;
  const int64_t n = *n_uniq;
#pragma unroll
  for (int r = 0; r < kRowsPerGroup; ++r) {
    const int64_t u = first + r;
    if (u >= n) continue;
    const int64_t row = uniq[u];
    float4* p = const_cast<float4*>(cur_row(r0, r1, ver, row, D4));
    float4* a = reinterpret_cast<float4*>(acc) + row * D4;
    for (int c = lane; c < D4; c += TPR) {
      float4 pv = p[c], av = ld_stream(a + c);
      const float4 gv = ld_stream(g + u * D4 + c);
      adagrad4(pv, av, gv, lr, eps);
      p[c] = pv;
      st_stream(a + c, av);
    }
  }
}

__global__ void __launch_bounds__(kThreads) k_sparse_adagrad_bias(float* __restrict__ bias, float* __restrict__ bacc,
                                                                  const int32_t* __restrict__ uniq,
                                                                  const int32_t* __restrict__ n_uniq,
                                                                  const float* __restrict__ gb, float lr, float eps) {
  const int64_t u = blockIdx.x * (int64_t)kThreads + threadIdx.x;
  if (u >= *n_uniq) return;
  const int64_t row = uniq[u];
  float p = bias[row], a = bacc[row];
  adagrad1(p, a, gb[u], lr, eps);
  bias[row] = p;
  bacc[row] = a;
}

template <int TPR>
__global__ void __launch_bounds__(kThreads) k_scatter_rows(float4* __restrict__ dst, int D4,
                                                           const int32_t* __restrict__ uniq,
                                                           const int32_t* __restrict__ n_uniq,
                                                           const float4* __restrict__ g, int accumulate) {
  const int lane = threadIdx.x % TPR;
  const int64_t group = (blockIdx.x * (int64_t)kThreads + threadIdx.x) / TPR;
  const int64_t first = group * kRowsPerGroup;
  const int64_t n = *n_uniq;
#pragma unroll
  for (int r = 0; r < kRowsPerGroup; ++r) {
    const int64_t u = first + r;
    if (u >= n) continue;
    float4* d = dst + (int64_t)uniq[u] * D4;
    for (int c = lane; c < D4; c += TPR) {
      float4 v = ld_stream(g + u * D4 + c);
      if (accumulate) f4_add(v, d[c]);
      d[c] = v;
    }
  }
}

// optax.adam: mu = b1 mu + (1-b1) g ; nu = b2 nu + (1-b2) g^2 ; p -= lr (mu/c1) / (sqrt(nu/c2) + eps)
struct AdamK {
  float lr, b1, b2, omb1, omb2, eps, c1, c2;  // omb = 1 - b, c = 1 - b^count: evaluated in double on the host
};

__device__ __forceinline__ void adam1(float& p, float g, float& mu, float& nu, const AdamK& k) {
  const float lr = k.lr, eps = k.eps, c1 = k.c1, c2 = k.c2;
  mu = k.b1 * mu + k.omb1 * g;
  nu = k.b2 * nu + k.omb2 * (g * g);
  const float mhat = mu / c1;
  const float vhat = nu / c2;
  p = p - lr * mhat / (sqrtf(vhat) + eps);
}

__global__ void __launch_bounds__(kThreads) k_dense_adam(float* __restrict__ p, const float* __restrict__ g,
                                                         float* __restrict__ mu, float* __restrict__ nu, int64_t n,
                                                         const AdamK k) {
  const int64_t n4 = n >> 2;
  const int64_t stride = (int64_t)gridDim.x * kThreads;
  for (int64_t i = blockIdx.x * (int64_t)kThreads + threadIdx.x; i < n4; i += stride) {
    float4 pv = reinterpret_cast<float4*>(p)[i];
    const float4 gv = ld_stream(reinterpret_cast<const float4*>(g) + i);
    float4 m = reinterpret_cast<float4*>(mu)[i], v = reinterpret_cast<float4*>(nu)[i];
    adam1(pv.x, gv.x, m.x, v.x, k);
    adam1(pv.y, gv.y, m.y, v.y, k);
    adam1(pv.z, gv.z, m.z, v.z, k);
    adam1(pv.w, gv.w, m.w, v.w, k);
    reinterpret_cast<float4*>(p)[i] = pv;
    reinterpret_cast<float4*>(mu)[i] = m;
    reinterpret_cast<float4*>(nu)[i] = v;
  }
  // tail (n % 4) by the first threads of block 0
  if (blockIdx.x == 0 && threadIdx.x < (n & 3)) {
    const int64_t i = (n4 << 2) + threadIdx.x;
    adam1(p[i], g[i], mu[i], nu[i], k);
  }
}

__global__ void __launch_bounds__(kThreads) k_dense_sgdm(float* __restrict__ p, const float* __restrict__ g,
                                                         float* __restrict__ tr, int64_t n, float lr, float mom) {
  const int64_t n4 = n >> 2;
  const int64_t stride = (int64_t)gridDim.x * kThreads;
  for (int64_t i = blockIdx.x * (int64_t)kThreads + threadIdx.x; i < n4; i += stride) {
    float4 pv = reinterpret_cast<float4*>(p)[i];
    const float4 gv = ld_stream(reinterpret_cast<const float4*>(g) + i);
    float4 t = reinterpret_cast<float4*>(tr)[i];
    t.x = gv.x + mom * t.x; t.y = gv.y + mom * t.y; t.z = gv.z + mom * t.z; t.w = gv.w + mom * t.w;
    pv.x -= lr * t.x; pv.y -= lr * t.y; pv.z -= lr * t.z; pv.w -= lr * t.w;
    reinterpret_cast<float4*>(p)[i] = pv;
    reinterpret_cast<float4*>(tr)[i] = t;
  }
  if (blockIdx.x == 0 && threadIdx.x < (n & 3)) {
    const int64_t i = (n4 << 2) + threadIdx.x;
    const float t = g[i] + mom * tr[i];
    tr[i] = t;
    p[i] -= lr * t;
  }
}

// out[b] = sum_d x[b,d] * y[b,d]   (jax.vmap(jnp.dot) wikipedia/models.py:35-36; sum(a*b,-1) pinterest/models.py:67-72)
__global__ void __launch_bounds__(kThreads) k_rowwise_dot(const float* __restrict__ x, const float* __restrict__ y,
                                                          int64_t B, int D, float* __restrict__ out) {
  const int lane = threadIdx.x & 31;
  const int64_t b = blockIdx.x * (int64_t)(kThreads / 32) + (threadIdx.x >> 5);
  if (b >= B) return;
  float acc = 0.f;
  for (int c = lane; c < D; c += 32) acc = fmaf(x[b * D + c], y[b * D + c], acc);
  acc = warp_sum(acc);
  if (lane == 0) out[b] = acc;
}

// scores[v, t] = E[v] . Q[t]   (Glove.score_all wikipedia/models.py:40-55; find_top_k
// pinterest/make_recommendations.py:57).  One warp per table row, queries staged in shared memory;
// the table is streamed once (HBM-bound: V*R bytes).
constexpr int kMaxQueries = 64;
__global__ void __launch_bounds__(kThreads) k_score_all(const float* __restrict__ r0, const float* __restrict__ r1,
                                                        const uint8_t* __restrict__ ver, int64_t V, int D,
                                                        const float* __restrict__ Q, int T, float* __restrict__ out) {
  extern __shared__ float qs[];  // [T][D]
  for (int i = threadIdx.x; i < T * D; i += kThreads) qs[i] = Q[i];
  __syncthreads();
  const int lane = threadIdx.x & 31;
  const int D4 = D / 4;
  for (int64_t v = blockIdx.x * (int64_t)(kThreads / 32) + (threadIdx.x >> 5); v < V;
       v += (int64_t)gridDim.x * (kThreads / 32)) {
    const float4* row = cur_row(r0, r1, ver, v, D4);
    for (int t0 = 0; t0 < T; t0 += 8) {
      float acc[8];
#pragma unroll
      for (int k = 0; k < 8; ++k) acc[k] = 0.f;
      for (int c = lane; c < D4; c += 32) {
        const float4 x = ld_stream(row + c);
#pragma unroll
        for (int k = 0; k < 8; ++k)
          if (t0 + k < T) acc[k] += f4_dot(x, reinterpret_cast<const float4*>(qs + (t0 + k) * D)[c]);
      }
#pragma unroll
      for (int k = 0; k < 8; ++k) {
        const float s = warp_sum(acc[k]);
        if (lane == 0 && t0 + k < T) out[v * T + t0 + k] = s;
      }
    }
  }
}

int tpr_for(int D4) {
  int t = 1;
  while (t < D4 && t < 32) t <<= 1;
  return t;
}

bool table_ok(const EsrTable* t) {
  return t != nullptr && t->struct_size >= sizeof(EsrTable) && t->D > 0 && (t->D % 4) == 0 && t->V >= 0 &&
         t->rows[0] != nullptr && (reinterpret_cast<uintptr_t>(t->rows[0]) % 16) == 0 &&
         (t->rows[1] == nullptr || (reinterpret_cast<uintptr_t>(t->rows[1]) % 16) == 0) &&
         ((t->ver == nullptr) || (t->rows[1] != nullptr));
}

#define ESR_DISPATCH_TPR(tpr, CALL) \
  switch (tpr) {                    \
    case 1: { constexpr int TPR = 1; CALL; } break;   \
    case 2: { constexpr int TPR = 2; CALL; } break;   \
    case 4: { constexpr int TPR = 4; CALL; } break;   \
    case 8: { constexpr int TPR = 8; CALL; } break;   \
    case 16: { constexpr int TPR = 16; CALL; } break; \
    default: { constexpr int TPR = 32; CALL; } break; \
  }

unsigned row_grid(int64_t rows, int tpr) {
  const int64_t groups = ceil_div(rows, kRowsPerGroup);
  return (unsigned)ceil_div(groups * tpr, kThreads);
}

int gather_impl(const EsrTable* t, const int32_t* ids, int64_t n, float* out, cudaStream_t stream) {
  if (n == 0) return ESR_OK;
  const int D4 = t->D / 4;
  const int tpr = tpr_for(D4);
  ESR_DISPATCH_TPR(tpr, (k_gather<TPR><<<row_grid(n, tpr), kThreads, 0, stream>>>(
                            t->rows[0], t->rows[1], t->ver, ids, n, D4, reinterpret_cast<float4*>(out))));
  ESR_LAUNCH_CHECK();
  return ESR_OK;
}

}  // namespace
}  // namespace esr

using namespace esr;

extern "C" int esr_table_gather_f32(const EsrTable* t, const int32_t* ids, int64_t n, float* out, esr_stream_t stream) {
  ESR_RANGE("esr_table_gather_f32");
  ESR_REQUIRE(table_ok(t) && n >= 0 && (n == 0 || (ids != nullptr && out != nullptr)));
  ESR_REQUIRE((reinterpret_cast<uintptr_t>(out) % 16) == 0);
  return gather_impl(t, ids, n, out, static_cast<cudaStream_t>(stream));
}

extern "C" int esr_table_export_f32(const EsrTable* t, float* out, esr_stream_t stream) {
  ESR_REQUIRE(table_ok(t) && (t->V == 0 || out != nullptr) && (reinterpret_cast<uintptr_t>(out) % 16) == 0);
  return gather_impl(t, nullptr, t->V, out, static_cast<cudaStream_t>(stream));
}

extern "C" int esr_sparse_adagrad_f32(EsrTable* t, const int32_t* uniq, const int32_t* n_uniq, int64_t cap,
                                      const float* g, const float* gb, float lr, float eps, esr_stream_t stream_) {
  ESR_RANGE("esr_sparse_adagrad_f32");
  ESR_REQUIRE(table_ok(t) && cap >= 0 && uniq != nullptr && n_uniq != nullptr);
  ESR_REQUIRE(g == nullptr || (t->acc != nullptr && (reinterpret_cast<uintptr_t>(g) % 16) == 0));
  ESR_REQUIRE(gb == nullptr || (t->bias != nullptr && t->bias_acc != nullptr));
  if (cap == 0) return ESR_OK;
  cudaStream_t stream = static_cast<cudaStream_t>(stream_);
  const int D4 = t->D / 4;
  const int tpr = tpr_for(D4);
  if (g != nullptr) {
    ESR_DISPATCH_TPR(tpr, (k_sparse_adagrad<TPR><<<row_grid(cap, tpr), kThreads, 0, stream>>>(
                              t->rows[0], t->rows[1], t->ver, t->acc, uniq, n_uniq, D4,
                              reinterpret_cast<const float4*>(g), lr, eps)));
    ESR_LAUNCH_CHECK();
  }
  if (gb != nullptr) {
    k_sparse_adagrad_bias<<<(unsigned)ceil_div(cap, kThreads), kThreads, 0, stream>>>(t->bias, t->bias_acc, uniq, n_uniq,
                                                                                      gb, lr, eps);
    ESR_LAUNCH_CHECK();
  }
  return ESR_OK;
}

extern "C" int esr_scatter_rows_f32(float* dst, int32_t D, const int32_t* uniq, const int32_t* n_uniq, int64_t cap,
                                    const float* g, int32_t accumulate, esr_stream_t stream_) {
  ESR_REQUIRE(dst != nullptr && D > 0 && (D % 4) == 0 && uniq != nullptr && n_uniq != nullptr && g != nullptr && cap >= 0);
  ESR_REQUIRE((reinterpret_cast<uintptr_t>(dst) % 16) == 0 && (reinterpret_cast<uintptr_t>(g) % 16) == 0);
  if (cap == 0) return ESR_OK;
  const int D4 = D / 4;
  const int tpr = tpr_for(D4);
  ESR_DISPATCH_TPR(tpr, (k_scatter_rows<TPR><<<row_grid(cap, tpr), kThreads, 0, static_cast<cudaStream_t>(stream_)>>>(
                            reinterpret_cast<float4*>(dst), D4, uniq, n_uniq, reinterpret_cast<const float4*>(g),
                            accumulate)));
  ESR_LAUNCH_CHECK();
  return ESR_OK;
}

static unsigned dense_grid(int64_t n) {
  const int64_t want = ceil_div(ceil_div(n, 4), kThreads);
  const int64_t cap = (int64_t)sm_count() * 16;
  return (unsigned)(want < 1 ? 1 : (want > cap ? cap : want));
}

extern "C" int esr_dense_adam_f32(float* p, const float* g, float* mu, float* nu, int64_t n, double lr, double b1,
                                  double b2, double eps, int64_t count, esr_stream_t stream_) {
  ESR_RANGE("esr_dense_adam_f32");
  ESR_REQUIRE(n >= 0 && count >= 1);
  if (n == 0) return ESR_OK;
  ESR_REQUIRE(p && g && mu && nu);
  ESR_REQUIRE(((reinterpret_cast<uintptr_t>(p) | reinterpret_cast<uintptr_t>(g) | reinterpret_cast<uintptr_t>(mu) |
                reinterpret_cast<uintptr_t>(nu)) % 16) == 0);
  // bias corrections in double on the host, like optax evaluates 1 - b^count in python floats / f32
  AdamK k;
  k.lr = (float)lr; k.b1 = (float)b1; k.b2 = (float)b2; k.omb1 = (float)(1.0 - b1); k.omb2 = (float)(1.0 - b2);
  k.eps = (float)eps;
  k.c1 = (float)(1.0 - pow(b1, (double)count));
  k.c2 = (float)(1.0 - pow(b2, (double)count));
  k_dense_adam<<<dense_grid(n), kThreads, 0, static_cast<cudaStream_t>(stream_)>>>(p, g, mu, nu, n, k);
  ESR_LAUNCH_CHECK();
  return ESR_OK;
}

extern "C" int esr_dense_sgdm_f32(float* p, const float* g, float* trace, int64_t n, float lr, float momentum,
                                  esr_stream_t stream_) {
  ESR_REQUIRE(n >= 0);
  if (n == 0) return ESR_OK;
  ESR_REQUIRE(p && g && trace);
  ESR_REQUIRE(((reinterpret_cast<uintptr_t>(p) | reinterpret_cast<uintptr_t>(g) | reinterpret_cast<uintptr_t>(trace)) % 16) == 0);
  k_dense_sgdm<<<dense_grid(n), kThreads, 0, static_cast<cudaStream_t>(stream_)>>>(p, g, trace, n, lr, momentum);
  ESR_LAUNCH_CHECK();
  return ESR_OK;
}

extern "C" int esr_rowwise_dot_f32(const float* x, const float* y, int64_t B, int32_t D, float* out, esr_stream_t stream_) {
  ESR_REQUIRE(B >= 0 && D > 0 && (B == 0 || (x && y && out)));
  if (B == 0) return ESR_OK;
  k_rowwise_dot<<<(unsigned)ceil_div(B, kThreads / 32), kThreads, 0, static_cast<cudaStream_t>(stream_)>>>(x, y, B, D, out);
  ESR_LAUNCH_CHECK();
  return ESR_OK;
}

extern "C" int esr_score_all_f32(const EsrTable* t, const float* queries, int32_t T, float* scores, esr_stream_t stream_) {
  ESR_RANGE("esr_score_all_f32");
  ESR_REQUIRE(table_ok(t) && T >= 1 && T <= kMaxQueries && queries && scores);
  ESR_REQUIRE((reinterpret_cast<uintptr_t>(queries) % 16) == 0);
  if (t->V == 0) return ESR_OK;
  const size_t smem = (size_t)T * t->D * sizeof(float);
  if (smem > 200 * 1024) return ESR_ENOTSUP;
  static SmemOptIn configured;
  if (smem > 48 * 1024 && configured.raise(smem))
    ESR_CUDA(cudaFuncSetAttribute(k_score_all, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
  const int64_t want = ceil_div(t->V, kThreads / 32);
  const int64_t cap = (int64_t)sm_count() * 8;
  k_score_all<<<(unsigned)(want < cap ? want : cap), kThreads, smem, static_cast<cudaStream_t>(stream_)>>>(
      t->rows[0], t->rows[1], t->ver, t->V, t->D, queries, T, scores);
  ESR_LAUNCH_CHECK();
  return ESR_OK;
}

// ---------------------------------------------------------------------------------------------
// On-device negative sampling.  Replaces sample_negative's
// jax.random.randint(subkey, [num_negatives], 0, N - 1) (spotify/train_spotify.py:139-150; the upper bound is
// EXCLUSIVE, so index N-1 is never drawn -- kept) and the pre-sampled negative of pinterest/train_shop_the_look.py:72-91.
// The stream is our own counter-based generator (splitmix64 of (seed, step, k)), not threefry: parity harnesses
// treat negatives as inputs; oracle.index.sample_uniform restates it bit for bit.
// ---------------------------------------------------------------------------------------------
namespace esr {
namespace {
__host__ __device__ inline uint64_t splitmix64(uint64_t x) {
  x += 0x9E3779B97F4A7C15ull;
  x = (x ^ (x >> 30)) * 0xBF58476D1CE4E5B9ull;
  x = (x ^ (x >> 27)) * 0x94D049BB133111EBull;
  return x ^ (x >> 31);
}
__global__ void __launch_bounds__(kThreads) k_sample_uniform(uint64_t seed, uint64_t step, int64_t n, uint32_t hi,
                                                             int32_t* __restrict__ out) {
  const int64_t k = blockIdx.x * (int64_t)kThreads + threadIdx.x;
  if (k >= n) return;
  const uint64_t r = splitmix64(splitmix64(seed ^ (step * 0xD1342543DE82EF95ull)) + (uint64_t)k);
  out[k] = (int32_t)(((r >> 32) * (uint64_t)hi) >> 32);  // multiply-shift: unbiased to 2^-32
}
}  // namespace
}  // namespace esr

extern "C" int esr_sample_uniform_i32(uint64_t seed, uint64_t step, int64_t n, int64_t hi, int32_t* out, esr_stream_t stream_) {
  ESR_REQUIRE(n >= 0 && hi > 0 && hi <= 0x7fffffffll && (out != nullptr || n == 0));
  if (n == 0) return ESR_OK;
  esr::k_sample_uniform<<<(unsigned)esr::ceil_div(n, esr::kThreads), esr::kThreads, 0, static_cast<cudaStream_t>(stream_)>>>(
      seed, step, n, (uint32_t)hi, out);
  ESR_LAUNCH_CHECK();
  return ESR_OK;
}
