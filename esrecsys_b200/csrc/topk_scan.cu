// Fused retrieval: ONE streaming pass over the candidate rows computes their scores against the queries and keeps the
// running top-k per query in shared memory -- no (V, T) score matrix in HBM, no full sort.  The retrieval step that
// follows training in every reference script:
//   dump_knn     wikipedia/train_cooccurence.py:114-126  top 10 of  E[v] . E[tok_t]   (read from the tail of jnp.argsort:
//                                                        ties come out HIGHER index first)
//   eval_step    spotify/train_spotify.py:113-131        jax.lax.top_k(neg_affinity, 500) over all 2.26 M tracks, where
//                spotify/models.py:78-80                 neg_affinity = max_k(track . ctx_k) + 0.1 isin(album, album_ctx)
//                                                        + 0.1 isin(artist, artist_ctx), track = concat(album_embed[album
//                                                        % 100000], artist_embed[artist])      (ties: LOWER index first)
//   find_top_k   pinterest/make_recommendations.py:49-65 jax.lax.top_k(scene . product_v, k)
// HBM-bound: N * D * 4 bytes of rows are read once (plus 4-8 bytes of ids per row for the gathered form).
//
// k_topk_scan   persistent CTAs over contiguous slabs of rows; a row is owned by G lanes (4 float4 each), the queries sit
//               in shared memory.  A (score, row) pair is ONE 64-bit key: order-preserving image of the float in the high
//               word, the row (complemented when ties must come out lower-index-first) in the low word, so "largest key
//               first" is exactly the reference's order.  Keys that beat the list's current threshold are appended to a
//               shared-memory buffer; when a buffer could overflow, the block bitonic-sorts it, keeps the best k and
//               raises the threshold to the k-th key (after the first few hundred rows almost nothing passes).
//               Every CTA leaves its best k keys per list in the workspace.
// k_topk_merge  one CTA per list runs the same select over the P x k surviving keys and writes (index, score) best first.
// The result does not depend on the grid: the k largest keys of a set are unique.
#include "esr_common.cuh"

namespace esr {
namespace {

constexpr int kThreads = 256;
constexpr int kMaxT = 64;
constexpr int kMaxK = 1024;
constexpr int kMaxCtx = 32;

struct TopkArgs {
  const float* A0;
  const float* A1;
  const uint8_t* ver;
  const float* Bt;
  const int32_t* idxA;
  const int32_t* idxB;
  int32_t modA;
  int32_t DA4, DB4;
  int64_t N;
  const float* Q;
  int32_t T;
  int32_t maxq;
  const int32_t* ctxA;
  const int32_t* ctxB;
  int32_t nA, nB;
  float boost;
  int32_t k, order, cap, tile_rows;
  int64_t slab;
  uint64_t* cand;
};

__device__ __forceinline__ uint64_t make_key(float s, uint32_t row, int order) {
  s += 0.f;  // -0.0 -> +0.0: the reference compares values, so the two zeros tie
  uint32_t u = __float_as_uint(s);
  u ^= (u >> 31) ? 0xffffffffu : 0x80000000u;
  return ((uint64_t)u << 32) | (order ? row : ~row);
}
__device__ __forceinline__ float key_score(uint64_t key) {
  uint32_t u = (uint32_t)(key >> 32);
  u ^= (u >> 31) ? 0x80000000u : 0xffffffffu;
  return __uint_as_float(u);
}
__device__ __forceinline__ uint32_t key_row(uint64_t key, int order) {
  const uint32_t lo = (uint32_t)key;
  return order ? lo : ~lo;
}

// descending bitonic sort of cap (power of two) keys in shared memory by the whole block
__device__ void block_sort_desc(uint64_t* keys, int cap) {
  for (int size = 2; size <= cap; size <<= 1) {
    for (int stride = size >> 1; stride > 0; stride >>= 1) {
      for (int i = threadIdx.x; i < (cap >> 1); i += kThreads) {
        const int lo = 2 * i - (i & (stride - 1));
        const int hi = lo + stride;
        const bool desc = (lo & size) == 0;
        const uint64_t a = keys[lo], b = keys[hi];
        if ((a < b) == desc) {
          keys[lo] = b;
          keys[hi] = a;
        }
      }
      __syncthreads();
    }
  }
}

// keep the best k keys of one list; threshold = the k-th key (0 while the list holds fewer than k)
__device__ void compact_list(uint64_t* keys, int* cnt, uint64_t* thr, int cap, int k) {
  const int c = *cnt;
  __syncthreads();
  for (int i = c + threadIdx.x; i < cap; i += kThreads) keys[i] = 0ull;  // 0 is below every real key
  __syncthreads();
  block_sort_desc(keys, cap);
  if (threadIdx.x == 0) {
    *cnt = c < k ? c : k;
    *thr = c >= k ? keys[k - 1] : 0ull;
  }
  __syncthreads();
}

// Sum G per-lane partial values of G different queries over the G lanes of a group so that lane gl ends up with the total of
// query gl: log2(G) exchange steps that halve the number of live values (G - 1 shuffles instead of G * log2(G)).
template <int G>
__device__ __forceinline__ float transpose_sum(float (&v)[G], int gl) {
#pragma unroll
  for (int o = G / 2, n = G; o > 0; o >>= 1, n >>= 1) {
    const bool up = (gl & o) != 0;
#pragma unroll
    for (int j = 0; j < n / 2; ++j) {
      const float send = up ? v[j] : v[j + n / 2];
      const float keep = up ? v[j + n / 2] : v[j];
      v[j] = keep + __shfl_xor_sync(FULL, send, o);
    }
  }
  return v[0];
}

// R rows per group and iteration (2 * NV independent 16-byte loads in flight per lane), G queries per pass: the query
// vectors are read from shared memory once per pass for all R rows.
// FULL: D4 == G * NV (every lane column exists: no per-column bounds checks -- they were half of the instructions).
template <int G, int NV, bool FULL>
__global__ void __launch_bounds__(kThreads) k_topk_scan(const TopkArgs a) {
  extern __shared__ __align__(16) unsigned char smem_raw[];
  constexpr int R = 2;
  const int D4 = FULL ? G * NV : a.DA4 + a.DB4;
  const int L = a.maxq ? 1 : a.T;
  const int Tp = (a.T + G - 1) / G * G;                                               // queries padded to whole passes
  float4* qs = reinterpret_cast<float4*>(smem_raw);                                   // [Tp][D4]
  uint64_t* keys = reinterpret_cast<uint64_t*>(smem_raw + (size_t)Tp * D4 * 16);      // [L][cap]
  __shared__ int cnt[kMaxT];
  __shared__ uint64_t thr[kMaxT];
  __shared__ int32_t ctx[2 * kMaxCtx];
  for (int i = threadIdx.x; i < Tp * D4; i += kThreads)
    qs[i] = i < a.T * D4 ? reinterpret_cast<const float4*>(a.Q)[i] : f4_zero();
  if (threadIdx.x < L) {
    cnt[threadIdx.x] = 0;
    thr[threadIdx.x] = 0ull;
  }
  if (threadIdx.x < a.nA) ctx[threadIdx.x] = a.ctxA[threadIdx.x];
  if (threadIdx.x < a.nB) ctx[kMaxCtx + threadIdx.x] = a.ctxB[threadIdx.x];
  __syncthreads();
  constexpr int GP = kThreads / G;  // groups per block; a block iteration covers GP * R rows
  const int gl = threadIdx.x % G, grp = threadIdx.x / G;
  const int64_t r0 = blockIdx.x * a.slab, r1 = min(a.N, r0 + a.slab);
  const float4* const A0 = reinterpret_cast<const float4*>(a.A0);
  const float4* const A1 = reinterpret_cast<const float4*>(a.A1);
  const float4* const Bt = reinterpret_cast<const float4*>(a.Bt);

  for (int64_t t0 = r0; t0 < r1; t0 += a.tile_rows) {
    const int64_t t1 = min(r1, t0 + a.tile_rows);
    for (int64_t base = t0; base < t1; base += GP * R) {
      float4 x[R][NV];
      int32_t ia[R], ib[R];
      bool on[R];
      int64_t n[R];
#pragma unroll
      for (int r = 0; r < R; ++r) {
        n[r] = base + (int64_t)r * GP + grp;
        on[r] = n[r] < t1;
        ia[r] = ib[r] = 0;
        if (on[r]) {
          ia[r] = a.idxA ? a.idxA[n[r]] : (int32_t)n[r];
          if (Bt) ib[r] = a.idxB[n[r]];
        }
      }
#pragma unroll
      for (int r = 0; r < R; ++r) {
        if (on[r]) {
          const int64_t ra = a.modA > 0 ? ia[r] % a.modA : ia[r];
          const float4* pa = ((a.ver != nullptr && a.ver[ra]) ? A1 : A0) + ra * a.DA4;
          const float4* pb = Bt ? Bt + (int64_t)ib[r] * a.DB4 : nullptr;
#pragma unroll
          for (int k = 0; k < NV; ++k) {
            const int c = k * G + gl;
            if (FULL && Bt == nullptr) x[r][k] = ld_stream(pa + c);
            else x[r][k] = c < a.DA4 ? ld_stream(pa + c) : ((FULL || c < D4) ? ld_stream(pb + (c - a.DA4)) : f4_zero());
          }
        } else {
#pragma unroll
          for (int k = 0; k < NV; ++k) x[r][k] = f4_zero();
        }
      }
      float best[R];
#pragma unroll
      for (int r = 0; r < R; ++r) best[r] = -INFINITY;
      for (int q0 = 0; q0 < Tp; q0 += G) {
        float d[R][G];
#pragma unroll
        for (int j = 0; j < G; ++j) {
          float4 qv[NV];
#pragma unroll
          for (int k = 0; k < NV; ++k) {
            const int c = k * G + gl;
            qv[k] = (FULL || c < D4) ? qs[(q0 + j) * D4 + c] : f4_zero();
          }
#pragma unroll
          for (int r = 0; r < R; ++r) {
            float acc = 0.f;
#pragma unroll
            for (int k = 0; k < NV; ++k) acc += f4_dot(x[r][k], qv[k]);
            d[r][j] = acc;
          }
        }
#pragma unroll
        for (int r = 0; r < R; ++r) {
          const float s = transpose_sum<G>(d[r], gl);   // lane gl: score of query q0 + gl
          const int q = q0 + gl;
          if (a.maxq) {
            if (q < a.T) best[r] = fmaxf(best[r], s);
          } else if (on[r] && q < a.T) {
            const uint64_t key = make_key(s, (uint32_t)n[r], a.order);
            if (key >= thr[q]) keys[(size_t)q * a.cap + atomicAdd(&cnt[q], 1)] = key;
          }
        }
      }
      if (a.maxq) {
#pragma unroll
        for (int r = 0; r < R; ++r) {
          float m = best[r];
#pragma unroll
          for (int o = G / 2; o > 0; o >>= 1) m = fmaxf(m, __shfl_xor_sync(FULL, m, o));
          if (on[r] && gl == 0) {
            // spotify/models.py:78-80: + 0.1 isin(album, album_context) + 0.1 isin(artist, artist_context), raw ids
            bool inA = false, inB = false;
            for (int j = 0; j < a.nA; ++j) inA |= ctx[j] == ia[r];
            for (int j = 0; j < a.nB; ++j) inB |= ctx[kMaxCtx + j] == ib[r];
            float s = m + a.boost * (inA ? 1.f : 0.f);
            s = s + a.boost * (inB ? 1.f : 0.f);
            const uint64_t key = make_key(s, (uint32_t)n[r], a.order);
            if (key >= thr[0]) keys[atomicAdd(&cnt[0], 1)] = key;
          }
        }
      }
    }
    __syncthreads();
    // a list that could overflow during the next tile is reduced to its best k now (block-uniform decision)
    for (int l = 0; l < L; ++l)
      if (cnt[l] + a.tile_rows > a.cap) compact_list(keys + (size_t)l * a.cap, &cnt[l], &thr[l], a.cap, a.k);
  }
  for (int l = 0; l < L; ++l) {
    compact_list(keys + (size_t)l * a.cap, &cnt[l], &thr[l], a.cap, a.k);
    uint64_t* out = a.cand + ((size_t)l * gridDim.x + blockIdx.x) * a.k;
    for (int i = threadIdx.x; i < a.k; i += kThreads) out[i] = i < cnt[l] ? keys[(size_t)l * a.cap + i] : 0ull;
    __syncthreads();
  }
}

__global__ void __launch_bounds__(kThreads) k_topk_merge(const uint64_t* __restrict__ cand, int64_t n_cand, int k, int cap,
                                                         int order, int32_t* __restrict__ out_idx,
                                                         float* __restrict__ out_val) {
  extern __shared__ __align__(16) unsigned char smem_raw[];
  uint64_t* keys = reinterpret_cast<uint64_t*>(smem_raw);
  __shared__ int cnt;
  __shared__ uint64_t thr;
  if (threadIdx.x == 0) {
    cnt = 0;
    thr = 0ull;
  }
  __syncthreads();
  const uint64_t* src = cand + (size_t)blockIdx.x * n_cand;
  const int tile = cap / 2;
  for (int64_t t0 = 0; t0 < n_cand; t0 += tile) {
    const int64_t t1 = min(n_cand, t0 + tile);
    for (int64_t i = t0 + threadIdx.x; i < t1; i += kThreads) {
      const uint64_t key = src[i];
      if (key != 0ull && key >= thr) keys[atomicAdd(&cnt, 1)] = key;
    }
    __syncthreads();
    if (cnt + tile > cap) compact_list(keys, &cnt, &thr, cap, k);
  }
  compact_list(keys, &cnt, &thr, cap, k);
  for (int i = threadIdx.x; i < k; i += kThreads) {
    const uint64_t key = keys[i];
    out_idx[(size_t)blockIdx.x * k + i] = i < cnt ? (int32_t)key_row(key, order) : -1;
    if (out_val) out_val[(size_t)blockIdx.x * k + i] = i < cnt ? key_score(key) : -INFINITY;
  }
}

int pow2_ceil(int v) {
  int p = 1;
  while (p < v) p <<= 1;
  return p;
}

struct TopkGeom {
  int cap, tile_rows, grid, G, lists;
  int64_t slab;
  size_t smem_scan, smem_merge;
};

bool topk_geom(int64_t N, int D4, int T, int maxq, int k, TopkGeom* g) {
  if (N <= 0 || D4 <= 0 || D4 > 128 || T <= 0 || T > kMaxT || k <= 0 || k > kMaxK || k > N) return false;
  g->lists = maxq ? 1 : T;
  int G = 1;
  while (G * 4 < D4 && G < 32) G <<= 1;
  g->G = G;
  const int Tp = (T + G - 1) / G * G;  // queries padded to whole passes of G
  int cap = pow2_ceil(4 * k);
  if (cap < 512) cap = 512;
  if (cap < 4 * (kThreads / G)) cap = 4 * (kThreads / G);  // at least two block iterations per tile
  if (cap > 4096) cap = 4096;
  // shared memory: queries + lists * cap keys; shrink the buffers (never below 2k rounded up) before giving up
  while ((size_t)Tp * D4 * 16 + (size_t)g->lists * cap * 8 > 100 * 1024 && cap / 2 >= pow2_ceil(2 * k) && cap > 128) cap >>= 1;
  if ((size_t)Tp * D4 * 16 + (size_t)g->lists * cap * 8 > 200 * 1024) return false;
  g->cap = cap;
  const int rows_per_iter = 2 * (kThreads / G);  // R = 2 rows per group
  int tile = cap / 2 < 256 ? cap / 2 : 256;
  tile = tile / rows_per_iter * rows_per_iter;
  if (tile < rows_per_iter) tile = rows_per_iter;
  if (k + tile > cap) return false;
  g->tile_rows = tile;
  const int64_t want = ceil_div(N, (int64_t)tile * 4);  // at least 4 tiles per CTA
  const int64_t cap_grid = 3 * (int64_t)sm_count();
  g->grid = (int)(want < 1 ? 1 : (want < cap_grid ? want : cap_grid));
  g->slab = ceil_div(ceil_div(N, (int64_t)g->grid), (int64_t)rows_per_iter) * rows_per_iter;
  g->grid = (int)ceil_div(N, g->slab);
  g->smem_scan = (size_t)Tp * D4 * 16 + (size_t)g->lists * cap * 8;
  g->smem_merge = (size_t)cap * 8;
  return true;
}

template <int G>
int launch_scan(const TopkArgs& a, const TopkGeom& g, cudaStream_t stream) {
  static SmemOptIn configured;
  if (g.smem_scan > 48 * 1024 && configured.raise(g.smem_scan)) {
    ESR_CUDA(cudaFuncSetAttribute(k_topk_scan<G, 4, true>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)g.smem_scan));
    ESR_CUDA(cudaFuncSetAttribute(k_topk_scan<G, 4, false>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)g.smem_scan));
  }
  if (a.DA4 + a.DB4 == G * 4) k_topk_scan<G, 4, true><<<g.grid, kThreads, g.smem_scan, stream>>>(a);
  else k_topk_scan<G, 4, false><<<g.grid, kThreads, g.smem_scan, stream>>>(a);
  ESR_LAUNCH_CHECK();
  return ESR_OK;
}

}  // namespace
}  // namespace esr

using namespace esr;

extern "C" size_t esr_topk_workspace_bytes(int64_t N, int32_t D, int32_t T, int32_t max_over_queries, int32_t k) {
  TopkGeom g;
  if (D <= 0 || (D % 4) != 0 || !topk_geom(N, D / 4, T, max_over_queries, k, &g)) return 0;
  return align_up((size_t)g.lists * g.grid * k * sizeof(uint64_t), 256) + 256;
}

extern "C" int esr_topk_scan_f32(const EsrTopkCfg* cfg, int32_t* out_idx, float* out_val, void* ws, size_t ws_bytes,
                                 esr_stream_t stream_) {
  ESR_RANGE("esr_topk_scan_f32");
  ESR_REQUIRE(cfg != nullptr && cfg->struct_size >= sizeof(EsrTopkCfg) && out_idx != nullptr && ws != nullptr);
  ESR_REQUIRE(cfg->rows_a != nullptr && cfg->queries != nullptr && cfg->Da > 0 && (cfg->Da % 4) == 0 && cfg->Db >= 0 &&
              (cfg->Db % 4) == 0);
  ESR_REQUIRE((reinterpret_cast<uintptr_t>(cfg->rows_a) % 16) == 0 && (reinterpret_cast<uintptr_t>(cfg->queries) % 16) == 0);
  ESR_REQUIRE(cfg->rows_b == nullptr ? cfg->Db == 0 : (cfg->Db > 0 && cfg->idx_b != nullptr &&
                                                        (reinterpret_cast<uintptr_t>(cfg->rows_b) % 16) == 0));
  ESR_REQUIRE(cfg->ver == nullptr || cfg->rows_a1 != nullptr);
  ESR_REQUIRE(cfg->n_ctx_a >= 0 && cfg->n_ctx_a <= kMaxCtx && cfg->n_ctx_b >= 0 && cfg->n_ctx_b <= kMaxCtx);
  ESR_REQUIRE((cfg->n_ctx_a == 0 || (cfg->ctx_a && cfg->idx_a && cfg->max_over_queries)) &&
              (cfg->n_ctx_b == 0 || (cfg->ctx_b && cfg->idx_b && cfg->max_over_queries)));
  ESR_REQUIRE(cfg->N > 0 && cfg->N < ((int64_t)1 << 31));
  const int D4 = (cfg->Da + cfg->Db) / 4;
  TopkGeom g;
  if (!topk_geom(cfg->N, D4, cfg->T, cfg->max_over_queries, cfg->k, &g)) return cfg->k > cfg->N ? ESR_EINVAL : ESR_ENOTSUP;
  if (ws_bytes < esr_topk_workspace_bytes(cfg->N, cfg->Da + cfg->Db, cfg->T, cfg->max_over_queries, cfg->k)) return ESR_EWORKSPACE;
  cudaStream_t stream = static_cast<cudaStream_t>(stream_);
  TopkArgs a;
  a.A0 = cfg->rows_a;
  a.A1 = cfg->rows_a1 ? cfg->rows_a1 : cfg->rows_a;
  a.ver = cfg->ver;
  a.Bt = cfg->rows_b;
  a.idxA = cfg->idx_a;
  a.idxB = cfg->idx_b;
  a.modA = cfg->mod_a;
  a.DA4 = cfg->Da / 4;
  a.DB4 = cfg->Db / 4;
  a.N = cfg->N;
  a.Q = cfg->queries;
  a.T = cfg->T;
  a.maxq = cfg->max_over_queries ? 1 : 0;
  a.ctxA = cfg->ctx_a;
  a.ctxB = cfg->ctx_b;
  a.nA = cfg->n_ctx_a;
  a.nB = cfg->n_ctx_b;
  a.boost = cfg->boost;
  a.k = cfg->k;
  a.order = cfg->ties_high_index_first ? 1 : 0;
  a.cap = g.cap;
  a.tile_rows = g.tile_rows;
  a.slab = g.slab;
  a.cand = static_cast<uint64_t*>(ws);
  int rc;
  switch (g.G) {
    case 1: rc = launch_scan<1>(a, g, stream); break;
    case 2: rc = launch_scan<2>(a, g, stream); break;
    case 4: rc = launch_scan<4>(a, g, stream); break;
    case 8: rc = launch_scan<8>(a, g, stream); break;
    case 16: rc = launch_scan<16>(a, g, stream); break;
    default: rc = launch_scan<32>(a, g, stream); break;
  }
  if (rc != ESR_OK) return rc;
  static SmemOptIn configured_merge;
  if (g.smem_merge > 48 * 1024 && configured_merge.raise(g.smem_merge))
    ESR_CUDA(cudaFuncSetAttribute(k_topk_merge, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)g.smem_merge));
  k_topk_merge<<<g.lists, kThreads, g.smem_merge, stream>>>(a.cand, (int64_t)g.grid * cfg->k, cfg->k, g.cap, a.order, out_idx,
                                                            out_val);
  ESR_LAUNCH_CHECK();
  return ESR_OK;
}
