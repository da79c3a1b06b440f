// Per-query ranking of a (V, T) score matrix: the retrieval step that follows training in every
// reference script:
//   find_knn / dump_knn  wikipedia/train_cooccurence.py:91-97, :114-126   jnp.argsort(scores, axis=0), top 10 from the tail
//   eval_step            spotify/train_spotify.py:113-131                 jax.lax.top_k(neg_affinity, 500)
//   find_top_k           pinterest/make_recommendations.py:49-65          jax.lax.top_k(scores, k)
// Ordering contract (oracle.glove.find_knn / top_k): ascending argsort is STABLE (ties keep index order, as
// jnp.argsort); descending top-k lists ties in index order too (as jax.lax.top_k).  An LSD radix sort of
// (score, index) pairs gives both; the sort is cub::DeviceRadixSort compiled into libesr, the column
// extraction / result scatter around it are the kernels below.  (-0.0 sorts before +0.0; the reference treats
// them as equal.  Scores of exactly -0.0 do not occur for dot products of generic rows.)
#include <cub/device/device_radix_sort.cuh>

#include "esr_common.cuh"

namespace esr {
namespace {

constexpr int kThreads = 256;

__global__ void __launch_bounds__(kThreads) k_col_extract(const float* __restrict__ scores, int64_t V, int32_t T, int32_t t,
                                                          float* __restrict__ keys, int32_t* __restrict__ idx) {
  const int64_t v = blockIdx.x * (int64_t)kThreads + threadIdx.x;
  if (v < V) {
    keys[v] = scores[v * T + t];
    idx[v] = (int32_t)v;
  }
}

__global__ void __launch_bounds__(kThreads) k_col_write(const float* __restrict__ keys, const int32_t* __restrict__ idx,
                                                        int64_t k, int32_t T, int32_t t, int32_t* __restrict__ out_idx,
                                                        float* __restrict__ out_val) {
  const int64_t r = blockIdx.x * (int64_t)kThreads + threadIdx.x;
  if (r < k) {
    out_idx[r * T + t] = idx[r];
    if (out_val) out_val[r * T + t] = keys[r];
  }
}

size_t sort_tmp_bound(int64_t V) {
  return align_up((size_t)V * 8, 256) + align_up((size_t)ceil_div(V, 1024) * 4 * 256 * 4 + (1 << 16), 256) + (1 << 20);
}

}  // namespace
}  // namespace esr

using namespace esr;

extern "C" size_t esr_sort_cols_workspace_bytes(int64_t V) {
  if (V <= 0) return 0;
  return 4 * align_up((size_t)V * 4, 256) + align_up(sort_tmp_bound(V), 256) + 1024;
}

extern "C" int esr_sort_cols_f32(const float* scores, int64_t V, int32_t T, int32_t descending, int64_t k, int32_t* out_idx,
                                 float* out_val, void* ws, size_t ws_bytes, esr_stream_t stream_) {
  ESR_REQUIRE(scores && out_idx && ws && V > 0 && V < ((int64_t)1 << 31) && T > 0 && k > 0 && k <= V);
  if (ws_bytes < esr_sort_cols_workspace_bytes(V)) return ESR_EWORKSPACE;
  cudaStream_t stream = static_cast<cudaStream_t>(stream_);
  Carver c(ws);
  float* keys_in = c.take<float>(V);
  float* keys_out = c.take<float>(V);
  int32_t* idx_in = c.take<int32_t>(V);
  int32_t* idx_out = c.take<int32_t>(V);
  const size_t tmp_bytes = sort_tmp_bound(V);
  void* tmp = c.take<char>(tmp_bytes);
  size_t need = 0;
  ESR_CUDA(cub::DeviceRadixSort::SortPairs(nullptr, need, keys_in, keys_out, idx_in, idx_out, (int)V, 0, 32, stream));
  if (need > tmp_bytes) return ESR_EWORKSPACE;
  const unsigned gv = (unsigned)ceil_div(V, kThreads), gk = (unsigned)ceil_div(k, kThreads);
  for (int32_t t = 0; t < T; ++t) {
    k_col_extract<<<gv, kThreads, 0, stream>>>(scores, V, T, t, keys_in, idx_in);
    ESR_LAUNCH_CHECK();
    size_t avail = tmp_bytes;
    if (descending)
      ESR_CUDA(cub::DeviceRadixSort::SortPairsDescending(tmp, avail, keys_in, keys_out, idx_in, idx_out, (int)V, 0, 32, stream));
    else
      ESR_CUDA(cub::DeviceRadixSort::SortPairs(tmp, avail, keys_in, keys_out, idx_in, idx_out, (int)V, 0, 32, stream));
    k_col_write<<<gk, kThreads, 0, stream>>>(keys_out, idx_out, k, T, t, out_idx, out_val);
    ESR_LAUNCH_CHECK();
  }
  return ESR_OK;
}
