// Row-sharded table: routing bookkeeping and owner-side gradient merge.
//
// The reference has no distributed code; BASELINE.json's north star row-shards the
// table across the GPUs of one box.  Ownership is CYCLIC (owner = row mod n, local = row div n)
// because GloVe row ids are frequency ranks (wikipedia/make_dictionary.py:113-116): block sharding
// would put every hot row on rank 0.  Bit-exact contract of the integer work: oracle/index.py
// (route_plan, local_row, global_row).
//
//  esr_route_plan_i32       unique rows of a rank's batch -> stable bucket by owner: order, owner-local
//                           ids in bucket order (payload of the id all-to-all), per-owner counts
//  esr_plan_compact_i32     re-express a batch's plan in unique-row indices, so the step can run on
//                           the compact table of fetched rows without sorting again
//  esr_gather_scalar_f32    out[k] = src[ids[k]]  (bias lookups)
//  esr_permute_rows_f32     out[k,:] = src[idx[k],:] / out[idx[k],:] = src[k,:]
//  esr_segment_sum_rows_f32 owner side: per-unique-row sum of the gradients received from all ranks,
//                           in stable sorted order (deterministic)
#include <cub/device/device_radix_sort.cuh>

#include "esr_common.cuh"

namespace esr {
namespace {

constexpr int kThreads = 256;
constexpr int kMaxRanks = 64;

__global__ void __launch_bounds__(kThreads) k_owner_keys(const int32_t* __restrict__ uniq,
                                                         const int32_t* __restrict__ n_uniq, int64_t cap, int n_ranks,
                                                         int32_t* __restrict__ owner, int32_t* __restrict__ iota) {
  const int64_t u = blockIdx.x * (int64_t)kThreads + threadIdx.x;
  if (u >= cap) return;
  // entries beyond n_uniq sort to the end (owner = n_ranks) and are never counted
  owner[u] = u < *n_uniq ? uniq[u] % n_ranks : n_ranks;
  iota[u] = (int32_t)u;
}

__global__ void __launch_bounds__(kThreads) k_route_finish(const int32_t* __restrict__ uniq,
                                                           const int32_t* __restrict__ n_uniq,
                                                           const int32_t* __restrict__ order, int64_t cap, int n_ranks,
                                                           int32_t* __restrict__ send_local,
                                                           int32_t* __restrict__ send_counts,
                                                           int32_t* __restrict__ inv_order) {
  __shared__ int cnt[kMaxRanks];
  if (threadIdx.x < kMaxRanks) cnt[threadIdx.x] = 0;
  __syncthreads();
  const int64_t k = blockIdx.x * (int64_t)kThreads + threadIdx.x;
  if (k < *n_uniq) {
    const int32_t u = order[k];
    const int32_t row = uniq[u];
    send_local[k] = row / n_ranks;
    if (inv_order) inv_order[u] = (int32_t)k;
    atomicAdd(&cnt[row % n_ranks], 1);  // integer counts: order-independent
  }
  __syncthreads();
  if (threadIdx.x < n_ranks && cnt[threadIdx.x]) atomicAdd(send_counts + threadIdx.x, cnt[threadIdx.x]);
}

__global__ void __launch_bounds__(kThreads) k_plan_compact(const int32_t* __restrict__ perm,
                                                           const int32_t* __restrict__ useg, int64_t n_cap,
                                                           const int32_t* __restrict__ n_valid,
                                                           int32_t* __restrict__ slot_u /* [n] scratch */) {
  const int64_t n = n_valid ? min(n_cap, (int64_t)__ldg(n_valid)) : n_cap;
  const int64_t p = blockIdx.x * (int64_t)kThreads + threadIdx.x;
  if (p < n) slot_u[perm[p]] = useg[p];
}

__global__ void __launch_bounds__(kThreads) k_plan_compact2(const int32_t* __restrict__ perm,
                                                            const int32_t* __restrict__ useg,
                                                            const int32_t* __restrict__ slot_u, int64_t n_cap,
                                                            const int32_t* __restrict__ n_valid,
                                                            int32_t* __restrict__ sorted_keys,
                                                            int32_t* __restrict__ partner, int32_t* __restrict__ uniq) {
  const int64_t n = n_valid ? min(n_cap, (int64_t)__ldg(n_valid)) : n_cap;
  const int64_t p = blockIdx.x * (int64_t)kThreads + threadIdx.x;
  if (p >= n) return;
  const int64_t half = n_cap >> 1;  // slot layout [i ; j] of the capacity
  const int64_t s = perm[p];
  const int32_t u = useg[p];
  sorted_keys[p] = u;
  partner[p] = slot_u[s < half ? s + half : s - half];
  uniq[u] = u;  // every slot of a segment writes the same value
}

// Owner-aware form (unified table of the owner-routed sharded step): a unique row this rank OWNS is addressed where it
// lives in the shard (row / n), a fetched one in the fetch region behind the shard (base + u) -- so local rows are never
// copied.  Equal rows still get equal keys, which is all the row pass needs from a key (address + segment equality).
__global__ void __launch_bounds__(kThreads) k_plan_compact_owner(const int32_t* __restrict__ perm,
                                                                 const int32_t* __restrict__ useg,
                                                                 const int32_t* __restrict__ slot_u,
                                                                 const int32_t* __restrict__ uniq, int64_t n_cap,
                                                                 const int32_t* __restrict__ n_valid, int n_ranks, int me,
                                                                 int32_t base, int32_t* __restrict__ sorted_keys,
                                                                 int32_t* __restrict__ partner) {
  const int64_t n = n_valid ? min(n_cap, (int64_t)__ldg(n_valid)) : n_cap;
  const int64_t p = blockIdx.x * (int64_t)kThreads + threadIdx.x;
  if (p >= n) return;
  const int64_t half = n_cap >> 1;
  const int64_t s = perm[p];
  const int32_t u = useg[p];
  const int32_t u2 = slot_u[s < half ? s + half : s - half];
  const int32_t r = uniq[u], r2 = uniq[u2];
  sorted_keys[p] = (r % n_ranks == me) ? r / n_ranks : base + u;
  partner[p] = (r2 % n_ranks == me) ? r2 / n_ranks : base + u2;
}

__global__ void __launch_bounds__(kThreads) k_gather_scalar(const float* __restrict__ src, const int32_t* __restrict__ ids,
                                                            int64_t n, float* __restrict__ out) {
  const int64_t k = blockIdx.x * (int64_t)kThreads + threadIdx.x;
  if (k < n) out[k] = src[ids[k]];
}

// gather: out[k] = src[idx[k]]; scatter: out[idx[k]] = src[k]; rows beyond *n_valid are skipped
template <int TPR>
__global__ void __launch_bounds__(kThreads) k_permute_rows(const float4* __restrict__ src, const int32_t* __restrict__ idx,
                                                           const int32_t* __restrict__ n_valid, int64_t cap, int D4,
                                                           int scatter, float4* __restrict__ out) {
  const int lane = threadIdx.x % TPR;
  const int64_t k = (blockIdx.x * (int64_t)kThreads + threadIdx.x) / TPR;
  const int64_t n = n_valid ? min((int64_t)*n_valid, cap) : cap;
  if (k >= n) return;
  const int64_t j = idx[k];
  const float4* s = src + (scatter ? k : j) * D4;
  float4* d = out + (scatter ? j : k) * D4;
  for (int c = lane; c < D4; c += TPR) d[c] = ld_stream(s + c);
}

// Sum of the rows g_in[perm[p]], p in [a, b), column c of this lane, in slot order; kSegBatch rows in flight (the adds
// stay in slot order: batching changes the latency, not the result).
constexpr int kSegBatch = 8;
__device__ __forceinline__ float4 sum_slots(const int32_t* __restrict__ perm, const float4* __restrict__ g_in, int D4, int c,
                                            int a, int b) {
  float4 acc = f4_zero();
  for (int p = a; p < b; p += kSegBatch) {
    int64_t r[kSegBatch];
#pragma unroll
    for (int q = 0; q < kSegBatch; ++q) r[q] = p + q < b ? perm[p + q] : -1;
    float4 v[kSegBatch];
#pragma unroll
    for (int q = 0; q < kSegBatch; ++q) v[q] = r[q] >= 0 ? ld_stream(g_in + r[q] * D4 + c) : f4_zero();
#pragma unroll
    for (int q = 0; q < kSegBatch; ++q)
      if (r[q] >= 0) f4_add(acc, v[q]);
  }
  return acc;
}

__device__ __forceinline__ float sum_slots_scalar(const int32_t* __restrict__ perm, const float* __restrict__ gb_in, int a,
                                                  int b) {
  float acc = 0.f;
  for (int p = a; p < b; p += kSegBatch) {
    float v[kSegBatch];
#pragma unroll
    for (int q = 0; q < kSegBatch; ++q) v[q] = p + q < b ? gb_in[perm[p + q]] : 0.f;
#pragma unroll
    for (int q = 0; q < kSegBatch; ++q)
      if (p + q < b) acc += v[q];
  }
  return acc;
}

// One group of TPR lanes per unique row: sums the rows g_in[perm[p]] over the row's sorted slots,
// in slot order (fixed => deterministic).
template <int TPR>
__global__ void __launch_bounds__(kThreads) k_segment_sum(const int32_t* __restrict__ perm,
                                                          const int32_t* __restrict__ seg_off,
                                                          const int32_t* __restrict__ n_uniq, int64_t cap, int D4,
                                                          const float4* __restrict__ g_in, const float* __restrict__ gb_in,
                                                          float4* __restrict__ g_out, float* __restrict__ gb_out) {
  const int lane = threadIdx.x % TPR;
  const int64_t u = (blockIdx.x * (int64_t)kThreads + threadIdx.x) / TPR;
  if (u >= cap || u >= *n_uniq) return;
  const int s0 = seg_off[u], s1 = seg_off[u + 1];
  for (int c = lane; c < D4; c += TPR) g_out[u * D4 + c] = sum_slots(perm, g_in, D4, c, s0, s1);
  if (lane == 0 && gb_in != nullptr) gb_out[u] = sum_slots_scalar(perm, gb_in, s0, s1);
}

// Rows of D >= 128: a warp per unique row, and the block's 8 warps TOGETHER on every row of theirs with more than
// kLongSeg slots (an in-batch id stream is Zipf: the hottest id owns 7 % of the slots, and one warp walking that segment
// alone was 230 us of a 0.76 ms two-tower step).  A long row is cut into 8 equal pieces, one per warp, summed in slot
// order and added up in piece order: a fixed tree, so the result does not depend on scheduling.
constexpr int kLongSeg = 64;
__global__ void __launch_bounds__(kThreads) k_segment_sum_coop(const int32_t* __restrict__ perm,
                                                               const int32_t* __restrict__ seg_off,
                                                               const int32_t* __restrict__ n_uniq, int64_t cap, int D4,
                                                               const float4* __restrict__ g_in,
                                                               const float* __restrict__ gb_in, float4* __restrict__ g_out,
                                                               float* __restrict__ gb_out) {
  constexpr int kWarps = kThreads / 32;
  extern __shared__ float4 part[];  // [kWarps][D4]
  __shared__ float part_b[kWarps];
  __shared__ int long_u[kWarps];
  const int lane = threadIdx.x & 31, w = threadIdx.x >> 5;
  const int64_t nu = min((int64_t)*n_uniq, cap);
  const int64_t u = (int64_t)blockIdx.x * kWarps + w;
  int s0 = 0, s1 = 0;
  if (u < nu) {
    s0 = seg_off[u];
    s1 = seg_off[u + 1];
  }
  const bool is_long = s1 - s0 > kLongSeg;
  if (lane == 0) long_u[w] = is_long ? 1 : 0;
  if (u < nu && !is_long) {
    for (int c = lane; c < D4; c += 32) g_out[u * D4 + c] = sum_slots(perm, g_in, D4, c, s0, s1);
    if (lane == 0 && gb_in != nullptr) gb_out[u] = sum_slots_scalar(perm, gb_in, s0, s1);
  }
  __syncthreads();
  for (int x = 0; x < kWarps; ++x) {
    if (!long_u[x]) continue;  // block-uniform
    const int64_t ux = (int64_t)blockIdx.x * kWarps + x;
    const int a = seg_off[ux], b = seg_off[ux + 1];
    const int piece = (b - a + kWarps - 1) / kWarps;
    const int pa = min(b, a + w * piece), pb = min(b, pa + piece);
    for (int c = lane; c < D4; c += 32) part[w * D4 + c] = sum_slots(perm, g_in, D4, c, pa, pb);
    if (lane == 0 && gb_in != nullptr) part_b[w] = sum_slots_scalar(perm, gb_in, pa, pb);
    __syncthreads();
    for (int c = threadIdx.x; c < D4; c += kThreads) {
      float4 acc = part[c];
#pragma unroll
      for (int y = 1; y < kWarps; ++y) f4_add(acc, part[y * D4 + c]);
      g_out[ux * D4 + c] = acc;
    }
    if (threadIdx.x == 0 && gb_in != nullptr) {
      float acc = part_b[0];
#pragma unroll
      for (int y = 1; y < kWarps; ++y) acc += part_b[y];
      gb_out[ux] = acc;
    }
    __syncthreads();
  }
}

int tpr_for(int D4) {
  int t = 1;
  while (t < D4 && t < 32) t <<= 1;
  return t;
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

size_t route_sort_bound(int64_t n) { return align_up((size_t)n * 8, 256) + (1 << 20); }

}  // namespace
}  // namespace esr

namespace esr {
int route_plan_small(const int32_t* uniq, const int32_t* n_uniq, int64_t cap, int32_t n_ranks, int32_t* order,
                     int32_t* send_local, int32_t* send_counts, int32_t* inv_order, void* ws, cudaStream_t stream);  // peer_ops.cu
}

using namespace esr;

extern "C" size_t esr_route_workspace_bytes(int64_t cap) {
  if (cap < 0) return 0;
  const int64_t n = cap > 0 ? cap : 1;
  return 2 * align_up((size_t)n * 4, 256) + align_up(route_sort_bound(n), 256) + 1024;
}

extern "C" int esr_route_plan_i32(const int32_t* uniq, const int32_t* n_uniq, int64_t cap, int32_t n_ranks,
                                  int32_t* order, int32_t* send_local, int32_t* send_counts, int32_t* inv_order, void* ws,
                                  size_t ws_bytes, esr_stream_t stream_) {
  ESR_RANGE("esr_route_plan_i32");
  ESR_REQUIRE(uniq && n_uniq && order && send_local && send_counts && cap >= 0 && n_ranks >= 1 && n_ranks <= kMaxRanks);
  cudaStream_t stream = static_cast<cudaStream_t>(stream_);
  ESR_CUDA(cudaMemsetAsync(send_counts, 0, sizeof(int32_t) * n_ranks, stream));
  if (cap == 0) return ESR_OK;
  ESR_REQUIRE(ws != nullptr && cap < ((int64_t)1 << 31));
  if (ws_bytes < esr_route_workspace_bytes(cap)) return ESR_EWORKSPACE;
  if (n_ranks <= ESR_MAX_PEERS)  // stable partition of the (sorted) unique rows: no sort needed
    return route_plan_small(uniq, n_uniq, cap, n_ranks, order, send_local, send_counts, inv_order, ws, stream);
  Carver c(ws);
  int32_t* owner = c.take<int32_t>(cap);
  int32_t* iota = c.take<int32_t>(cap);
  size_t tmp_bytes = route_sort_bound(cap);
  void* tmp = c.take<char>(tmp_bytes);
  const unsigned grid = (unsigned)ceil_div(cap, kThreads);
  k_owner_keys<<<grid, kThreads, 0, stream>>>(uniq, n_uniq, cap, n_ranks, owner, iota);
  ESR_LAUNCH_CHECK();
  int bits = 1;
  while ((1 << bits) <= n_ranks) ++bits;  // owners 0..n_ranks (n_ranks = padding)
  size_t need = 0;
  // sorted owners are not needed: reuse send_local as the key output, then overwrite it
  ESR_CUDA(cub::DeviceRadixSort::SortPairs(nullptr, need, owner, send_local, iota, order, (int)cap, 0, bits, stream));
  if (need > tmp_bytes) return ESR_EWORKSPACE;
  ESR_CUDA(cub::DeviceRadixSort::SortPairs(tmp, tmp_bytes, owner, send_local, iota, order, (int)cap, 0, bits, stream));
  k_route_finish<<<grid, kThreads, 0, stream>>>(uniq, n_uniq, order, cap, n_ranks, send_local, send_counts, inv_order);
  ESR_LAUNCH_CHECK();
  return ESR_OK;
}

extern "C" int esr_plan_compact_i32(const EsrPlan* plan, int32_t* sorted_keys, int32_t* partner, int32_t* uniq,
                                    int32_t* scratch, esr_stream_t stream_) {
  ESR_RANGE("esr_plan_compact_i32");
  ESR_REQUIRE(plan && plan->struct_size >= sizeof(EsrPlan) && sorted_keys && partner && uniq && scratch);
  const int64_t n = plan->n_slots;
  ESR_REQUIRE(n >= 0 && (n % 2) == 0);
  if (n == 0) return ESR_OK;
  ESR_REQUIRE(plan->perm && plan->useg);
  cudaStream_t stream = static_cast<cudaStream_t>(stream_);
  const unsigned grid = (unsigned)ceil_div(n, kThreads);
  k_plan_compact<<<grid, kThreads, 0, stream>>>(plan->perm, plan->useg, n, plan->n_valid, scratch);
  ESR_LAUNCH_CHECK();
  k_plan_compact2<<<grid, kThreads, 0, stream>>>(plan->perm, plan->useg, scratch, n, plan->n_valid, sorted_keys, partner,
                                                 uniq);
  ESR_LAUNCH_CHECK();
  return ESR_OK;
}

extern "C" int esr_plan_compact_owner_i32(const EsrPlan* plan, int32_t n_ranks, int32_t me, int32_t base, int32_t* sorted_keys,
                                          int32_t* partner, int32_t* scratch, esr_stream_t stream_) {
  ESR_RANGE("esr_plan_compact_owner_i32");
  ESR_REQUIRE(plan && plan->struct_size >= sizeof(EsrPlan) && sorted_keys && partner && scratch);
  ESR_REQUIRE(n_ranks >= 1 && me >= 0 && me < n_ranks && base >= 0);
  const int64_t n = plan->n_slots;
  ESR_REQUIRE(n >= 0 && (n % 2) == 0 && (int64_t)base + n < ((int64_t)1 << 30));
  if (n == 0) return ESR_OK;
  ESR_REQUIRE(plan->perm && plan->useg && plan->uniq);
  cudaStream_t stream = static_cast<cudaStream_t>(stream_);
  const unsigned grid = (unsigned)ceil_div(n, kThreads);
  k_plan_compact<<<grid, kThreads, 0, stream>>>(plan->perm, plan->useg, n, plan->n_valid, scratch);
  ESR_LAUNCH_CHECK();
  k_plan_compact_owner<<<grid, kThreads, 0, stream>>>(plan->perm, plan->useg, scratch, plan->uniq, n, plan->n_valid, n_ranks, me,
                                                      base, sorted_keys, partner);
  ESR_LAUNCH_CHECK();
  return ESR_OK;
}

extern "C" int esr_gather_scalar_f32(const float* src, const int32_t* ids, int64_t n, float* out, esr_stream_t stream_) {
  ESR_REQUIRE(n >= 0 && (n == 0 || (src && ids && out)));
  if (n == 0) return ESR_OK;
  k_gather_scalar<<<(unsigned)ceil_div(n, kThreads), kThreads, 0, static_cast<cudaStream_t>(stream_)>>>(src, ids, n, out);
  ESR_LAUNCH_CHECK();
  return ESR_OK;
}

extern "C" int esr_permute_rows_f32(const float* src, const int32_t* idx, const int32_t* n_valid, int64_t cap, int32_t D,
                                    int32_t scatter, float* out, esr_stream_t stream_) {
  ESR_REQUIRE(cap >= 0 && D > 0 && (D % 4) == 0 && (cap == 0 || (src && idx && out)));
  ESR_REQUIRE(((reinterpret_cast<uintptr_t>(src) | reinterpret_cast<uintptr_t>(out)) % 16) == 0);
  if (cap == 0) return ESR_OK;
  const int D4 = D / 4;
  const int tpr = tpr_for(D4);
  const unsigned grid = (unsigned)ceil_div(cap * tpr, kThreads);
  ESR_DISPATCH_TPR(tpr, (k_permute_rows<TPR><<<grid, kThreads, 0, static_cast<cudaStream_t>(stream_)>>>(
                            reinterpret_cast<const float4*>(src), idx, n_valid, cap, D4, scatter,
                            reinterpret_cast<float4*>(out))));
  ESR_LAUNCH_CHECK();
  return ESR_OK;
}

extern "C" int esr_segment_sum_rows_f32(const EsrPlan* plan, int32_t D, const float* g_in, const float* gb_in, float* g_out,
                                        float* gb_out, esr_stream_t stream_) {
  ESR_RANGE("esr_segment_sum_rows_f32");
  ESR_REQUIRE(plan && plan->struct_size >= sizeof(EsrPlan) && D > 0 && (D % 4) == 0);
  const int64_t n = plan->n_slots;
  if (n == 0) return ESR_OK;
  ESR_REQUIRE(plan->perm && plan->seg_off && plan->n_uniq && g_in && g_out && (gb_in == nullptr || gb_out != nullptr));
  ESR_REQUIRE(((reinterpret_cast<uintptr_t>(g_in) | reinterpret_cast<uintptr_t>(g_out)) % 16) == 0);
  const int D4 = D / 4;
  const int tpr = tpr_for(D4);
  const unsigned grid = (unsigned)ceil_div(n * tpr, kThreads);
  const size_t coop_smem = (size_t)(kThreads / 32) * D4 * sizeof(float4);
  if (tpr == 32 && coop_smem <= 40 * 1024) {
    k_segment_sum_coop<<<grid, kThreads, coop_smem, static_cast<cudaStream_t>(stream_)>>>(
        plan->perm, plan->seg_off, plan->n_uniq, n, D4, reinterpret_cast<const float4*>(g_in), gb_in,
        reinterpret_cast<float4*>(g_out), gb_out);
    ESR_LAUNCH_CHECK();
    return ESR_OK;
  }
  ESR_DISPATCH_TPR(tpr, (k_segment_sum<TPR><<<grid, kThreads, 0, static_cast<cudaStream_t>(stream_)>>>(
                            plan->perm, plan->seg_off, plan->n_uniq, n, D4, reinterpret_cast<const float4*>(g_in), gb_in,
                            reinterpret_cast<float4*>(g_out), gb_out)));
  ESR_LAUNCH_CHECK();
  return ESR_OK;
}
