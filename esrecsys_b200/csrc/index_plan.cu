// Index plan of one batch: stable sort of slots by table row, unique rows, segment offsets,
// partner rows.  Integer bookkeeping only -- the bit-exact contract is oracle/index.py.
//
// Replaces the index handling XLA performs inside jnp.take and its scatter-add VJP
// (reference: nn.Embed lookups wikipedia/models.py:31-34; jax.value_and_grad
// wikipedia/train_cooccurence.py:86-87).
//
// The radix sort itself is cub::DeviceRadixSort (CUDA toolkit header library, compiled here for
// sm_100a; LSD radix sort is stable, which the determinism of the segment sums relies on).  The
// rest (iota, head flags, scan, compaction, partner lookup) is hand-written below.
#include <cub/block/block_scan.cuh>
#include <cub/device/device_radix_sort.cuh>

#include "esr_common.cuh"

namespace esr {
namespace {

constexpr int kTileThreads = 256;
constexpr int kItems = 8;
constexpr int kTile = kTileThreads * kItems;  // 2048 sorted slots per block

__global__ void k_iota(int32_t* v, int64_t n) {
  int64_t i = blockIdx.x * (int64_t)blockDim.x + threadIdx.x;
  if (i < n) v[i] = (int32_t)i;
}

__device__ __forceinline__ int head_flag(const int32_t* __restrict__ sk, int64_t p) {
  return p == 0 ? 1 : (sk[p] != sk[p - 1]);
}

// Real slots of the plan: the host-side capacity clamped by EsrPlan.n_valid (padding sorts to the end and is skipped).
__device__ __forceinline__ int64_t eff_slots(int64_t n, const int32_t* n_valid) {
  return n_valid ? min(n, (int64_t)__ldg(n_valid)) : n;
}

// Pass A: number of segment heads in each tile.
__global__ void __launch_bounds__(kTileThreads) k_head_count(const int32_t* __restrict__ sk, int64_t n_cap,
                                                            const int32_t* __restrict__ n_valid,
                                                            int32_t* __restrict__ tile_count) {
  __shared__ float red[32];
  const int64_t n = eff_slots(n_cap, n_valid);
  const int64_t base = (int64_t)blockIdx.x * kTile;
  int c = 0;
#pragma unroll
  for (int k = 0; k < kItems; ++k) {
    int64_t p = base + (int64_t)k * kTileThreads + threadIdx.x;
    if (p < n) c += head_flag(sk, p);
  }
  // integer block reduce through warp shuffles
  for (int o = 16; o > 0; o >>= 1) c += __shfl_xor_sync(FULL, c, o);
  int* ired = reinterpret_cast<int*>(red);
  if ((threadIdx.x & 31) == 0) ired[threadIdx.x >> 5] = c;
  __syncthreads();
  if (threadIdx.x < 32) {
    int x = threadIdx.x < (kTileThreads / 32) ? ired[threadIdx.x] : 0;
    for (int o = 16; o > 0; o >>= 1) x += __shfl_xor_sync(FULL, x, o);
    if (threadIdx.x == 0) tile_count[blockIdx.x] = x;
  }
}

// Pass B: exclusive scan of the tile counts (single block), total -> n_uniq, seg_off[total] = n.
__global__ void __launch_bounds__(1024) k_tile_scan(const int32_t* __restrict__ tile_count, int32_t n_tiles,
                                                    int32_t* __restrict__ tile_base, int32_t* __restrict__ n_uniq,
                                                    int32_t* __restrict__ seg_off, int64_t n_cap,
                                                    const int32_t* __restrict__ n_valid) {
  using Scan = cub::BlockScan<int, 1024>;
  const int64_t n = eff_slots(n_cap, n_valid);
  __shared__ typename Scan::TempStorage tmp;
  __shared__ int carry_s;
  if (threadIdx.x == 0) carry_s = 0;
  __syncthreads();
  for (int32_t b = 0; b < n_tiles; b += 1024) {
    int32_t t = b + threadIdx.x;
    int v = t < n_tiles ? tile_count[t] : 0;
    int ex, total;
    Scan(tmp).ExclusiveSum(v, ex, total);
    int carry = carry_s;
    if (t < n_tiles) tile_base[t] = carry + ex;
    __syncthreads();
    if (threadIdx.x == 0) carry_s = carry + total;
    __syncthreads();
  }
  if (threadIdx.x == 0) {
    *n_uniq = carry_s;
    seg_off[carry_s] = (int32_t)n;
  }
}

// Pass C: useg, uniq, seg_off; optionally partner rows.
__global__ void __launch_bounds__(kTileThreads) k_head_write(
    const int32_t* __restrict__ sk, const int32_t* __restrict__ perm, const int32_t* __restrict__ keys, int64_t n_cap,
    const int32_t* __restrict__ n_valid, const int32_t* __restrict__ tile_base, int32_t* __restrict__ useg,
    int32_t* __restrict__ uniq, int32_t* __restrict__ seg_off, int32_t* __restrict__ partner) {
  using Scan = cub::BlockScan<int, kTileThreads>;
  __shared__ typename Scan::TempStorage tmp;
  const int64_t n = eff_slots(n_cap, n_valid);
  // blocked arrangement: thread t owns kItems consecutive slots so the scan order is slot order
  const int64_t base = (int64_t)blockIdx.x * kTile + (int64_t)threadIdx.x * kItems;
  int f[kItems], key[kItems];
  int tsum = 0;
#pragma unroll
  for (int k = 0; k < kItems; ++k) {
    int64_t p = base + k;
    f[k] = 0;
    key[k] = 0;
    if (p < n) {
      key[k] = sk[p];
      f[k] = head_flag(sk, p);
    }
    tsum += f[k];
  }
  int ex;
  Scan(tmp).ExclusiveSum(tsum, ex);
  int run = tile_base[blockIdx.x] + ex;  // heads strictly before this thread's first slot
  const int64_t half = n_cap >> 1;  // slot layout [i ; j] of the CAPACITY, whatever part of it is real
#pragma unroll
  for (int k = 0; k < kItems; ++k) {
    int64_t p = base + k;
    if (p < n) {
      run += f[k];
      int u = run - 1;
      useg[p] = u;
      if (f[k]) {
        uniq[u] = key[k];
        seg_off[u] = (int32_t)p;
      }
      if (partner != nullptr) {
        int64_t s = perm[p];
        int64_t o = s < half ? s + half : s - half;
        partner[p] = keys[o];
      }
    }
  }
}

__global__ void k_remap(const int32_t* __restrict__ perm, const int32_t* __restrict__ useg, int32_t* __restrict__ out,
                        int64_t n) {
  int64_t p = blockIdx.x * (int64_t)blockDim.x + threadIdx.x;
  if (p < n) out[perm[p]] = useg[p];
}

__global__ void k_check_ids(const int32_t* __restrict__ ids, int64_t n, int64_t V, int32_t* __restrict__ n_bad) {
  int64_t i = blockIdx.x * (int64_t)blockDim.x + threadIdx.x;
  int bad = 0;
  if (i < n) bad = (ids[i] < 0 || (int64_t)ids[i] >= V);
  unsigned m = __ballot_sync(FULL, bad);
  if ((threadIdx.x & 31) == 0 && m) atomicAdd(n_bad, __popc(m));
}

size_t sort_temp_bound(int64_t n) {
  // Upper bound of cub::DeviceRadixSort::SortPairs temp storage for <int32,int32>: the onesweep
  // path keeps alternate key/value buffers plus per-pass histograms and look-back state.
  return align_up((size_t)n * 8, 256) + align_up((size_t)ceil_div(n, 1024) * 4 * 256 * 4 + (1 << 16), 256) + (1 << 20);
}

struct PlanWs {
  int32_t* iota;
  int32_t* tile_count;
  int32_t* tile_base;
  void* sort_tmp;
  size_t sort_tmp_bytes;
};

PlanWs carve(void* ws, int64_t n) {
  Carver c(ws);
  PlanWs w;
  w.iota = c.take<int32_t>(n);
  const int64_t tiles = ceil_div(n, kTile);
  w.tile_count = c.take<int32_t>(tiles + 1);
  w.tile_base = c.take<int32_t>(tiles + 1);
  w.sort_tmp_bytes = sort_temp_bound(n);
  w.sort_tmp = c.take<char>(w.sort_tmp_bytes);
  return w;
}

}  // namespace
}  // namespace esr

using namespace esr;

extern "C" size_t esr_plan_workspace_bytes(int64_t n_slots) {
  if (n_slots < 0) return 0;
  int64_t n = n_slots > 0 ? n_slots : 1;
  const int64_t tiles = ceil_div(n, kTile);
  return align_up((size_t)n * 4, 256) + 2 * align_up((size_t)(tiles + 1) * 4, 256) + align_up(sort_temp_bound(n), 256) + 1024;
}

extern "C" int esr_plan_build_i32(const EsrPlan* plan, void* ws, size_t ws_bytes, esr_stream_t stream_) {
  ESR_REQUIRE(plan != nullptr && plan->struct_size >= sizeof(EsrPlan));
  const int64_t n = plan->n_slots;
  ESR_REQUIRE(n >= 0 && n < (int64_t)1 << 31);
  ESR_REQUIRE(plan->sorted_keys && plan->perm && plan->useg && plan->uniq && plan->seg_off && plan->n_uniq);
  ESR_REQUIRE(plan->partner == nullptr || (n % 2) == 0);
  cudaStream_t stream = static_cast<cudaStream_t>(stream_);
  if (n == 0) {
    ESR_CUDA(cudaMemsetAsync(plan->n_uniq, 0, sizeof(int32_t), stream));
    ESR_CUDA(cudaMemsetAsync(plan->seg_off, 0, sizeof(int32_t), stream));
    return ESR_OK;
  }
  ESR_REQUIRE(plan->keys != nullptr && ws != nullptr);
  if (ws_bytes < esr_plan_workspace_bytes(n)) return ESR_EWORKSPACE;
  PlanWs w = carve(ws, n);

  k_iota<<<(unsigned)ceil_div(n, 256), 256, 0, stream>>>(w.iota, n);
  ESR_LAUNCH_CHECK();
  int end_bit = plan->key_bits > 0 && plan->key_bits <= 32 ? plan->key_bits : 32;
  if (end_bit == 32) end_bit = 31;  // ids are non-negative int32
  size_t need = 0;
  ESR_CUDA(cub::DeviceRadixSort::SortPairs(nullptr, need, plan->keys, plan->sorted_keys, w.iota, plan->perm, (int)n, 0,
                                           end_bit, stream));
  if (need > w.sort_tmp_bytes) return ESR_EWORKSPACE;
  size_t avail = w.sort_tmp_bytes;
  ESR_CUDA(cub::DeviceRadixSort::SortPairs(w.sort_tmp, avail, plan->keys, plan->sorted_keys, w.iota, plan->perm, (int)n,
                                           0, end_bit, stream));
  const int32_t tiles = (int32_t)ceil_div(n, kTile);
  k_head_count<<<tiles, kTileThreads, 0, stream>>>(plan->sorted_keys, n, plan->n_valid, w.tile_count);
  ESR_LAUNCH_CHECK();
  k_tile_scan<<<1, 1024, 0, stream>>>(w.tile_count, tiles, w.tile_base, plan->n_uniq, plan->seg_off, n, plan->n_valid);
  ESR_LAUNCH_CHECK();
  k_head_write<<<tiles, kTileThreads, 0, stream>>>(plan->sorted_keys, plan->perm, plan->keys, n, plan->n_valid, w.tile_base,
                                                   plan->useg, plan->uniq, plan->seg_off, plan->partner);
  ESR_LAUNCH_CHECK();
  return ESR_OK;
}

extern "C" int esr_plan_remap_ids_i32(const EsrPlan* plan, int32_t* ids_out, esr_stream_t stream_) {
  ESR_REQUIRE(plan != nullptr && plan->struct_size >= sizeof(EsrPlan) && ids_out != nullptr);
  const int64_t n = plan->n_slots;
  if (n == 0) return ESR_OK;
  k_remap<<<(unsigned)ceil_div(n, 256), 256, 0, static_cast<cudaStream_t>(stream_)>>>(plan->perm, plan->useg, ids_out, n);
  ESR_LAUNCH_CHECK();
  return ESR_OK;
}

extern "C" int esr_check_ids_i32(const int32_t* ids, int64_t n, int64_t V, int32_t* n_bad, esr_stream_t stream_) {
  ESR_REQUIRE(n_bad != nullptr && n >= 0 && (ids != nullptr || n == 0));
  cudaStream_t stream = static_cast<cudaStream_t>(stream_);
  ESR_CUDA(cudaMemsetAsync(n_bad, 0, sizeof(int32_t), stream));
  if (n == 0) return ESR_OK;
  k_check_ids<<<(unsigned)ceil_div(n, 256), 256, 0, stream>>>(ids, n, V, n_bad);
  ESR_LAUNCH_CHECK();
  return ESR_OK;
}
