// Row-sharded table over NVLink PEER MEMORY: the exchange steps of the sharded GloveE step done by
// kernels that load directly from the other GPUs' HBM (pointers from a symmetric-memory rendezvous),
// instead of NCCL all-to-alls with host-known split sizes.  No id exchange, no send buffers, no host
// synchronisation: every size is read on the device.
//
//  esr_peer_gather_f32        rank r reads the current value of its batch's unique rows straight from
//                             their owners' shards:  out[u,:] = shard[owner(uniq[u])][local(uniq[u]),:]
//                             (the "index all-to-all + row all-to-all" of SURVEY.md 8(e) in one kernel)
//  esr_peer_pull_ids_i32      owner side: copy, from every source rank's published route plan, the
//                             owner-local ids destined to me into one local array (+ per-source
//                             offsets); tiny (4 B per row) so the merge can binary-search locally
//  esr_peer_merge_adagrad_f32 owner side: for every received (source, row) entry that is the FIRST
//                             source naming that row, sum the row's gradients over all sources in
//                             source order (peer loads of the 512-byte gradient rows), then
//                             optax.adagrad on the local shard row.  Deterministic; replaces the
//                             gradient all-to-all + sort + segment-sum.
// Ownership is cyclic (owner = row % n, local = row / n).  The caller separates the phases with
// device barriers (all fetches done before any update; all updates done before the next fetch).
#include "esr_common.cuh"

namespace esr {
namespace {

constexpr int kThreads = 256;

struct PeerPtrs {
  const void* p[ESR_MAX_PEERS];
};

// owner / local row of a global row id under cyclic sharding; shift/mask when n is a power of two
struct Cyclic {
  int n, shift;  // shift >= 0: n == 1 << shift
  __device__ __forceinline__ int owner(int32_t row) const { return shift >= 0 ? (row & (n - 1)) : row % n; }
  __device__ __forceinline__ int64_t local(int32_t row) const { return shift >= 0 ? (row >> shift) : row / n; }
};

// Persistent grid: groups of TPR lanes stride over quads of unique rows (sizes are only known on the
// device, so a capacity-sized grid would launch mostly empty blocks).
template <int TPR, int ROWS>
__global__ void __launch_bounds__(kThreads) k_peer_gather(const __grid_constant__ PeerPtrs rows, const __grid_constant__ PeerPtrs bias, const int32_t* __restrict__ uniq,
                                                          const int32_t* __restrict__ n_uniq, int64_t cap, Cyclic cyc,
                                                          int D4, float4* __restrict__ out, float* __restrict__ out_bias) {
  const int lane = threadIdx.x % TPR;
  const int64_t n = min((int64_t)*n_uniq, cap);
  const int64_t groups = (int64_t)gridDim.x * (kThreads / TPR);
  for (int64_t first = ((blockIdx.x * (int64_t)kThreads + threadIdx.x) / TPR) * ROWS; first < n; first += groups * ROWS) {
    const float4* src[ROWS];
#pragma unroll
    for (int r = 0; r < ROWS; ++r) {
      const int64_t u = first + r;
      src[r] = nullptr;
      if (u < n) {
        const int32_t row = uniq[u];
        const int owner = cyc.owner(row);
        const int64_t local = cyc.local(row);
        src[r] = reinterpret_cast<const float4*>(rows.p[owner]) + local * D4;
        if (lane == 0) out_bias[u] = reinterpret_cast<const float*>(bias.p[owner])[local];
      }
    }
    for (int c = lane; c < D4; c += TPR) {
      float4 v[ROWS];
#pragma unroll
      for (int r = 0; r < ROWS; ++r)
        if (src[r]) v[r] = src[r][c];  // plain ld.global: peer addresses bypass L2, nothing to hint
#pragma unroll
      for (int r = 0; r < ROWS; ++r)
        if (src[r]) st_stream(out + (first + r) * D4 + c, v[r]);
    }
  }
}

// src_meta[s] = {offset of source s in recv_ids, count, displacement inside source s's bucket list}
// slot_map[s][x] = position of owner-local row x in source s's list (-1: source s does not name x)
__global__ void __launch_bounds__(kThreads) k_peer_pull_ids(const __grid_constant__ PeerPtrs counts, const __grid_constant__ PeerPtrs send_local, int n_ranks, int me,
                                                            int64_t recv_cap, int32_t* __restrict__ recv_ids,
                                                            int32_t* __restrict__ src_meta, int32_t* __restrict__ slot_map,
                                                            int64_t map_stride) {
  __shared__ int off[ESR_MAX_PEERS + 1], cnt[ESR_MAX_PEERS], dsp[ESR_MAX_PEERS];
  __shared__ int cm[ESR_MAX_PEERS][ESR_MAX_PEERS];
  if ((int)threadIdx.x < n_ranks * n_ranks) {  // all n*n peer loads in flight at once
    const int s = threadIdx.x / n_ranks, q = threadIdx.x % n_ranks;
    cm[s][q] = reinterpret_cast<const int32_t*>(counts.p[s])[q];
  }
  __syncthreads();
  if (threadIdx.x == 0) {
    int o = 0;
    for (int s = 0; s < n_ranks; ++s) {
      int d = 0;
      for (int q = 0; q < me; ++q) d += cm[s][q];
      off[s] = o;
      cnt[s] = cm[s][me];
      dsp[s] = d;
      o += cm[s][me];
    }
    off[n_ranks] = o;
  }
  __syncthreads();
  if (blockIdx.x == 0 && (int)threadIdx.x <= n_ranks) {
    const int s = threadIdx.x;
    if (s < n_ranks) {
      src_meta[s * 3 + 0] = off[s];
      src_meta[s * 3 + 1] = cnt[s];
      src_meta[s * 3 + 2] = dsp[s];
    } else {
      src_meta[n_ranks * 3] = min((int64_t)off[n_ranks], recv_cap);  // total received
      src_meta[n_ranks * 3 + 1] = 0;                                 // owner entries (counted by k_peer_resolve)
    }
  }
  const int64_t total = min((int64_t)off[n_ranks], recv_cap);
  for (int64_t k = blockIdx.x * (int64_t)kThreads + threadIdx.x; k < total; k += (int64_t)gridDim.x * kThreads) {
    int s = 0;
    while (s + 1 < n_ranks && k >= off[s + 1]) ++s;
    const int pos = (int)(k - off[s]);
    const int32_t x = reinterpret_cast<const int32_t*>(send_local.p[s])[dsp[s] + pos];
    recv_ids[k] = x;
    slot_map[s * map_stride + x] = pos;
  }
}

__global__ void __launch_bounds__(kThreads) k_peer_clear_map(const int32_t* __restrict__ recv_ids,
                                                             const int32_t* __restrict__ src_meta, int n_ranks,
                                                             int32_t* __restrict__ slot_map, int64_t map_stride) {
  const int64_t total = src_meta[n_ranks * 3];
  for (int64_t k = blockIdx.x * (int64_t)kThreads + threadIdx.x; k < total; k += (int64_t)gridDim.x * kThreads) {
    int s = 0;
    while (s + 1 < n_ranks && k >= src_meta[(s + 1) * 3]) ++s;
    slot_map[s * map_stride + recv_ids[k]] = -1;
  }
}

// One thread per received (source, row) entry: resolve, through slot_map, where every source keeps
// that row's gradient.  The entry of the FIRST source naming a row owns it: desc[k][q] = row index in
// my inbox of source q's gradient (offset_q + position), -1 if source q does not name the row; entries that do
// not own their row get all -1.  Splitting this off keeps the heavy kernel below free of dependent
// 4-byte lookups, so its row loads are issued back to back.
__global__ void __launch_bounds__(kThreads) k_peer_resolve(int n_ranks, const int32_t* __restrict__ recv_ids,
                                                           int32_t* __restrict__ src_meta,
                                                           const int32_t* __restrict__ slot_map, int64_t map_stride,
                                                           int32_t* __restrict__ desc, int32_t* __restrict__ own_list) {
  const int64_t total = src_meta[n_ranks * 3];
  const int64_t span = (int64_t)gridDim.x * kThreads;
  for (int64_t k0 = blockIdx.x * (int64_t)kThreads; k0 < total; k0 += span) {
    const int64_t k = k0 + threadIdx.x;
    bool own = false;
    if (k < total) {
    int s = 0;
    while (s + 1 < n_ranks && k >= src_meta[(s + 1) * 3]) ++s;
    const int32_t x = recv_ids[k];
    int pos[ESR_MAX_PEERS];
#pragma unroll
    for (int q = 0; q < ESR_MAX_PEERS; ++q) pos[q] = q < n_ranks ? slot_map[q * map_stride + x] : -1;
    bool first = true;
#pragma unroll
    for (int q = 0; q < ESR_MAX_PEERS; ++q)
      if (q < s && pos[q] >= 0) first = false;
#pragma unroll
    for (int q = 0; q < ESR_MAX_PEERS; ++q)
      if (q < n_ranks) desc[k * n_ranks + q] = (first && q >= s && pos[q] >= 0) ? src_meta[q * 3 + 0] + pos[q] : -1;
    own = first;
    }
    // compact the entries that own their row (warp-aggregated append; the order of the list only decides which
    // group processes which row, never a summation order)
    const unsigned m = __ballot_sync(FULL, own);
    if (m) {
      const int lane = threadIdx.x & 31;
      int base = 0;
      if (lane == __ffs(m) - 1) base = atomicAdd(src_meta + n_ranks * 3 + 1, __popc(m));
      base = __shfl_sync(FULL, base, __ffs(m) - 1);
      if (own) own_list[base + __popc(m & ((1u << lane) - 1u))] = (int32_t)k;
    }
  }
}

// One group of TPR lanes per OWNER entry (compacted list), EB entries per iteration: sum the row's gradients
// over the sources in source order (from my inbox, where the sources' row passes scattered them), then
// optax.adagrad on the local shard row.  NR = number of sources rounded up to a power of two (compile time,
// so the per-source loops carry no dead code); every load of the EB entries is issued before its first use.
// (The first version looped over ESR_MAX_PEERS with predicates: 128 registers, 2 CTAs per SM, 2 TB/s.)
template <int TPR, int NR, int EB, int MINB>
__global__ void __launch_bounds__(kThreads, MINB) k_peer_merge_adagrad(const float4* __restrict__ inbox_dE,
                                                                       const float* __restrict__ inbox_db, int n_ranks,
                                                                       const int32_t* __restrict__ recv_ids,
                                                                       const int32_t* __restrict__ src_meta,
                                                                       const int32_t* __restrict__ desc,
                                                                       const int32_t* __restrict__ own_list, int D4,
                                                                       float* __restrict__ rows, float* __restrict__ acc,
                                                                       float* __restrict__ bias, float* __restrict__ bias_acc,
                                                                       float lr, float eps) {
  const int lane = threadIdx.x % TPR;
  const int64_t total = src_meta[n_ranks * 3 + 1];
  const int64_t groups = (int64_t)gridDim.x * (kThreads / TPR);
  for (int64_t i0 = (blockIdx.x * (int64_t)kThreads + threadIdx.x) / TPR; i0 < total; i0 += groups * EB) {
    int gi[EB][NR];
    int64_t x[EB];
    bool on[EB];
#pragma unroll
    for (int e = 0; e < EB; ++e) {
      const int64_t i = i0 + e * groups;
      on[e] = i < total;
      const int64_t k = on[e] ? own_list[i] : 0;
      x[e] = on[e] ? recv_ids[k] : 0;
#pragma unroll
      for (int q = 0; q < NR; ++q) gi[e][q] = (on[e] && q < n_ranks) ? desc[k * n_ranks + q] : -1;
    }
    // bias scalars (lane 0) -- loaded up front so their latency overlaps the row traffic
    float bp[EB], ba[EB], bg[EB];
#pragma unroll
    for (int e = 0; e < EB; ++e) {
      bp[e] = ba[e] = bg[e] = 0.f;
      if (on[e] && lane == 0) {
        bp[e] = bias[x[e]];
        ba[e] = bias_acc[x[e]];
#pragma unroll
        for (int q = 0; q < NR; ++q)
          if (gi[e][q] >= 0) bg[e] += inbox_db[gi[e][q]];
      }
    }
    for (int c = lane; c < D4; c += TPR) {
      float4 g[EB][NR], pv[EB], av[EB];
#pragma unroll
      for (int e = 0; e < EB; ++e) {
        if (on[e]) {
          pv[e] = reinterpret_cast<const float4*>(rows)[x[e] * D4 + c];
          av[e] = ld_stream(reinterpret_cast<const float4*>(acc) + x[e] * D4 + c);
        }
#pragma unroll
        for (int q = 0; q < NR; ++q) g[e][q] = gi[e][q] >= 0 ? ld_stream(inbox_dE + (int64_t)gi[e][q] * D4 + c) : f4_zero();
      }
#pragma unroll
      for (int e = 0; e < EB; ++e) {
        if (on[e]) {
          float4 s = g[e][0];
#pragma unroll
          for (int q = 1; q < NR; ++q) f4_add(s, g[e][q]);   // source order: a missing source adds +0.0
          adagrad4(pv[e], av[e], s, lr, eps);
          reinterpret_cast<float4*>(rows)[x[e] * D4 + c] = pv[e];
          st_stream(reinterpret_cast<float4*>(acc) + x[e] * D4 + c, av[e]);
        }
      }
    }
#pragma unroll
    for (int e = 0; e < EB; ++e) {
      if (on[e] && lane == 0) {
        adagrad1(bp[e], ba[e], bg[e], lr, eps);
        bias[x[e]] = bp[e];
        bias_acc[x[e]] = ba[e];
      }
    }
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

bool load_ptrs(PeerPtrs* out, const void* const* in, int n) {
  if (in == nullptr) return false;
  for (int i = 0; i < ESR_MAX_PEERS; ++i) out->p[i] = i < n ? in[i] : nullptr;
  for (int i = 0; i < n; ++i)
    if (in[i] == nullptr) return false;
  return true;
}

}  // namespace
}  // namespace esr

using namespace esr;

extern "C" int esr_peer_gather_f32(const void* const* peer_rows, const void* const* peer_bias, int32_t n_ranks,
                                   const int32_t* uniq, const int32_t* n_uniq, int64_t cap, int32_t D, float* out,
                                   float* out_bias, esr_stream_t stream_) {
  ESR_REQUIRE(n_ranks >= 1 && n_ranks <= ESR_MAX_PEERS && cap >= 0 && D > 0 && (D % 4) == 0);
  if (cap == 0) return ESR_OK;
  PeerPtrs pr, pb;
  ESR_REQUIRE(load_ptrs(&pr, peer_rows, n_ranks) && load_ptrs(&pb, peer_bias, n_ranks));
  ESR_REQUIRE(uniq && n_uniq && out && out_bias && (reinterpret_cast<uintptr_t>(out) % 16) == 0);
  const int D4 = D / 4;
  const int tpr = tpr_for(D4);
  constexpr int ROWS = 4;
  const int64_t want = ceil_div(ceil_div(cap, ROWS) * tpr, kThreads);
  const int64_t persistent = (int64_t)sm_count() * 8;
  const unsigned grid = (unsigned)(want < persistent ? want : persistent);
  Cyclic cyc;
  cyc.n = n_ranks;
  cyc.shift = -1;
  for (int b = 0; b < 4; ++b)
    if ((1 << b) == n_ranks) cyc.shift = b;
  ESR_DISPATCH_TPR(tpr, (k_peer_gather<TPR, ROWS><<<grid, kThreads, 0, static_cast<cudaStream_t>(stream_)>>>(
                            pr, pb, uniq, n_uniq, cap, cyc, D4, reinterpret_cast<float4*>(out), out_bias)));
  ESR_LAUNCH_CHECK();
  return ESR_OK;
}

extern "C" int esr_peer_pull_ids_i32(const void* const* peer_counts, const void* const* peer_send_local, int32_t n_ranks,
                                     int32_t me, int64_t recv_cap, int32_t* recv_ids, int32_t* src_meta, int32_t* slot_map,
                                     int64_t map_stride, esr_stream_t stream_) {
  ESR_REQUIRE(n_ranks >= 1 && n_ranks <= ESR_MAX_PEERS && me >= 0 && me < n_ranks && recv_cap > 0 && recv_ids && src_meta);
  ESR_REQUIRE(slot_map != nullptr && map_stride > 0);
  PeerPtrs pc, ps;
  ESR_REQUIRE(load_ptrs(&pc, peer_counts, n_ranks) && load_ptrs(&ps, peer_send_local, n_ranks));
  const int grid = 2 * sm_count();
  k_peer_pull_ids<<<grid, kThreads, 0, static_cast<cudaStream_t>(stream_)>>>(pc, ps, n_ranks, me, recv_cap, recv_ids,
                                                                             src_meta, slot_map, map_stride);
  ESR_LAUNCH_CHECK();
  return ESR_OK;
}

// Owner side, step 1 (ids only -- can run on a side stream while the row pass is still producing gradients):
// per received entry, where every source keeps that row's gradient (desc), and the compacted list of the entries
// that own their row (own_list = desc + recv_cap * n_ranks, counter in src_meta[3n + 1]).
extern "C" int esr_peer_resolve_i32(int32_t n_ranks, const int32_t* recv_ids, int32_t* src_meta, const int32_t* slot_map,
                                    int64_t map_stride, int32_t* desc, int64_t recv_cap, esr_stream_t stream_) {
  ESR_REQUIRE(n_ranks >= 1 && n_ranks <= ESR_MAX_PEERS && recv_ids && src_meta && slot_map && map_stride > 0 && desc &&
              recv_cap > 0);
  k_peer_resolve<<<4 * sm_count(), kThreads, 0, static_cast<cudaStream_t>(stream_)>>>(n_ranks, recv_ids, src_meta, slot_map,
                                                                                     map_stride, desc, desc + recv_cap * n_ranks);
  ESR_LAUNCH_CHECK();
  return ESR_OK;
}

// Owner side, step 2 (after every rank's gradients have landed): merge in source order + optax.adagrad, then restore
// slot_map to -1.
extern "C" int esr_peer_apply_adagrad_f32(EsrTable* shard, const float* inbox_dE, const float* inbox_db, int32_t n_ranks,
                                          const int32_t* recv_ids, const int32_t* src_meta, int32_t* slot_map,
                                          int64_t map_stride, const int32_t* desc, int64_t recv_cap, float lr, float eps,
                                          esr_stream_t stream_) {
  ESR_REQUIRE(shard && shard->struct_size >= sizeof(EsrTable) && shard->D > 0 && (shard->D % 4) == 0 && desc);
  ESR_REQUIRE(shard->rows[0] && shard->acc && shard->bias && shard->bias_acc && shard->ver == nullptr);
  ESR_REQUIRE(n_ranks >= 1 && n_ranks <= ESR_MAX_PEERS && recv_ids && src_meta && slot_map && map_stride > 0 && recv_cap > 0);
  ESR_REQUIRE(inbox_dE && inbox_db && (reinterpret_cast<uintptr_t>(inbox_dE) % 16) == 0);
  cudaStream_t stream = static_cast<cudaStream_t>(stream_);
  const int D4 = shard->D / 4;
  const int32_t* own_list = desc + recv_cap * n_ranks;
  const int tpr = tpr_for(D4);
  const int grid = 8 * sm_count();
#define ESR_MERGE(NR, EB, MINB)                                                                                          \
  ESR_DISPATCH_TPR(tpr, (k_peer_merge_adagrad<TPR, NR, EB, MINB><<<grid, kThreads, 0, stream>>>(                          \
                            reinterpret_cast<const float4*>(inbox_dE), inbox_db, n_ranks, recv_ids, src_meta, desc, own_list, \
                            D4, shard->rows[0], shard->acc, shard->bias, shard->bias_acc, lr, eps)))
  if (n_ranks == 1) {
    ESR_MERGE(1, 4, 3);
  } else if (n_ranks == 2) {
    ESR_MERGE(2, 4, 2);
  } else if (n_ranks <= 4) {
    ESR_MERGE(4, 2, 3);
  } else {
    ESR_MERGE(8, 2, 2);
  }
#undef ESR_MERGE
  ESR_LAUNCH_CHECK();
  k_peer_clear_map<<<2 * sm_count(), kThreads, 0, stream>>>(recv_ids, src_meta, n_ranks, slot_map, map_stride);
  ESR_LAUNCH_CHECK();
  return ESR_OK;
}

extern "C" int esr_peer_merge_adagrad_f32(EsrTable* shard, const float* inbox_dE, const float* inbox_db, int32_t n_ranks,
                                          const int32_t* recv_ids, const int32_t* src_meta, int32_t* slot_map,
                                          int64_t map_stride, int32_t* desc, int64_t recv_cap, float lr, float eps,
                                          esr_stream_t stream_) {
  const int rc = esr_peer_resolve_i32(n_ranks, recv_ids, const_cast<int32_t*>(src_meta), slot_map, map_stride, desc, recv_cap,
                                      stream_);
  if (rc != ESR_OK) return rc;
  return esr_peer_apply_adagrad_f32(shard, inbox_dE, inbox_db, n_ranks, recv_ids, src_meta, slot_map, map_stride, desc,
                                    recv_cap, lr, eps, stream_);
}

// ---------------------------------------------------------------------------------------------
// esr_peer_allreduce_f32: all-reduce of a handful of floats (or, with count == 0, a pure barrier) over symmetric peer
// memory -- the step's two batch-sum reductions and its phase barriers as ONE plain kernel each, so the sharded step
// holds no NCCL call and stays CUDA-graph capturable, and a synchronisation costs one NVLink round trip.
// Every rank owns a sync block of kSyncRing x ESR_MAX_PEERS slots of 8 words: [0..5] payload, [7] sequence flag.
// Call k (k = 1, 2, ...; the counter lives on the device so a replayed graph advances it):
//   publish : lane r < n stores my payload, then (release, system scope) the flag k into slot [k % ring][me] of rank r;
//   wait    : lane s < n polls slot [k % ring][s] of MY block until its flag is k (acquire, system scope);
//   reduce  : lane c < count adds the n payloads in rank order -> bit-identical sums on every rank.
// A rank can run at most one call ahead of the slowest one (it cannot leave call k before everyone published k), so a
// ring of 2 would do; 4 is used.  The poll is bounded by wall time: a lost peer traps instead of hanging the GPU.
// EXPERIMENTAL in round 1 (written after the GPU budget was spent): PeerShardedGloveTrainer(fast_sync=True).
// ---------------------------------------------------------------------------------------------
constexpr int kSyncRing = 4;
constexpr int kSyncWords = 8;
constexpr int kSyncMaxCount = 6;
constexpr unsigned long long kSyncTimeoutNs = 20ull * 1000 * 1000 * 1000;

__device__ __forceinline__ void st_release_sys(uint32_t* p, uint32_t v) {
  asm volatile("st.release.sys.global.u32 [%0], %1;" ::"l"(p), "r"(v) : "memory");
}
__device__ __forceinline__ uint32_t ld_acquire_sys(const uint32_t* p) {
  uint32_t v;
  asm volatile("ld.acquire.sys.global.u32 %0, [%1];" : "=r"(v) : "l"(p) : "memory");
  return v;
}
__device__ __forceinline__ void st_relaxed_sys(uint32_t* p, uint32_t v) {
  asm volatile("st.relaxed.sys.global.u32 [%0], %1;" ::"l"(p), "r"(v) : "memory");
}
__device__ __forceinline__ uint32_t ld_relaxed_sys(const uint32_t* p) {
  uint32_t v;
  asm volatile("ld.relaxed.sys.global.u32 %0, [%1];" : "=r"(v) : "l"(p) : "memory");
  return v;
}
__device__ __forceinline__ unsigned long long global_ns() {
  unsigned long long t;
  asm volatile("mov.u64 %0, %%globaltimer;" : "=l"(t));
  return t;
}

__global__ void __launch_bounds__(32) k_peer_allreduce(const __grid_constant__ PeerPtrs sync, int n, int me, const float* in, float* out, int count,
                                                       uint32_t* seq_counter) {
  const int lane = threadIdx.x;
  const uint32_t seq = *seq_counter + 1u;
  const int ring = (int)(seq % kSyncRing);
  float mine[kSyncMaxCount];
#pragma unroll
  for (int c = 0; c < kSyncMaxCount; ++c) mine[c] = c < count ? in[c] : 0.f;
  if (lane < n) {
    uint32_t* dst = reinterpret_cast<uint32_t*>(const_cast<void*>(sync.p[lane])) + ((size_t)ring * ESR_MAX_PEERS + me) * kSyncWords;
#pragma unroll
    for (int c = 0; c < kSyncMaxCount; ++c)
      if (c < count) st_relaxed_sys(dst + c, __float_as_uint(mine[c]));
    st_release_sys(dst + 7, seq);  // release: the payload stores above are visible before the flag
  }
  const uint32_t* my = reinterpret_cast<const uint32_t*>(sync.p[me]) + (size_t)ring * ESR_MAX_PEERS * kSyncWords;
  if (lane < n) {
    const uint32_t* flag = my + (size_t)lane * kSyncWords + 7;
    const unsigned long long t0 = global_ns();
    while (ld_acquire_sys(flag) != seq) {
      if (global_ns() - t0 > kSyncTimeoutNs) __trap();
      __nanosleep(40);
    }
  }
  __syncwarp();
  __threadfence_system();
  if (lane < count) {
    float sum = 0.f;
    for (int r = 0; r < n; ++r) sum += __uint_as_float(ld_relaxed_sys(my + (size_t)r * kSyncWords + lane));
    out[lane] = sum;
  }
  __syncwarp();
  if (lane == 0) *seq_counter = seq;
}

extern "C" size_t esr_peer_sync_bytes(void) { return (size_t)kSyncRing * ESR_MAX_PEERS * kSyncWords * sizeof(uint32_t); }

extern "C" int esr_peer_allreduce_f32(void* const* peer_sync, int32_t n_ranks, int32_t me, const float* in, float* out,
                                      int32_t count, uint32_t* seq_counter, esr_stream_t stream_) {
  ESR_REQUIRE(n_ranks >= 1 && n_ranks <= ESR_MAX_PEERS && me >= 0 && me < n_ranks && count >= 0 && count <= kSyncMaxCount);
  ESR_REQUIRE(seq_counter != nullptr && (count == 0 || (in != nullptr && out != nullptr)));
  PeerPtrs ps;
  ESR_REQUIRE(load_ptrs(&ps, reinterpret_cast<const void* const*>(peer_sync), n_ranks));
  k_peer_allreduce<<<1, 32, 0, static_cast<cudaStream_t>(stream_)>>>(ps, n_ranks, me, in, out, count, seq_counter);
  ESR_LAUNCH_CHECK();
  return ESR_OK;
}

// emit_map[u] = owner << 27 | (offset of my bucket in owner's inbox + position inside the bucket)
__global__ void __launch_bounds__(kThreads) k_peer_emit_plan(const __grid_constant__ PeerPtrs counts, int n_ranks, int me, Cyclic cyc,
                                                             const int32_t* __restrict__ uniq,
                                                             const int32_t* __restrict__ n_uniq, int64_t cap,
                                                             const int32_t* __restrict__ inv_order, int64_t inbox_cap,
                                                             int32_t* __restrict__ emit_map, int32_t* __restrict__ err) {
  __shared__ int cm[ESR_MAX_PEERS][ESR_MAX_PEERS];
  __shared__ int off[ESR_MAX_PEERS], dsp[ESR_MAX_PEERS];
  if ((int)threadIdx.x < n_ranks * n_ranks) {
    const int s = threadIdx.x / n_ranks, q = threadIdx.x % n_ranks;
    cm[s][q] = reinterpret_cast<const int32_t*>(counts.p[s])[q];
  }
  __syncthreads();
  if ((int)threadIdx.x < n_ranks) {
    const int o = threadIdx.x;
    int a = 0, d = 0;
    for (int s = 0; s < me; ++s) a += cm[s][o];  // sources before me in owner o's inbox
    for (int q = 0; q < o; ++q) d += cm[me][q];  // my buckets before owner o
    off[o] = a;
    dsp[o] = d;
  }
  __syncthreads();
  const int64_t n = min((int64_t)*n_uniq, cap);
  for (int64_t u = blockIdx.x * (int64_t)kThreads + threadIdx.x; u < n; u += (int64_t)gridDim.x * kThreads) {
    const int o = cyc.owner(uniq[u]);
    const int64_t idx = (int64_t)off[o] + inv_order[u] - dsp[o];
    if (idx >= inbox_cap || idx >= (1 << 27)) {
      *err = 1;
      emit_map[u] = o << 27;  // clamp: keep the store in bounds
    } else {
      emit_map[u] = (o << 27) | (int32_t)idx;
    }
  }
}

extern "C" int esr_peer_emit_plan_i32(const void* const* peer_counts, int32_t n_ranks, int32_t me, const int32_t* uniq,
                                      const int32_t* n_uniq, int64_t cap, const int32_t* inv_order, int64_t inbox_cap,
                                      int32_t* emit_map, int32_t* err, esr_stream_t stream_) {
  ESR_REQUIRE(n_ranks >= 1 && n_ranks <= ESR_MAX_PEERS && me >= 0 && me < n_ranks && cap >= 0 && inbox_cap > 0);
  if (cap == 0) return ESR_OK;
  ESR_REQUIRE(uniq && n_uniq && inv_order && emit_map && err);
  PeerPtrs pc;
  ESR_REQUIRE(load_ptrs(&pc, peer_counts, n_ranks));
  Cyclic cyc;
  cyc.n = n_ranks;
  cyc.shift = -1;
  for (int b = 0; b < 4; ++b)
    if ((1 << b) == n_ranks) cyc.shift = b;
  k_peer_emit_plan<<<2 * sm_count(), kThreads, 0, static_cast<cudaStream_t>(stream_)>>>(
      pc, n_ranks, me, cyc, uniq, n_uniq, cap, inv_order, inbox_cap, emit_map, err);
  ESR_LAUNCH_CHECK();
  return ESR_OK;
}
