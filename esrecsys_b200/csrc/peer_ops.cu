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
    float bv[ROWS];
#pragma unroll
    for (int r = 0; r < ROWS; ++r) {
      const int64_t u = first + r;
      src[r] = nullptr;
      bv[r] = 0.f;
      if (u < n) {
        const int32_t row = uniq[u];
        const int owner = cyc.owner(row);
        const int64_t local = cyc.local(row);
        src[r] = reinterpret_cast<const float4*>(rows.p[owner]) + local * D4;
        // the bias is a 4-byte PEER load too: keep it in a register and store it after the row loads are in flight -- storing
        // it here put one full NVLink round trip in front of every batch of row loads (measured: 430 vs 665 GB/s)
        if (lane == 0) bv[r] = reinterpret_cast<const float*>(bias.p[owner])[local];
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
    if (lane == 0) {
#pragma unroll
      for (int r = 0; r < ROWS; ++r)
        if (src[r]) out_bias[first + r] = bv[r];
    }
  }
}

// Remote rows only (owner-routed step: the rows this rank owns are read where they live).  Walks the route plan's bucket
// order -- unique rows grouped by owner -- and skips my own bucket as one contiguous range, so every group has work.
template <int TPR, int ROWS>
__global__ void __launch_bounds__(kThreads) k_peer_gather_remote(const __grid_constant__ PeerPtrs rows, const __grid_constant__ PeerPtrs bias,
                                                                 const int32_t* __restrict__ uniq, const int32_t* __restrict__ order,
                                                                 const int32_t* __restrict__ counts, int me, Cyclic cyc, int D4,
                                                                 float4* __restrict__ out, float* __restrict__ out_bias, int parts) {
  const int lane = threadIdx.x % TPR;
  int before = 0, mine = 0, total = 0;
  for (int o = 0; o < cyc.n; ++o) {
    const int c = counts[o];
    if (o < me) before += c;
    if (o == me) mine = c;
    total += c;
  }
  const int64_t n = total - mine;
  const int64_t groups = (int64_t)gridDim.x * (kThreads / TPR);
  for (int64_t first = ((blockIdx.x * (int64_t)kThreads + threadIdx.x) / TPR) * ROWS; first < n; first += groups * ROWS) {
    const float4* src[ROWS];
    int32_t u[ROWS];
    float bv[ROWS];
#pragma unroll
    for (int r = 0; r < ROWS; ++r) {
      const int64_t k = first + r;
      src[r] = nullptr;
      bv[r] = 0.f;
      u[r] = 0;
      if (k < n) {
        u[r] = order[k < before ? k : k + mine];
        const int32_t row = uniq[u[r]];
        const int owner = cyc.owner(row);
        const int64_t local = cyc.local(row);
        src[r] = reinterpret_cast<const float4*>(rows.p[owner]) + local * D4;
        if (lane == 0 && (parts & 2)) bv[r] = reinterpret_cast<const float*>(bias.p[owner])[local];
      }
    }
    if (parts & 1) {
      for (int c = lane; c < D4; c += TPR) {
        float4 v[ROWS];
#pragma unroll
        for (int r = 0; r < ROWS; ++r)
          if (src[r]) v[r] = src[r][c];
#pragma unroll
        for (int r = 0; r < ROWS; ++r)
          if (src[r]) out[(int64_t)u[r] * D4 + c] = v[r];  // default caching: the row pass re-reads these from L2
      }
    }
    if (lane == 0 && (parts & 2)) {
#pragma unroll
      for (int r = 0; r < ROWS; ++r)
        if (src[r]) out_bias[u[r]] = bv[r];
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
    bool own = false, multi = false;
    int32_t xk = 0;
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
    xk = x;
#pragma unroll
    for (int q = 0; q < ESR_MAX_PEERS; ++q)
      if (q > s && q < n_ranks && pos[q] >= 0) multi = true;
    }
    // compact the entries that own their row (warp-aggregated append; the order of the list only decides which
    // group processes which row, never a summation order).  Record = {entry k | several sources << 31, owner-local row}:
    // one coalesced 8-byte load per row in the merge instead of three dependent gathers.
    const unsigned m = __ballot_sync(FULL, own);
    if (m) {
      const int lane = threadIdx.x & 31;
      int base = 0;
      if (lane == __ffs(m) - 1) base = atomicAdd(src_meta + n_ranks * 3 + 1, __popc(m));
      base = __shfl_sync(FULL, base, __ffs(m) - 1);
      if (own)
        reinterpret_cast<int2*>(own_list)[base + __popc(m & ((1u << lane) - 1u))] =
            make_int2((int32_t)((uint32_t)k | (multi ? 0x80000000u : 0u)), xk);
    }
  }
}

// One group of TPR lanes per OWNER record, EB records per iteration: the row's gradient from the first source naming it is
// inbox row k itself (recv order == inbox order); shard row, accumulator row and that gradient row of all EB records are
// issued before the first use, and the records of the NEXT iteration are fetched while this one is in flight.  Further
// sources -- rare: only rows several ranks touched in the same step -- are added in source order, four independent
// loads at a time.  Sum order: source order (a missing source adds +0.0) -> deterministic.
template <int TPR, int EB>
__global__ void __launch_bounds__(kThreads, 2) k_peer_merge_adagrad(const float4* __restrict__ inbox_dE, int n_ranks,
                                                                    const int32_t* __restrict__ src_meta,
                                                                    const int32_t* __restrict__ desc,
                                                                    const int2* __restrict__ own_rec, int D4,
                                                                    float* __restrict__ rows, float* __restrict__ acc,
                                                                    float lr, float eps) {
  const int lane = threadIdx.x % TPR;
  const int64_t total = src_meta[n_ranks * 3 + 1];
  const int64_t groups = (int64_t)gridDim.x * (kThreads / TPR);
  const int64_t g0 = (blockIdx.x * (int64_t)kThreads + threadIdx.x) / TPR;
  int2 rec[EB], nxt[EB];
#pragma unroll
  for (int e = 0; e < EB; ++e) {
    const int64_t i = g0 + e * groups;
    rec[e] = i < total ? own_rec[i] : make_int2(-1, 0);
  }
  for (int64_t i0 = g0; i0 < total; i0 += groups * EB) {
    for (int c = lane; c < D4; c += TPR) {
      float4 g[EB], pv[EB], av[EB];
#pragma unroll
      for (int e = 0; e < EB; ++e) {
        const bool on = i0 + e * groups < total;
        if (on) {
          const int64_t k = rec[e].x & 0x7fffffff, x = rec[e].y;
          pv[e] = reinterpret_cast<const float4*>(rows)[x * D4 + c];
          av[e] = ld_stream(reinterpret_cast<const float4*>(acc) + x * D4 + c);
          g[e] = ld_stream(inbox_dE + k * D4 + c);
        }
      }
      if (c == lane) {  // next iteration's records: in flight behind the row loads
#pragma unroll
        for (int e = 0; e < EB; ++e) {
          const int64_t i = i0 + (e + EB) * groups;
          nxt[e] = i < total ? own_rec[i] : make_int2(-1, 0);
        }
      }
#pragma unroll
      for (int e = 0; e < EB; ++e) {
        if (i0 + e * groups < total) {
          const int64_t k = rec[e].x & 0x7fffffff, x = rec[e].y;
          if (rec[e].x < 0) {  // several sources: add the others in source order
            int q0 = 0;
            while (q0 < n_ranks && desc[k * n_ranks + q0] < 0) ++q0;
            for (int qb = q0 + 1; qb < n_ranks; qb += 4) {
              int d[4];
              float4 t[4];
#pragma unroll
              for (int j = 0; j < 4; ++j) d[j] = qb + j < n_ranks ? desc[k * n_ranks + qb + j] : -1;
#pragma unroll
              for (int j = 0; j < 4; ++j) t[j] = d[j] >= 0 ? ld_stream(inbox_dE + (int64_t)d[j] * D4 + c) : f4_zero();
#pragma unroll
              for (int j = 0; j < 4; ++j) f4_add(g[e], t[j]);
            }
          }
          adagrad4(pv[e], av[e], g[e], lr, eps);
          reinterpret_cast<float4*>(rows)[x * D4 + c] = pv[e];
          st_stream(reinterpret_cast<float4*>(acc) + x * D4 + c, av[e]);
        }
      }
    }
#pragma unroll
    for (int e = 0; e < EB; ++e) rec[e] = nxt[e];
  }
}

// Bias of the owned rows: one thread per owner record (the per-row scalars would otherwise put dependent 4-byte loads of
// one lane in front of every row of the kernel above).
__global__ void __launch_bounds__(kThreads) k_peer_merge_bias(const float* __restrict__ inbox_db, int n_ranks,
                                                              const int32_t* __restrict__ src_meta,
                                                              const int32_t* __restrict__ desc,
                                                              const int2* __restrict__ own_rec, float* __restrict__ bias,
                                                              float* __restrict__ bias_acc, float lr, float eps) {
  const int64_t total = src_meta[n_ranks * 3 + 1];
  for (int64_t i = blockIdx.x * (int64_t)kThreads + threadIdx.x; i < total; i += (int64_t)gridDim.x * kThreads) {
    const int2 r = own_rec[i];
    const int64_t k = r.x & 0x7fffffff, x = r.y;
    float g = inbox_db[k];
    if (r.x < 0) {
      int q0 = 0;
      while (q0 < n_ranks && desc[k * n_ranks + q0] < 0) ++q0;
      for (int q = q0 + 1; q < n_ranks; ++q) {
        const int d = desc[k * n_ranks + q];
        if (d >= 0) g += inbox_db[d];
      }
    }
    float p = bias[x], a = bias_acc[x];
    adagrad1(p, a, g, lr, eps);
    bias[x] = p;
    bias_acc[x] = a;
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
  ESR_RANGE("esr_peer_gather_f32");
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

extern "C" int esr_peer_gather_remote_f32(const void* const* peer_rows, const void* const* peer_bias, int32_t n_ranks, int32_t me,
                                          const int32_t* uniq, const int32_t* order, const int32_t* counts, int64_t cap,
                                          int32_t D, float* out, float* out_bias, int32_t parts, esr_stream_t stream_) {
  ESR_RANGE("esr_peer_gather_remote_f32");
  ESR_REQUIRE(n_ranks >= 1 && n_ranks <= ESR_MAX_PEERS && me >= 0 && me < n_ranks && cap >= 0 && D > 0 && (D % 4) == 0);
  ESR_REQUIRE((parts & 3) != 0);
  if (cap == 0 || n_ranks == 1) return ESR_OK;  // one rank: nothing is remote
  PeerPtrs pr, pb;
  ESR_REQUIRE(load_ptrs(&pr, peer_rows, n_ranks) && load_ptrs(&pb, peer_bias, n_ranks));
  ESR_REQUIRE(uniq && order && counts && out && out_bias && (reinterpret_cast<uintptr_t>(out) % 16) == 0);
  const int D4 = D / 4;
  // biases only: one lane per row does the work, so use the narrowest group
  const int tpr = (parts & 1) ? tpr_for(D4) : 1;
  constexpr int ROWS = 4;
  const int64_t want = ceil_div(ceil_div(cap, ROWS) * tpr, kThreads);
  const int64_t persistent = (int64_t)sm_count() * ((parts & 1) ? 8 : 2);
  const unsigned grid = (unsigned)(want < persistent ? want : persistent);
  Cyclic cyc;
  cyc.n = n_ranks;
  cyc.shift = -1;
  for (int b = 0; b < 4; ++b)
    if ((1 << b) == n_ranks) cyc.shift = b;
  ESR_DISPATCH_TPR(tpr, (k_peer_gather_remote<TPR, ROWS><<<grid, kThreads, 0, static_cast<cudaStream_t>(stream_)>>>(
                            pr, pb, uniq, order, counts, me, cyc, D4, reinterpret_cast<float4*>(out), out_bias, parts)));
  ESR_LAUNCH_CHECK();
  return ESR_OK;
}

extern "C" int esr_peer_pull_ids_i32(const void* const* peer_counts, const void* const* peer_send_local, int32_t n_ranks,
                                     int32_t me, int64_t recv_cap, int32_t* recv_ids, int32_t* src_meta, int32_t* slot_map,
                                     int64_t map_stride, esr_stream_t stream_) {
  ESR_RANGE("esr_peer_pull_ids_i32");
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
// that own their row, as 8-byte records {entry k | several-sources flag << 31, owner-local row} at desc + recv_cap * n_ranks
// (so desc holds recv_cap * (n_ranks + 2) ints; counter in src_meta[3n + 1]).
extern "C" int esr_peer_resolve_i32(int32_t n_ranks, const int32_t* recv_ids, int32_t* src_meta, const int32_t* slot_map,
                                    int64_t map_stride, int32_t* desc, int64_t recv_cap, esr_stream_t stream_) {
  ESR_RANGE("esr_peer_resolve_i32");
  ESR_REQUIRE(n_ranks >= 1 && n_ranks <= ESR_MAX_PEERS && recv_ids && src_meta && slot_map && map_stride > 0 && desc &&
              recv_cap > 0 && ((recv_cap * n_ranks) % 2) == 0 && (reinterpret_cast<uintptr_t>(desc) % 8) == 0);
  k_peer_resolve<<<4 * sm_count(), kThreads, 0, static_cast<cudaStream_t>(stream_)>>>(n_ranks, recv_ids, src_meta, slot_map,
                                                                                     map_stride, desc, desc + recv_cap * n_ranks);
  ESR_LAUNCH_CHECK();
  return ESR_OK;
}

// Owner side, step 2 (after every rank's gradients have landed): merge in source order + optax.adagrad, then restore
// slot_map to -1.
// parts: 1 = embedding rows, 2 = biases + restore slot_map.  The two halves touch disjoint state, so a caller may run them on
// different streams: the rows as soon as every rank's gradient ROWS have landed, the biases after the bias gradients have.
extern "C" int esr_peer_apply_parts_f32(EsrTable* shard, const float* inbox_dE, const float* inbox_db, int32_t n_ranks,
                                        const int32_t* recv_ids, const int32_t* src_meta, int32_t* slot_map,
                                        int64_t map_stride, const int32_t* desc, int64_t recv_cap, float lr, float eps,
                                        int32_t parts, esr_stream_t stream_) {
  ESR_RANGE("esr_peer_apply_parts_f32");
  ESR_REQUIRE(shard && shard->struct_size >= sizeof(EsrTable) && shard->D > 0 && (shard->D % 4) == 0 && desc);
  ESR_REQUIRE(shard->rows[0] && shard->acc && shard->bias && shard->bias_acc && shard->ver == nullptr);
  ESR_REQUIRE(n_ranks >= 1 && n_ranks <= ESR_MAX_PEERS && recv_ids && src_meta && slot_map && map_stride > 0 && recv_cap > 0);
  ESR_REQUIRE(inbox_dE && inbox_db && (reinterpret_cast<uintptr_t>(inbox_dE) % 16) == 0 && (parts & 3) != 0);
  cudaStream_t stream = static_cast<cudaStream_t>(stream_);
  const int D4 = shard->D / 4;
  const int2* own_rec = reinterpret_cast<const int2*>(desc + recv_cap * n_ranks);  // 8-byte aligned (resolve checks it)
  if (parts & 1) {
    const int tpr = tpr_for(D4);
    const int grid = 8 * sm_count();
    ESR_DISPATCH_TPR(tpr, (k_peer_merge_adagrad<TPR, 4><<<grid, kThreads, 0, stream>>>(
                              reinterpret_cast<const float4*>(inbox_dE), n_ranks, src_meta, desc, own_rec, D4, shard->rows[0],
                              shard->acc, lr, eps)));
    ESR_LAUNCH_CHECK();
  }
  if (parts & 2) {
    k_peer_merge_bias<<<2 * sm_count(), kThreads, 0, stream>>>(inbox_db, n_ranks, src_meta, desc, own_rec, shard->bias,
                                                               shard->bias_acc, lr, eps);
    ESR_LAUNCH_CHECK();
    k_peer_clear_map<<<2 * sm_count(), kThreads, 0, stream>>>(recv_ids, src_meta, n_ranks, slot_map, map_stride);
    ESR_LAUNCH_CHECK();
  }
  return ESR_OK;
}

extern "C" int esr_peer_apply_adagrad_f32(EsrTable* shard, const float* inbox_dE, const float* inbox_db, int32_t n_ranks,
                                          const int32_t* recv_ids, const int32_t* src_meta, int32_t* slot_map,
                                          int64_t map_stride, const int32_t* desc, int64_t recv_cap, float lr, float eps,
                                          esr_stream_t stream_) {
  return esr_peer_apply_parts_f32(shard, inbox_dE, inbox_db, n_ranks, recv_ids, src_meta, slot_map, map_stride, desc, recv_cap,
                                  lr, eps, 3, stream_);
}

extern "C" int esr_peer_merge_adagrad_f32(EsrTable* shard, const float* inbox_dE, const float* inbox_db, int32_t n_ranks,
                                          const int32_t* recv_ids, const int32_t* src_meta, int32_t* slot_map,
                                          int64_t map_stride, int32_t* desc, int64_t recv_cap, float lr, float eps,
                                          esr_stream_t stream_) {
  ESR_RANGE("esr_peer_merge_adagrad_f32");
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
  ESR_RANGE("esr_peer_allreduce_f32");
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
      atomicOr(err, 1);
      emit_map[u] = o << 27;  // clamp: keep the store in bounds
    } else {
      emit_map[u] = (o << 27) | (int32_t)idx;
    }
  }
}

extern "C" int esr_peer_emit_plan_i32(const void* const* peer_counts, int32_t n_ranks, int32_t me, const int32_t* uniq,
                                      const int32_t* n_uniq, int64_t cap, const int32_t* inv_order, int64_t inbox_cap,
                                      int32_t* emit_map, int32_t* err, esr_stream_t stream_) {
  ESR_RANGE("esr_peer_emit_plan_i32");
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

// ---------------------------------------------------------------------------------------------
// OWNER-COMPUTES pair routing.  A pair (i, j, x) is processed by the rank that OWNS row i (owner = i % n): that
// rank reads E[i] from its own shard, so only the unique PARTNER rows j of its pairs cross NVLink (ids are frequency
// ranks and i > j -- wikipedia/make_dictionary.py:113-116, make_cooccurrence.py:48 -- so the j side is the hot,
// heavily duplicated side), and only their gradients travel back.  Counted on the bench stream (Zipf(1), V = 1M,
// B = 262 144 per GPU, 8 ranks): 28 062 remote unique rows per rank and step instead of 119 861 when every rank
// keeps the pairs it was handed (4.3x fewer NVLink bytes in each direction); V = 100M: 56 770 instead of 208 461.
// The triples themselves are 12 bytes per pair: routing them costs 2 % of what it saves.
//
//  esr_peer_route_pairs_i32   source side (ids only: side stream): STABLE partition of my B pairs by owner(i) written
//                             straight into the owners' pair inboxes (region of source `me`, one 16-byte record
//                             {i, j, count, 0} = one NVLink store per pair), plus my per-owner counts.  Stable => the received batch is a deterministic function of the global batch.
//  esr_peer_collect_pairs_i32 owner side, after a barrier: concatenates the regions in source order into the flat
//                             [i ; j] slot array of the plan (capacity B_cap per half, padding key = pad_key sorts to
//                             the end), the counts, and *n_valid = 2 m.  err |= 2 if m > B_cap (pairs dropped).
// Bit-exact contract: oracle/index.py route_pairs / collect_pairs.
// ---------------------------------------------------------------------------------------------
namespace esr {
namespace {

constexpr int kRouteItems = 8;
constexpr int kRouteTile = kThreads * kRouteItems;  // 2048 pairs per block, blocked arrangement (order preserving)

struct OwnerCounts {  // per-owner counters packed 16 bits each (a tile holds 2048 pairs < 65536)
  unsigned long long lo, hi;  // owners 0..3, 4..7
  __device__ __forceinline__ void add(int o) {
    if (o < 4) lo += 1ull << (16 * o);
    else hi += 1ull << (16 * (o - 4));
  }
  __device__ __forceinline__ int get(int o) const { return (int)(((o < 4 ? lo : hi) >> (16 * (o & 3))) & 0xffffull); }
};

__device__ __forceinline__ unsigned long long warp_excl_scan_u64(unsigned long long v, unsigned long long* total) {
  const int lane = threadIdx.x & 31;
  unsigned long long x = v;
#pragma unroll
  for (int o = 1; o < 32; o <<= 1) {
    const unsigned long long y = __shfl_up_sync(FULL, x, o);
    if (lane >= o) x += y;
  }
  *total = __shfl_sync(FULL, x, 31);
  return x - v;
}

// per block and owner: number of pairs of the tile whose row i lives on that owner
__global__ void __launch_bounds__(kThreads) k_pair_count(const int32_t* __restrict__ ids_i, int64_t B, Cyclic cyc,
                                                         int32_t* __restrict__ blk_cnt /* [blocks][8] */) {
  __shared__ int cnt[ESR_MAX_PEERS];
  if (threadIdx.x < ESR_MAX_PEERS) cnt[threadIdx.x] = 0;
  __syncthreads();
  const int64_t base = (int64_t)blockIdx.x * kRouteTile + (int64_t)threadIdx.x * kRouteItems;
  OwnerCounts c{0ull, 0ull};
#pragma unroll
  for (int k = 0; k < kRouteItems; ++k)
    if (base + k < B) c.add(cyc.owner(ids_i[base + k]));
  for (int o = 0; o < cyc.n; ++o) {
    int v = c.get(o);
    for (int s = 16; s > 0; s >>= 1) v += __shfl_xor_sync(FULL, v, s);
    if ((threadIdx.x & 31) == 0 && v) atomicAdd(&cnt[o], v);  // integer counts: order-independent
  }
  __syncthreads();
  if (threadIdx.x < ESR_MAX_PEERS) blk_cnt[blockIdx.x * ESR_MAX_PEERS + threadIdx.x] = cnt[threadIdx.x];
}

// one block: exclusive prefix over the blocks for every owner, totals to the owners' count tables
__global__ void __launch_bounds__(ESR_MAX_PEERS * 32) k_pair_scan(const int32_t* __restrict__ blk_cnt, int n_blocks, int n_ranks,
                                                                  int me, int32_t* __restrict__ blk_base,
                                                                  const __grid_constant__ PeerPtrs peer_counts,
                                                                  int32_t* __restrict__ my_counts) {
  const int o = threadIdx.x >> 5, lane = threadIdx.x & 31;  // warp o scans owner o
  if (o >= n_ranks) return;
  int carry = 0;
  for (int b0 = 0; b0 < n_blocks; b0 += 32) {
    const int b = b0 + lane;
    const int v = b < n_blocks ? blk_cnt[b * ESR_MAX_PEERS + o] : 0;
    int x = v;
#pragma unroll
    for (int s = 1; s < 32; s <<= 1) {
      const int y = __shfl_up_sync(FULL, x, s);
      if (lane >= s) x += y;
    }
    if (b < n_blocks) blk_base[b * ESR_MAX_PEERS + o] = carry + x - v;
    carry += __shfl_sync(FULL, x, 31);
  }
  if (lane == 0) {
    my_counts[o] = carry;
    reinterpret_cast<int32_t*>(const_cast<void*>(peer_counts.p[o]))[me] = carry;  // owner o: "source me sends you carry pairs"
  }
}

// stable scatter of the tile's pairs into the owners' inbox regions of source `me`.  The tile is first ordered by owner in
// shared memory (stable), then every owner's run is copied to its region with consecutive lanes writing consecutive
// 16-byte records: 512-byte NVLink stores per warp instruction instead of 32 scattered ones (63 -> ~30 us at 8 ranks).
__global__ void __launch_bounds__(kThreads) k_pair_scatter(const int32_t* __restrict__ ids, const float* __restrict__ counts,
                                                           int64_t B, Cyclic cyc, int me, const int32_t* __restrict__ blk_base,
                                                           const __grid_constant__ PeerPtrs peer_ids) {
  __shared__ unsigned long long wlo[kThreads / 32], whi[kThreads / 32];
  __shared__ int4 tile[kRouteTile];
  __shared__ int tile_off[ESR_MAX_PEERS + 1];
  const int64_t base = (int64_t)blockIdx.x * kRouteTile + (int64_t)threadIdx.x * kRouteItems;
  int32_t vi[kRouteItems], vj[kRouteItems];
  float vx[kRouteItems];
  int own[kRouteItems];
  OwnerCounts c{0ull, 0ull};
#pragma unroll
  for (int k = 0; k < kRouteItems; ++k) {
    own[k] = -1;
    if (base + k < B) {
      vi[k] = ids[base + k];
      vj[k] = ids[B + base + k];
      vx[k] = counts[base + k];
      own[k] = cyc.owner(vi[k]);
      c.add(own[k]);
    }
  }
  // exclusive prefix of the per-thread counters over the block, thread order (= pair order)
  unsigned long long tlo, thi;
  OwnerCounts ex;
  ex.lo = warp_excl_scan_u64(c.lo, &tlo);
  ex.hi = warp_excl_scan_u64(c.hi, &thi);
  const int wid = threadIdx.x >> 5;
  if ((threadIdx.x & 31) == 31) {
    wlo[wid] = tlo;
    whi[wid] = thi;
  }
  __syncthreads();
  OwnerCounts tot{0ull, 0ull};
  for (int w = 0; w < kThreads / 32; ++w) {
    if (w < wid) {
      ex.lo += wlo[w];
      ex.hi += whi[w];
    }
    tot.lo += wlo[w];
    tot.hi += whi[w];
  }
  if (threadIdx.x == 0) {
    int o_off = 0;
    for (int o = 0; o < ESR_MAX_PEERS; ++o) {
      tile_off[o] = o_off;
      o_off += o < cyc.n ? tot.get(o) : 0;
    }
    tile_off[ESR_MAX_PEERS] = o_off;
  }
  __syncthreads();
  int run[ESR_MAX_PEERS];
#pragma unroll
  for (int o = 0; o < ESR_MAX_PEERS; ++o) run[o] = o < cyc.n ? tile_off[o] + ex.get(o) : 0;
#pragma unroll
  for (int k = 0; k < kRouteItems; ++k) {
    if (own[k] >= 0) {
      int pos = 0;
#pragma unroll
      for (int o = 0; o < ESR_MAX_PEERS; ++o)
        if (o == own[k]) pos = run[o]++;
      tile[pos] = make_int4(vi[k], vj[k], __float_as_int(vx[k]), 0);
    }
  }
  __syncthreads();
  // region of source `me` in owner's inbox: B records of 16 bytes {i, j, count bits, 0}
  const int n_tile = tile_off[ESR_MAX_PEERS];
  for (int t = threadIdx.x; t < n_tile; t += kThreads) {
    int o = 0;
#pragma unroll
    for (int q = 1; q < ESR_MAX_PEERS; ++q)
      if (t >= tile_off[q]) o = q;   // tile_off is non-decreasing; empty owners share an offset and the last match wins
    int4* dst = reinterpret_cast<int4*>(const_cast<void*>(peer_ids.p[o])) + (int64_t)me * B;
    dst[blk_base[blockIdx.x * ESR_MAX_PEERS + o] + (t - tile_off[o])] = tile[t];
  }
}

// owner side: regions -> flat [i ; j] keys of capacity 2 * B_cap, counts, n_valid; padding keys sort to the end
__global__ void __launch_bounds__(kThreads) k_pair_collect(const int4* __restrict__ in_rec,
                                                           const int32_t* __restrict__ in_counts, int n_ranks, int64_t B,
                                                           int64_t B_cap, int32_t pad_key, int32_t* __restrict__ keys,
                                                           float* __restrict__ counts, int32_t* __restrict__ n_valid,
                                                           int32_t* __restrict__ err) {
  __shared__ int64_t off[ESR_MAX_PEERS + 1];
  if (threadIdx.x == 0) {
    int64_t o = 0;
    for (int s = 0; s < n_ranks; ++s) {
      off[s] = o;
      o += min((int64_t)max(in_counts[s], 0), B);
    }
    off[n_ranks] = o;
  }
  __syncthreads();
  const int64_t m_all = off[n_ranks];
  const int64_t m = min(m_all, B_cap);
  if (blockIdx.x == 0 && threadIdx.x == 0) {
    *n_valid = (int32_t)(2 * m);
    if (m_all > B_cap) atomicOr(err, 2);
  }
  for (int64_t p = blockIdx.x * (int64_t)kThreads + threadIdx.x; p < B_cap; p += (int64_t)gridDim.x * kThreads) {
    if (p < m) {
      int s = 0;
      while (s + 1 < n_ranks && p >= off[s + 1]) ++s;
      const int64_t q = p - off[s];
      const int4 r = in_rec[(int64_t)s * B + q];
      keys[p] = r.x;
      keys[B_cap + p] = r.y;
      counts[p] = __int_as_float(r.z);
    } else {
      keys[p] = pad_key;
      keys[B_cap + p] = pad_key;
      counts[p] = 0.f;
    }
  }
}

Cyclic make_cyclic(int n_ranks) {
  Cyclic cyc;
  cyc.n = n_ranks;
  cyc.shift = -1;
  for (int b = 0; b < 4; ++b)
    if ((1 << b) == n_ranks) cyc.shift = b;
  return cyc;
}

}  // namespace
}  // namespace esr

// ---------------------------------------------------------------------------------------------
// Route plan for up to ESR_MAX_PEERS ranks without a sort: the unique rows arrive sorted, so bucketing them by owner is a
// stable partition -- tile counts, a scan over the tiles, a stable scatter (the same three small kernels as the pair
// routing above) instead of a cub radix sort + two kernels (35 -> ~12 us at 137k rows).  Same outputs, bit for bit
// (oracle/index.py route_plan): order (bucket position -> unique index), owner-local ids in bucket order, per-owner counts,
// inverse order.
// ---------------------------------------------------------------------------------------------
namespace esr {
namespace {

__global__ void __launch_bounds__(kThreads) k_route_count(const int32_t* __restrict__ uniq, const int32_t* __restrict__ n_uniq,
                                                          int64_t cap, Cyclic cyc, int32_t* __restrict__ blk_cnt) {
  __shared__ int cnt[ESR_MAX_PEERS];
  if (threadIdx.x < ESR_MAX_PEERS) cnt[threadIdx.x] = 0;
  __syncthreads();
  const int64_t n = min((int64_t)*n_uniq, cap);
  const int64_t base = (int64_t)blockIdx.x * kRouteTile + (int64_t)threadIdx.x * kRouteItems;
  OwnerCounts c{0ull, 0ull};
#pragma unroll
  for (int k = 0; k < kRouteItems; ++k)
    if (base + k < n) c.add(cyc.owner(uniq[base + k]));
  for (int o = 0; o < cyc.n; ++o) {
    int v = c.get(o);
    for (int s = 16; s > 0; s >>= 1) v += __shfl_xor_sync(FULL, v, s);
    if ((threadIdx.x & 31) == 0 && v) atomicAdd(&cnt[o], v);
  }
  __syncthreads();
  if (threadIdx.x < ESR_MAX_PEERS) blk_cnt[blockIdx.x * ESR_MAX_PEERS + threadIdx.x] = cnt[threadIdx.x];
}

// one block, warp o scans owner o over the tiles; then the bucket displacements (prefix over the owners)
__global__ void __launch_bounds__(ESR_MAX_PEERS * 32) k_route_scan(const int32_t* __restrict__ blk_cnt, int n_blocks, int n_ranks,
                                                                   int32_t* __restrict__ blk_base, int32_t* __restrict__ send_counts) {
  __shared__ int tot[ESR_MAX_PEERS];
  const int o = threadIdx.x >> 5, lane = threadIdx.x & 31;
  int carry = 0;
  if (o < n_ranks) {
    for (int b0 = 0; b0 < n_blocks; b0 += 32) {
      const int b = b0 + lane;
      const int v = b < n_blocks ? blk_cnt[b * ESR_MAX_PEERS + o] : 0;
      int x = v;
#pragma unroll
      for (int s = 1; s < 32; s <<= 1) {
        const int y = __shfl_up_sync(FULL, x, s);
        if (lane >= s) x += y;
      }
      if (b < n_blocks) blk_base[b * ESR_MAX_PEERS + o] = carry + x - v;
      carry += __shfl_sync(FULL, x, 31);
    }
  }
  if (lane == 0) tot[o] = o < n_ranks ? carry : 0;
  __syncthreads();
  if (threadIdx.x < n_ranks) send_counts[threadIdx.x] = tot[threadIdx.x];
  // fold the bucket displacement into every tile base
  int dsp = 0;
  for (int q = 0; q < o; ++q) dsp += tot[q];
  if (o < n_ranks)
    for (int b = lane; b < n_blocks; b += 32) blk_base[b * ESR_MAX_PEERS + o] += dsp;
}

__global__ void __launch_bounds__(kThreads) k_route_scatter(const int32_t* __restrict__ uniq, const int32_t* __restrict__ n_uniq,
                                                            int64_t cap, Cyclic cyc, const int32_t* __restrict__ blk_base,
                                                            int32_t* __restrict__ order, int32_t* __restrict__ send_local,
                                                            int32_t* __restrict__ inv_order) {
  __shared__ unsigned long long wlo[kThreads / 32], whi[kThreads / 32];
  const int64_t n = min((int64_t)*n_uniq, cap);
  const int64_t base = (int64_t)blockIdx.x * kRouteTile + (int64_t)threadIdx.x * kRouteItems;
  int32_t row[kRouteItems];
  int own[kRouteItems];
  OwnerCounts c{0ull, 0ull};
#pragma unroll
  for (int k = 0; k < kRouteItems; ++k) {
    own[k] = -1;
    if (base + k < n) {
      row[k] = uniq[base + k];
      own[k] = cyc.owner(row[k]);
      c.add(own[k]);
    }
  }
  unsigned long long tlo, thi;
  OwnerCounts ex;
  ex.lo = warp_excl_scan_u64(c.lo, &tlo);
  ex.hi = warp_excl_scan_u64(c.hi, &thi);
  const int wid = threadIdx.x >> 5;
  if ((threadIdx.x & 31) == 31) {
    wlo[wid] = tlo;
    whi[wid] = thi;
  }
  __syncthreads();
  for (int w = 0; w < wid; ++w) {
    ex.lo += wlo[w];
    ex.hi += whi[w];
  }
  int run[ESR_MAX_PEERS];
#pragma unroll
  for (int o = 0; o < ESR_MAX_PEERS; ++o) run[o] = o < cyc.n ? blk_base[blockIdx.x * ESR_MAX_PEERS + o] + ex.get(o) : 0;
#pragma unroll
  for (int k = 0; k < kRouteItems; ++k) {
    if (own[k] >= 0) {
      int pos = 0;
#pragma unroll
      for (int o = 0; o < ESR_MAX_PEERS; ++o)
        if (o == own[k]) pos = run[o]++;
      order[pos] = (int32_t)(base + k);
      send_local[pos] = (int32_t)cyc.local(row[k]);
      if (inv_order) inv_order[base + k] = pos;
    }
  }
}

}  // namespace

// called by esr_route_plan_i32 (shard_ops.cu) when n_ranks <= ESR_MAX_PEERS; ws: 2 * ceil(cap / 2048) * 8 ints
int route_plan_small(const int32_t* uniq, const int32_t* n_uniq, int64_t cap, int32_t n_ranks, int32_t* order,
                     int32_t* send_local, int32_t* send_counts, int32_t* inv_order, void* ws, cudaStream_t stream) {
  const int blocks = (int)ceil_div(cap, (int64_t)kRouteTile);
  Carver c(ws);
  int32_t* blk_cnt = c.take<int32_t>((size_t)blocks * ESR_MAX_PEERS);
  int32_t* blk_base = c.take<int32_t>((size_t)blocks * ESR_MAX_PEERS);
  const Cyclic cyc = make_cyclic(n_ranks);
  k_route_count<<<blocks, kThreads, 0, stream>>>(uniq, n_uniq, cap, cyc, blk_cnt);
  ESR_LAUNCH_CHECK();
  k_route_scan<<<1, ESR_MAX_PEERS * 32, 0, stream>>>(blk_cnt, blocks, n_ranks, blk_base, send_counts);
  ESR_LAUNCH_CHECK();
  k_route_scatter<<<blocks, kThreads, 0, stream>>>(uniq, n_uniq, cap, cyc, blk_base, order, send_local, inv_order);
  ESR_LAUNCH_CHECK();
  return ESR_OK;
}

}  // namespace esr

extern "C" size_t esr_peer_route_pairs_workspace_bytes(int64_t B) {
  if (B < 0) return 0;
  const int64_t blocks = ceil_div(B > 0 ? B : 1, (int64_t)kRouteTile);
  return 2 * align_up((size_t)blocks * ESR_MAX_PEERS * sizeof(int32_t), 256) + 256;
}

extern "C" int esr_peer_route_pairs_i32(const int32_t* ids, const float* counts, int64_t B, int32_t n_ranks, int32_t me,
                                        void* const* peer_pair_rec, void* const* peer_pair_counts, int32_t* my_counts,
                                        void* ws, size_t ws_bytes, esr_stream_t stream_) {
  ESR_RANGE("esr_peer_route_pairs_i32");
  ESR_REQUIRE(n_ranks >= 1 && n_ranks <= ESR_MAX_PEERS && me >= 0 && me < n_ranks && B >= 0 && B < ((int64_t)1 << 30));
  ESR_REQUIRE(my_counts != nullptr);
  PeerPtrs pi, pn;
  ESR_REQUIRE(load_ptrs(&pi, reinterpret_cast<const void* const*>(peer_pair_rec), n_ranks) &&
              load_ptrs(&pn, reinterpret_cast<const void* const*>(peer_pair_counts), n_ranks));
  for (int r = 0; r < n_ranks; ++r) ESR_REQUIRE((reinterpret_cast<uintptr_t>(pi.p[r]) % 16) == 0);
  cudaStream_t stream = static_cast<cudaStream_t>(stream_);
  const int blocks = (int)ceil_div(B > 0 ? B : 1, (int64_t)kRouteTile);
  ESR_REQUIRE(ws != nullptr && (B == 0 || (ids && counts)));
  if (ws_bytes < esr_peer_route_pairs_workspace_bytes(B)) return ESR_EWORKSPACE;
  Carver c(ws);
  int32_t* blk_cnt = c.take<int32_t>((size_t)blocks * ESR_MAX_PEERS);
  int32_t* blk_base = c.take<int32_t>((size_t)blocks * ESR_MAX_PEERS);
  const Cyclic cyc = make_cyclic(n_ranks);
  k_pair_count<<<blocks, kThreads, 0, stream>>>(ids, B, cyc, blk_cnt);
  ESR_LAUNCH_CHECK();
  k_pair_scan<<<1, ESR_MAX_PEERS * 32, 0, stream>>>(blk_cnt, blocks, n_ranks, me, blk_base, pn, my_counts);
  ESR_LAUNCH_CHECK();
  if (B > 0) {
    k_pair_scatter<<<blocks, kThreads, 0, stream>>>(ids, counts, B, cyc, me, blk_base, pi);
    ESR_LAUNCH_CHECK();
  }
  return ESR_OK;
}

extern "C" int esr_peer_collect_pairs_i32(const void* in_rec, const int32_t* in_counts, int32_t n_ranks,
                                          int64_t B, int64_t B_cap, int32_t pad_key, int32_t* keys, float* counts,
                                          int32_t* n_valid, int32_t* err, esr_stream_t stream_) {
  ESR_RANGE("esr_peer_collect_pairs_i32");
  ESR_REQUIRE(n_ranks >= 1 && n_ranks <= ESR_MAX_PEERS && B >= 0 && B_cap > 0 && B_cap < ((int64_t)1 << 30));
  ESR_REQUIRE(in_rec && in_counts && keys && counts && n_valid && err && (reinterpret_cast<uintptr_t>(in_rec) % 16) == 0);
  const int64_t want = ceil_div(B_cap, (int64_t)kThreads);
  const int64_t cap = (int64_t)sm_count() * 8;
  k_pair_collect<<<(unsigned)(want < cap ? want : cap), kThreads, 0, static_cast<cudaStream_t>(stream_)>>>(
      static_cast<const int4*>(in_rec), in_counts, n_ranks, B, B_cap, pad_key, keys, counts, n_valid, err);
  ESR_LAUNCH_CHECK();
  return ESR_OK;
}
