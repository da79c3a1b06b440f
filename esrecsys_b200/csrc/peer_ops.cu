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

template <int TPR, int ROWS>
__global__ void __launch_bounds__(kThreads) k_peer_gather(PeerPtrs rows, PeerPtrs bias, const int32_t* __restrict__ uniq,
                                                          const int32_t* __restrict__ n_uniq, int64_t cap, int n_ranks,
                                                          int D4, float4* __restrict__ out, float* __restrict__ out_bias) {
  const int lane = threadIdx.x % TPR;
  const int64_t group = (blockIdx.x * (int64_t)kThreads + threadIdx.x) / TPR;
  const int64_t n = min((int64_t)*n_uniq, cap);
  const int64_t first = group * ROWS;
  const float4* src[ROWS];
#pragma unroll
  for (int r = 0; r < ROWS; ++r) {
    const int64_t u = first + r;
    src[r] = nullptr;
    if (u < n) {
      const int32_t row = uniq[u];
      const int owner = row % n_ranks;
      const int64_t local = row / n_ranks;
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
      if (src[r]) out[(first + r) * D4 + c] = v[r];
  }
}

// src_meta[s] = {offset of source s in recv_ids, count, displacement inside source s's bucket list}
__global__ void __launch_bounds__(kThreads) k_peer_pull_ids(PeerPtrs counts, PeerPtrs send_local, int n_ranks, int me,
                                                            int64_t recv_cap, int32_t* __restrict__ recv_ids,
                                                            int32_t* __restrict__ src_meta) {
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
    }
  }
  const int64_t total = min((int64_t)off[n_ranks], recv_cap);
  for (int64_t k = blockIdx.x * (int64_t)kThreads + threadIdx.x; k < total; k += (int64_t)gridDim.x * kThreads) {
    int s = 0;
    while (s + 1 < n_ranks && k >= off[s + 1]) ++s;
    recv_ids[k] = reinterpret_cast<const int32_t*>(send_local.p[s])[dsp[s] + (k - off[s])];
  }
}

// index of x in the ascending array a[0..n), or -1
__device__ __forceinline__ int find_sorted(const int32_t* __restrict__ a, int n, int32_t x) {
  int lo = 0, hi = n;
  while (lo < hi) {
    const int mid = (lo + hi) >> 1;
    if (a[mid] < x) lo = mid + 1;
    else hi = mid;
  }
  return (lo < n && a[lo] == x) ? lo : -1;
}

template <int TPR>
__global__ void __launch_bounds__(kThreads) k_peer_merge_adagrad(PeerPtrs order, PeerPtrs dE, PeerPtrs db, int n_ranks,
                                                                 const int32_t* __restrict__ recv_ids,
                                                                 const int32_t* __restrict__ src_meta, int D4,
                                                                 float* __restrict__ rows, float* __restrict__ acc,
                                                                 float* __restrict__ bias, float* __restrict__ bias_acc,
                                                                 float lr, float eps) {
  const int lane = threadIdx.x % TPR;
  const int64_t total = src_meta[n_ranks * 3];
  const int64_t groups = (int64_t)gridDim.x * (kThreads / TPR);
  for (int64_t k = (blockIdx.x * (int64_t)kThreads + threadIdx.x) / TPR; k < total; k += groups) {
    int s = 0;
    while (s + 1 < n_ranks && k >= src_meta[(s + 1) * 3]) ++s;
    const int32_t x = recv_ids[k];
    bool first = true;  // is s the first source that names row x?
    for (int q = 0; q < s && first; ++q)
      first = find_sorted(recv_ids + src_meta[q * 3], src_meta[q * 3 + 1], x) < 0;
    if (!first) continue;
    // gradient row index of x inside each source's dE (-1: that source does not name x):
    // order_q[displ_q + position of x in source q's list], for this source and the later ones
    int gi[ESR_MAX_PEERS];
#pragma unroll
    for (int q = 0; q < ESR_MAX_PEERS; ++q) {
      gi[q] = -1;
      if (q < n_ranks && q >= s) {
        const int pos = q == s ? (int)(k - src_meta[s * 3])
                               : find_sorted(recv_ids + src_meta[q * 3], src_meta[q * 3 + 1], x);
        if (pos >= 0) gi[q] = reinterpret_cast<const int32_t*>(order.p[q])[src_meta[q * 3 + 2] + pos];
      }
    }
    float4* p = reinterpret_cast<float4*>(rows) + (int64_t)x * D4;
    float4* a = reinterpret_cast<float4*>(acc) + (int64_t)x * D4;
    for (int c = lane; c < D4; c += TPR) {
      float4 g = f4_zero();
#pragma unroll
      for (int q = 0; q < ESR_MAX_PEERS; ++q)
        if (q < n_ranks && gi[q] >= 0) f4_add(g, reinterpret_cast<const float4*>(dE.p[q])[(int64_t)gi[q] * D4 + c]);
      float4 pv = p[c], av = ld_stream(a + c);
      adagrad4(pv, av, g, lr, eps);
      p[c] = pv;
      st_stream(a + c, av);
    }
    if (lane == 0) {
      float g = 0.f;
#pragma unroll
      for (int q = 0; q < ESR_MAX_PEERS; ++q)
        if (q < n_ranks && gi[q] >= 0) g += reinterpret_cast<const float*>(db.p[q])[gi[q]];
      float pv = bias[x], av = bias_acc[x];
      adagrad1(pv, av, g, lr, eps);
      bias[x] = pv;
      bias_acc[x] = av;
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
  const unsigned grid = (unsigned)ceil_div(ceil_div(cap, ROWS) * tpr, kThreads);
  ESR_DISPATCH_TPR(tpr, (k_peer_gather<TPR, ROWS><<<grid, kThreads, 0, static_cast<cudaStream_t>(stream_)>>>(
                            pr, pb, uniq, n_uniq, cap, n_ranks, D4, reinterpret_cast<float4*>(out), out_bias)));
  ESR_LAUNCH_CHECK();
  return ESR_OK;
}

extern "C" int esr_peer_pull_ids_i32(const void* const* peer_counts, const void* const* peer_send_local, int32_t n_ranks,
                                     int32_t me, int64_t recv_cap, int32_t* recv_ids, int32_t* src_meta,
                                     esr_stream_t stream_) {
  ESR_REQUIRE(n_ranks >= 1 && n_ranks <= ESR_MAX_PEERS && me >= 0 && me < n_ranks && recv_cap > 0 && recv_ids && src_meta);
  PeerPtrs pc, ps;
  ESR_REQUIRE(load_ptrs(&pc, peer_counts, n_ranks) && load_ptrs(&ps, peer_send_local, n_ranks));
  const int grid = 2 * sm_count();
  k_peer_pull_ids<<<grid, kThreads, 0, static_cast<cudaStream_t>(stream_)>>>(pc, ps, n_ranks, me, recv_cap, recv_ids,
                                                                             src_meta);
  ESR_LAUNCH_CHECK();
  return ESR_OK;
}

extern "C" int esr_peer_merge_adagrad_f32(EsrTable* shard, const void* const* peer_order, const void* const* peer_dE,
                                          const void* const* peer_db, int32_t n_ranks, const int32_t* recv_ids,
                                          const int32_t* src_meta, float lr, float eps, esr_stream_t stream_) {
  ESR_REQUIRE(shard && shard->struct_size >= sizeof(EsrTable) && shard->D > 0 && (shard->D % 4) == 0);
  ESR_REQUIRE(shard->rows[0] && shard->acc && shard->bias && shard->bias_acc && shard->ver == nullptr);
  ESR_REQUIRE(n_ranks >= 1 && n_ranks <= ESR_MAX_PEERS && recv_ids && src_meta);
  PeerPtrs po, pe, pb;
  ESR_REQUIRE(load_ptrs(&po, peer_order, n_ranks) && load_ptrs(&pe, peer_dE, n_ranks) && load_ptrs(&pb, peer_db, n_ranks));
  const int D4 = shard->D / 4;
  const int tpr = tpr_for(D4);
  const int grid = 8 * sm_count();
  ESR_DISPATCH_TPR(tpr, (k_peer_merge_adagrad<TPR><<<grid, kThreads, 0, static_cast<cudaStream_t>(stream_)>>>(
                            po, pe, pb, n_ranks, recv_ids, src_meta, D4, shard->rows[0], shard->acc, shard->bias,
                            shard->bias_acc, lr, eps)));
  ESR_LAUNCH_CHECK();
  return ESR_OK;
}
