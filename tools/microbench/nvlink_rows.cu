// Microbenchmark (tools only, not product): how fast can 512-byte table rows cross NVLink between two B200s, by access
// shape?  One process, two devices, peer access enabled; both directions can run at once (the sharded step is symmetric).
//   pull  : reader LDGs the owner's rows (what esr_peer_gather_f32 does today), UN rows in flight per warp
//   push  : owner reads its local rows and STGs them into the reader's buffer (posted writes)
//   bulk  : cp.async.bulk (TMA, UBLKCP) peer global -> shared -> local global (pull) or local -> shared -> peer (push)
//   nvcc -O3 -std=c++17 -gencode arch=compute_100a,code=sm_100a -o tools/microbench/nvlink_rows tools/microbench/nvlink_rows.cu
#include <cstdio>
#include <cstdlib>
#include <cstdint>
#include <vector>
#include <algorithm>
#include <random>
#include <cuda_runtime.h>

#define CK(x) do { cudaError_t e = (x); if (e != cudaSuccess) { printf("CUDA error %s at line %d\n", cudaGetErrorString(e), __LINE__); exit(1); } } while (0)

__device__ __forceinline__ void stcs(float4* p, float4 v) { asm volatile("st.global.cs.v4.f32 [%0], {%1,%2,%3,%4};" ::"l"(p), "f"(v.x), "f"(v.y), "f"(v.z), "f"(v.w) : "memory"); }

// out[u] = src[idx[u]]   (src may be a peer pointer: pull; out may be a peer pointer: push)
template <int UN>
__global__ void __launch_bounds__(256) k_copy_rows(const float4* __restrict__ src, const int* __restrict__ idx, int n, float4* __restrict__ out) {
  const int lane = threadIdx.x & 31;
  const int warp = (blockIdx.x * blockDim.x + threadIdx.x) >> 5;
  const int nwarps = (gridDim.x * blockDim.x) >> 5;
  for (int u0 = warp * UN; u0 < n; u0 += nwarps * UN) {
    float4 v[UN];
    int row[UN];
#pragma unroll
    for (int k = 0; k < UN; ++k) row[k] = u0 + k < n ? idx[u0 + k] : -1;
#pragma unroll
    for (int k = 0; k < UN; ++k)
      if (row[k] >= 0) v[k] = src[(size_t)row[k] * 32 + lane];
#pragma unroll
    for (int k = 0; k < UN; ++k)
      if (row[k] >= 0) stcs(out + (size_t)(u0 + k) * 32 + lane, v[k]);
  }
}

__device__ __forceinline__ uint32_t smem_u32(const void* p) { return (uint32_t)__cvta_generic_to_shared(p); }
__device__ __forceinline__ void mbar_init(uint32_t bar, uint32_t count) { asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(bar), "r"(count) : "memory"); }
__device__ __forceinline__ void mbar_expect_tx(uint32_t bar, uint32_t bytes) { asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(bar), "r"(bytes) : "memory"); }
__device__ __forceinline__ void mbar_wait(uint32_t bar, uint32_t parity) {
  uint32_t ok = 0;
  for (uint32_t spin = 0; !ok; ++spin) {
    asm volatile("{\n .reg .pred p;\n mbarrier.try_wait.parity.shared::cta.b64 p, [%1], %2;\n selp.u32 %0, 1, 0, p;\n}\n" : "=r"(ok) : "r"(bar), "r"(parity) : "memory");
    if (!ok && spin > (1u << 26)) __trap();
  }
}
__device__ __forceinline__ void bulk_g2s(uint32_t dst, const void* src, uint32_t bytes, uint32_t bar) {
  asm volatile("cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1], %2, [%3];" ::"r"(dst), "l"(src), "r"(bytes), "r"(bar) : "memory");
}
__device__ __forceinline__ void bulk_s2g(void* dst, uint32_t src, uint32_t bytes) {
  asm volatile("cp.async.bulk.global.shared::cta.bulk_group [%0], [%1], %2;" ::"l"(dst), "r"(src), "r"(bytes) : "memory");
}

// Bulk-copy ring: one warp per ring of NS stages of RPS rows each; lane 0..RPS-1 issue one 512-byte row copy each into
// the stage (rows are scattered in the source), then ONE bulk store of the whole stage (RPS * 512 contiguous bytes of out).
template <int NS, int RPS>
__global__ void __launch_bounds__(128) k_bulk_rows(const float4* __restrict__ src, const int* __restrict__ idx, int n, float4* __restrict__ out) {
  extern __shared__ __align__(128) unsigned char smem[];
  __shared__ __align__(8) uint64_t bars[4][NS];
  const int lane = threadIdx.x & 31, wid = threadIdx.x >> 5;
  unsigned char* ring = smem + (size_t)wid * NS * RPS * 512;
  const uint32_t ring_u32 = smem_u32(ring);
  const uint32_t bar0 = smem_u32(&bars[wid][0]);
  if (lane == 0) {
    for (int i = 0; i < NS; ++i) mbar_init(bar0 + 8u * i, 1);
    asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
  }
  __syncwarp();
  const int warp = blockIdx.x * 4 + wid, nwarps = gridDim.x * 4;
  const int ngroups = (n + RPS - 1) / RPS;
  uint32_t phase = 0;
  int issued = 0, done = 0;
  // groups of this warp: g = warp + k * nwarps
  auto issue = [&](int k) {
    const int g = warp + k * nwarps;
    const int stg = k % NS;
    const int r0 = g * RPS, cnt = min(RPS, n - r0);
    if (lane == 0) mbar_expect_tx(bar0 + 8u * stg, 512u * cnt);
    __syncwarp();
    if (lane < cnt) bulk_g2s(ring_u32 + (uint32_t)(stg * RPS + lane) * 512u, src + (size_t)idx[r0 + lane] * 32, 512u, bar0 + 8u * stg);
  };
  const int mine = warp < ngroups ? (ngroups - warp + nwarps - 1) / nwarps : 0;
  for (; issued < mine && issued < NS; ++issued) issue(issued);
  for (; done < mine; ++done) {
    const int stg = done % NS;
    mbar_wait(bar0 + 8u * stg, (phase >> stg) & 1u);
    phase ^= 1u << stg;
    const int g = warp + done * nwarps;
    const int r0 = g * RPS, cnt = min(RPS, n - r0);
    if (lane == 0) {
      asm volatile("fence.proxy.async.shared::cta;" ::: "memory");
      bulk_s2g(out + (size_t)r0 * 32, ring_u32 + (uint32_t)(stg * RPS) * 512u, 512u * cnt);
      asm volatile("cp.async.bulk.commit_group;" ::: "memory");
    }
    if (issued < mine) {
      // the stage being refilled is `issued % NS` == stg: its store must have finished READING shared memory
      if (lane == 0) asm volatile("cp.async.bulk.wait_group.read 0;" ::: "memory");
      __syncwarp();
      issue(issued);
      ++issued;
    }
  }
  if (lane == 0) asm volatile("cp.async.bulk.wait_group 0;" ::: "memory");
}

struct Dev {
  int id;
  float4 *table, *out;
  int *idx;
  cudaStream_t st;
  cudaEvent_t e0, e1;
};

int main(int argc, char** argv) {
  int ndev = 0;
  CK(cudaGetDeviceCount(&ndev));
  if (ndev < 2) { printf("needs 2 GPUs\n"); return 0; }
  const int V = 12500000;            // 6.4 GB shard (100M rows / 8)
  const int N = argc > 1 ? atoi(argv[1]) : 120000;  // rows per transfer (~61 MB)
  Dev d[2];
  std::mt19937 rng(3);
  std::vector<int> h(N);
  for (int g = 0; g < 2; ++g) {
    d[g].id = g;
    CK(cudaSetDevice(g));
    int can = 0;
    CK(cudaDeviceCanAccessPeer(&can, g, 1 - g));
    if (!can) { printf("no peer access\n"); return 0; }
    CK(cudaDeviceEnablePeerAccess(1 - g, 0));
    CK(cudaMalloc(&d[g].table, (size_t)V * 512));
    CK(cudaMemset(d[g].table, 1, (size_t)V * 512));
    CK(cudaMalloc(&d[g].out, (size_t)N * 512));
    CK(cudaMalloc(&d[g].idx, N * 4));
    for (auto& x : h) x = rng() % V;
    std::sort(h.begin(), h.end());
    CK(cudaMemcpy(d[g].idx, h.data(), N * 4, cudaMemcpyHostToDevice));
    CK(cudaStreamCreate(&d[g].st));
    CK(cudaEventCreate(&d[g].e0));
    CK(cudaEventCreate(&d[g].e1));
  }
  int sms = 0;
  CK(cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, 0));
  const double bytes = (double)N * 512;
  // launch(g, src_dev, dst_dev): device g runs a kernel reading table of src_dev into out of dst_dev
  auto run = [&](const char* name, bool both, auto launch) {
    float best[2] = {1e9f, 1e9f};
    for (int rep = 0; rep < 6; ++rep) {
      for (int g = 0; g < 2; ++g) { CK(cudaSetDevice(g)); CK(cudaDeviceSynchronize()); }
      for (int g = 0; g < (both ? 2 : 1); ++g) {
        CK(cudaSetDevice(g));
        CK(cudaEventRecord(d[g].e0, d[g].st));
        launch(g);
        CK(cudaEventRecord(d[g].e1, d[g].st));
      }
      for (int g = 0; g < (both ? 2 : 1); ++g) {
        CK(cudaSetDevice(g));
        CK(cudaStreamSynchronize(d[g].st));
        CK(cudaGetLastError());
        float ms;
        CK(cudaEventElapsedTime(&ms, d[g].e0, d[g].e1));
        if (rep >= 1) best[g] = std::min(best[g], ms);
      }
    }
    const float t = both ? std::max(best[0], best[1]) : best[0];
    printf("%-58s %s  %7.1f us  -> %6.1f GB/s per direction\n", name, both ? "both dirs" : "one dir  ", t * 1e3, bytes / (t * 1e-3) / 1e9);
    fflush(stdout);
  };
  for (int both = 0; both < 2; ++both) {
    for (int per_sm : {4, 8}) {
      char nm[128];
#define PULLPUSH(UN)                                                                                                       \
      snprintf(nm, sizeof nm, "pull LDG  UN=%d ctas/sm=%d", UN, per_sm);                                                     \
      run(nm, both, [&](int g) { k_copy_rows<UN><<<sms * per_sm, 256, 0, d[g].st>>>(d[1 - g].table, d[g].idx, N, d[g].out); }); \
      snprintf(nm, sizeof nm, "push STG  UN=%d ctas/sm=%d", UN, per_sm);                                                     \
      run(nm, both, [&](int g) { k_copy_rows<UN><<<sms * per_sm, 256, 0, d[g].st>>>(d[g].table, d[g].idx, N, d[1 - g].out); });
      PULLPUSH(2) PULLPUSH(4) PULLPUSH(8)
    }
    {
      constexpr int NS = 4, RPS = 8;
      const size_t smem = 4 * NS * RPS * 512;
      CK(cudaSetDevice(0)); CK(cudaFuncSetAttribute(k_bulk_rows<NS, RPS>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
      CK(cudaSetDevice(1)); CK(cudaFuncSetAttribute(k_bulk_rows<NS, RPS>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
      for (int per_sm : {1, 2, 3}) {
        char nm[128];
        snprintf(nm, sizeof nm, "pull bulk (TMA) NS=%d RPS=%d ctas/sm=%d", NS, RPS, per_sm);
        run(nm, both, [&](int g) { k_bulk_rows<NS, RPS><<<sms * per_sm, 128, smem, d[g].st>>>(d[1 - g].table, d[g].idx, N, d[g].out); });
        snprintf(nm, sizeof nm, "push bulk (TMA) NS=%d RPS=%d ctas/sm=%d", NS, RPS, per_sm);
        run(nm, both, [&](int g) { k_bulk_rows<NS, RPS><<<sms * per_sm, 128, smem, d[g].st>>>(d[g].table, d[g].idx, N, d[1 - g].out); });
      }
    }
    // local gather for scale (same kernel, both pointers local)
    run("local LDG UN=4 ctas/sm=8", both, [&](int g) { k_copy_rows<4><<<sms * 8, 256, 0, d[g].st>>>(d[g].table, d[g].idx, N, d[g].out); });
    // cudaMemcpyPeer of the same byte count, contiguous
    run("cudaMemcpyPeerAsync contiguous", both, [&](int g) { CK(cudaMemcpyPeerAsync(d[g].out, g, d[1 - g].table, 1 - g, (size_t)N * 512, d[g].st)); });
  }
  return 0;
}
