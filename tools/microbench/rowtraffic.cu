// Microbenchmark (tools only, not product): what HBM rate can the ROW PASS's access pattern reach on a B200 when nothing
// but the memory system limits it?  Same table geometry as bench.py (V rows x D floats, two row buffers + accumulator),
// same index statistics (sorted unique self rows, random partner rows), no arithmetic dependencies between rows.
//   A: per unique row: read row + acc, write row' + acc          (4*U*R bytes: the algorithmic bytes of SURVEY 8(d))
//   B: A + one partner-row read per slot (random rows of the touched set)
//   C: partner reads only
// Each variant is run with several (rows in flight per warp, warps per SM) shapes.
//   nvcc -O3 -gencode arch=compute_100a,code=sm_100a -o tools/microbench/rowtraffic tools/microbench/rowtraffic.cu
#include <cstdio>
#include <cstdlib>
#include <cstdint>
#include <vector>
#include <algorithm>
#include <random>
#include <cuda_runtime.h>

#define CK(x) do { cudaError_t e = (x); if (e != cudaSuccess) { printf("CUDA error %s at %d\n", cudaGetErrorString(e), __LINE__); exit(1); } } while (0)

__device__ __forceinline__ float4 ldcs(const float4* p) { float4 v; asm volatile("ld.global.cs.v4.f32 {%0,%1,%2,%3}, [%4];" : "=f"(v.x), "=f"(v.y), "=f"(v.z), "=f"(v.w) : "l"(p)); return v; }
__device__ __forceinline__ void stcs(float4* p, float4 v) { asm volatile("st.global.cs.v4.f32 [%0], {%1,%2,%3,%4};" ::"l"(p), "f"(v.x), "f"(v.y), "f"(v.z), "f"(v.w) : "memory"); }

// warp per row, lane owns one float4 (D = 128); UN rows in flight per warp
template <int UN, bool UPDATE, bool PARTNER>
__global__ void __launch_bounds__(256) k_traffic(const float4* __restrict__ rows0, float4* __restrict__ rows1, float4* __restrict__ acc,
                                                 const int* __restrict__ uniq, int U, const int* __restrict__ partner, int nslots,
                                                 float* __restrict__ sink) {
  const int lane = threadIdx.x & 31;
  const int warp = (blockIdx.x * blockDim.x + threadIdx.x) >> 5;
  const int nwarps = (gridDim.x * blockDim.x) >> 5;
  float4 s = make_float4(0.f, 0.f, 0.f, 0.f);
  if (UPDATE) {
    for (int u0 = warp * UN; u0 < U; u0 += nwarps * UN) {
      float4 r[UN], a[UN];
      int row[UN];
#pragma unroll
      for (int k = 0; k < UN; ++k) row[k] = u0 + k < U ? uniq[u0 + k] : -1;
#pragma unroll
      for (int k = 0; k < UN; ++k)
        if (row[k] >= 0) { r[k] = rows0[(size_t)row[k] * 32 + lane]; a[k] = ldcs(acc + (size_t)row[k] * 32 + lane); }
#pragma unroll
      for (int k = 0; k < UN; ++k)
        if (row[k] >= 0) {
          a[k].x += r[k].x * r[k].x; a[k].y += r[k].y; a[k].z += r[k].z; a[k].w += r[k].w;
          r[k].x -= a[k].x * 1e-9f;
          stcs(rows1 + (size_t)row[k] * 32 + lane, r[k]);
          stcs(acc + (size_t)row[k] * 32 + lane, a[k]);
        }
    }
  }
  if (PARTNER) {
    for (int p0 = warp * UN; p0 < nslots; p0 += nwarps * UN) {
      float4 r[UN];
      int row[UN];
#pragma unroll
      for (int k = 0; k < UN; ++k) row[k] = p0 + k < nslots ? partner[p0 + k] : -1;
#pragma unroll
      for (int k = 0; k < UN; ++k)
        if (row[k] >= 0) r[k] = rows0[(size_t)row[k] * 32 + lane];
#pragma unroll
      for (int k = 0; k < UN; ++k)
        if (row[k] >= 0) { s.x += r[k].x; s.y += r[k].y; s.z += r[k].z; s.w += r[k].w; }
    }
  }
  if (s.x + s.y + s.z + s.w == 123.456f) sink[0] = s.x;
}

// interleaved: a warp takes a sorted range of slots like the row pass: for each unique row, its update traffic, plus the partner
// reads of that row's slots (slots sorted by self row; seg_off gives each row's slot range)
template <int UN>
__global__ void __launch_bounds__(256) k_interleaved(const float4* __restrict__ rows0, float4* __restrict__ rows1, float4* __restrict__ acc,
                                                     const int* __restrict__ uniq, const int* __restrict__ seg_off, int U,
                                                     const int* __restrict__ partner_sorted, float* __restrict__ sink, int* counter, int item_rows) {
  const int lane = threadIdx.x & 31;
  float4 s = make_float4(0.f, 0.f, 0.f, 0.f);
  for (;;) {
    int item = 0;
    if (lane == 0) item = atomicAdd(counter, 1);
    item = __shfl_sync(0xffffffffu, item, 0);
    const int u_begin = item * item_rows;
    if (u_begin >= U) break;
    const int u_end = min(U, u_begin + item_rows);
    for (int u0 = u_begin; u0 < u_end; u0 += UN) {
      float4 r[UN], a[UN], p[UN];
      int row[UN], q[UN];
#pragma unroll
      for (int k = 0; k < UN; ++k) {
        row[k] = u0 + k < u_end ? uniq[u0 + k] : -1;
        q[k] = u0 + k < u_end ? partner_sorted[seg_off[u0 + k]] : -1;  // first slot's partner (uniform: ~1.3 slots per row)
      }
#pragma unroll
      for (int k = 0; k < UN; ++k)
        if (row[k] >= 0) {
          r[k] = rows0[(size_t)row[k] * 32 + lane];
          a[k] = ldcs(acc + (size_t)row[k] * 32 + lane);
          p[k] = rows0[(size_t)q[k] * 32 + lane];
        }
#pragma unroll
      for (int k = 0; k < UN; ++k)
        if (row[k] >= 0) {
          // remaining slots of the row (beyond the first): serial partner reads
          for (int t = seg_off[u0 + k] + 1; t < seg_off[u0 + k + 1]; ++t) {
            const float4 x = rows0[(size_t)partner_sorted[t] * 32 + lane];
            s.x += x.x; s.y += x.y;
          }
          a[k].x += p[k].x * r[k].x; a[k].y += p[k].y; a[k].z += r[k].z; a[k].w += r[k].w;
          r[k].x -= a[k].x * 1e-9f;
          stcs(rows1 + (size_t)row[k] * 32 + lane, r[k]);
          stcs(acc + (size_t)row[k] * 32 + lane, a[k]);
        }
    }
  }
  if (s.x + s.y + s.z + s.w == 123.456f) sink[0] = s.x;
}

int main(int argc, char** argv) {
  const int V = 1000000, D4 = 32, B = 262144;
  const int zipf = argc > 1 ? atoi(argv[1]) : 0;
  std::mt19937_64 rng(1);
  std::vector<int> ids(2 * B);
  if (zipf) {
    std::vector<double> cdf(V);
    double h = 0;
    for (int k = 0; k < V; ++k) { h += 1.0 / (k + 1); cdf[k] = h; }
    std::uniform_real_distribution<double> ud(0.0, h);
    for (auto& x : ids) x = (int)(std::lower_bound(cdf.begin(), cdf.end(), ud(rng)) - cdf.begin());
  } else {
    std::uniform_int_distribution<int> ud(0, V - 1);
    for (auto& x : ids) x = ud(rng);
  }
  // slots sorted by self row; partner of slot s is ids[s ^ B-offset]
  std::vector<int> order(2 * B);
  for (int i = 0; i < 2 * B; ++i) order[i] = i;
  std::stable_sort(order.begin(), order.end(), [&](int a, int b) { return ids[a] < ids[b]; });
  std::vector<int> uniq, seg_off, partner_sorted(2 * B);
  for (int p = 0; p < 2 * B; ++p) {
    const int s = order[p];
    if (p == 0 || ids[s] != ids[order[p - 1]]) { uniq.push_back(ids[s]); seg_off.push_back(p); }
    partner_sorted[p] = ids[s < B ? s + B : s - B];
  }
  const int U = (int)uniq.size();
  seg_off.push_back(2 * B);
  printf("stream=%s U=%d slots=%d alg_bytes=%.1f MB\n", zipf ? "zipf" : "uniform", U, 2 * B, 4.0 * U * 512 / 1e6);
  float4 *rows0, *rows1, *acc;
  const size_t tb = (size_t)V * D4 * sizeof(float4);
  CK(cudaMalloc(&rows0, tb)); CK(cudaMalloc(&rows1, tb)); CK(cudaMalloc(&acc, tb));
  CK(cudaMemset(rows0, 0, tb)); CK(cudaMemset(rows1, 0, tb)); CK(cudaMemset(acc, 0, tb));
  int *d_uniq, *d_part, *d_seg, *d_counter; float* sink;
  CK(cudaMalloc(&d_uniq, U * 4)); CK(cudaMalloc(&d_part, 2 * B * 4)); CK(cudaMalloc(&d_seg, (U + 1) * 4)); CK(cudaMalloc(&d_counter, 4)); CK(cudaMalloc(&sink, 4));
  CK(cudaMemcpy(d_uniq, uniq.data(), U * 4, cudaMemcpyHostToDevice));
  CK(cudaMemcpy(d_part, partner_sorted.data(), 2 * B * 4, cudaMemcpyHostToDevice));
  CK(cudaMemcpy(d_seg, seg_off.data(), (U + 1) * 4, cudaMemcpyHostToDevice));
  // a 1 GiB flush buffer between repetitions so that nothing survives in L2
  char* flush; CK(cudaMalloc(&flush, 1u << 30));
  cudaEvent_t e0, e1; CK(cudaEventCreate(&e0)); CK(cudaEventCreate(&e1));
  auto timeit = [&](const char* name, double bytes, auto launch) {
    float best = 1e9f, tot = 0.f;
    const int reps = 6;
    for (int r = 0; r < reps + 2; ++r) {
      CK(cudaMemsetAsync(flush, r, 1u << 30));
      CK(cudaMemsetAsync(d_counter, 0, 4));
      CK(cudaEventRecord(e0));
      launch();
      CK(cudaEventRecord(e1));
      CK(cudaDeviceSynchronize());
      float ms; CK(cudaEventElapsedTime(&ms, e0, e1));
      if (r >= 2) { best = std::min(best, ms); tot += ms; }
    }
    printf("%-44s best %7.1f us  avg %7.1f us  -> %6.0f GB/s (of the bytes named)\n", name, best * 1e3, tot / reps * 1e3, bytes / (best * 1e-3) / 1e9);
  };
  const double bytesA = 4.0 * U * 512, bytesP = 2.0 * B * 512;
  int sms = 0; CK(cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, 0));
  for (int per_sm : {4, 8}) {
    const int grid = sms * per_sm;
    char nm[128];
#define RUN(UN)                                                                                                         \
    snprintf(nm, sizeof nm, "A update-only      UN=%d ctas/sm=%d", UN, per_sm);                                            \
    timeit(nm, bytesA, [&] { k_traffic<UN, true, false><<<grid, 256>>>(rows0, rows1, acc, d_uniq, U, d_part, 2 * B, sink); }); \
    snprintf(nm, sizeof nm, "C partner-only     UN=%d ctas/sm=%d", UN, per_sm);                                            \
    timeit(nm, bytesP, [&] { k_traffic<UN, false, true><<<grid, 256>>>(rows0, rows1, acc, d_uniq, U, d_part, 2 * B, sink); }); \
    snprintf(nm, sizeof nm, "B update+partner   UN=%d ctas/sm=%d (alg bytes)", UN, per_sm);                                \
    timeit(nm, bytesA, [&] { k_traffic<UN, true, true><<<grid, 256>>>(rows0, rows1, acc, d_uniq, U, d_part, 2 * B, sink); }); \
    snprintf(nm, sizeof nm, "I interleaved      UN=%d ctas/sm=%d (alg bytes)", UN, per_sm);                                \
    timeit(nm, bytesA, [&] { k_interleaved<UN><<<grid, 256>>>(rows0, rows1, acc, d_uniq, d_seg, U, d_part, sink, d_counter, 64); });
    RUN(2) RUN(4) RUN(8)
  }
  // plain copy of the same number of bytes, for scale (what MEASURED_PEAKS.json calls the HBM peak)
  timeit("copy 512 MB (read+write bytes)", 2.0 * tb, [&] { CK(cudaMemcpyAsync(rows1, rows0, tb, cudaMemcpyDeviceToDevice)); });
  return 0;
}
