// Native host runtime of the single-GPU training loop: the body of train_epoch (wikipedia/train_cooccurence.py:103-112:
// next(train_it) -> apply_model -> update_model) as ONE C call per step.
//
// A step is three stream-ordered stages on three streams owned by this object
//   copy stream : stage the batch (H2D from pinned host memory on a copy engine, or D2D)
//   side stream : index plan of the batch (ids only)                       -- a captured CUDA graph per buffer parity
//   main stream : prep -> rows -> combine -> finish on the table           -- a captured CUDA graph per buffer parity
// with `depth` staging / plan buffers in flight (3: batch t+2 uploads while plan(t+1) is built and step(t) trains).
// Round 1 drove this from Python through torch (stream context managers, tensor.copy_, torch events, CUDAGraph.replay):
// ~14 framework calls, 170-200 us of host time per step -- more than the 157 us the GPU needs, so the end-to-end rate
// from host batches was HOST-bound (1.35 G pairs/s against 1.67 G device-timed).  Here a step is ~12 CUDA runtime calls.
// The kernels are captured while the caller launches them through the usual libesr entry points on these streams
// (esr_pipeline_capture_begin / _end bracket the launches), so no kernel knowledge lives here.
#include <cstdlib>
#include <new>

#include "esr_common.cuh"

namespace {
// Host batches are staged by a FEW CTAs reading the pinned block over PCIe (UVA maps pinned host memory into the device
// address space), not by the copy engine.  Measured (esr_pipeline_trace, profiles/r2_pipeline_timeline.txt): while a 3 MB
// upload runs at full PCIe rate -- DMA or 24 CTAs alike, ~60 us -- the plan / step graphs of the steps in flight start
// 20-35 us late (their launches also have to come over PCIe), 184 us per step against 149 us from device-resident batches;
// throttled to 2-4 CTAs the upload takes ~110 us, still hidden behind the 149 us step, and a step costs 158-165 us.
__global__ void __launch_bounds__(256) k_stage_copy(const int4* __restrict__ src, int4* __restrict__ dst, int64_t n16) {
  const int64_t stride = (int64_t)gridDim.x * blockDim.x;
  int64_t i = blockIdx.x * (int64_t)blockDim.x + threadIdx.x;
  for (; i + 3 * stride < n16; i += 4 * stride) {
    const int4 a = src[i], b = src[i + stride], c = src[i + 2 * stride], d = src[i + 3 * stride];
    dst[i] = a;
    dst[i + stride] = b;
    dst[i + 2 * stride] = c;
    dst[i + 3 * stride] = d;
  }
  for (; i < n16; i += stride) dst[i] = src[i];
}

int stage(void* dst, const void* src, size_t bytes, bool by_kernel, cudaStream_t stream) {
  if (by_kernel && (bytes % 16) == 0 && (reinterpret_cast<uintptr_t>(dst) % 16) == 0 && (reinterpret_cast<uintptr_t>(src) % 16) == 0) {
    static const int ctas = getenv("ESR_PIPE_STAGE_CTAS") ? atoi(getenv("ESR_PIPE_STAGE_CTAS")) : 3;  // throttled on purpose: see the comment on k_stage_copy
    static const int thr = getenv("ESR_PIPE_STAGE_THREADS") ? atoi(getenv("ESR_PIPE_STAGE_THREADS")) : 256;
    k_stage_copy<<<ctas, thr, 0, stream>>>(static_cast<const int4*>(src), static_cast<int4*>(dst), (int64_t)(bytes / 16));
    ESR_LAUNCH_CHECK();
    return ESR_OK;
  }
  ESR_CUDA(cudaMemcpyAsync(dst, src, bytes, cudaMemcpyDefault, stream));
  return ESR_OK;
}

constexpr int kMaxDepth = 4;
constexpr uint32_t kMagic = 0x45535250u;  // "ESRP"
}  // namespace

struct EsrPipeline {
  uint32_t magic;
  int depth;
  cudaStream_t copy, side, main, d2h;
  cudaEvent_t ev_in, ev_copy[kMaxDepth], ev_plan[kMaxDepth], ev_done[kMaxDepth];
  cudaGraphExec_t g_plan[kMaxDepth], g_step[kMaxDepth];
  void* ids_dev[kMaxDepth];
  void* counts_dev[kMaxDepth];
  size_t ids_bytes, counts_bytes;
  const float* loss_src;
  float* loss_log;
  float* loss_host;
  int64_t loss_len;
  int64_t t;
  int capturing;  // -1 none, 0 plan, 1 step
  // optional timeline (esr_pipeline_trace): timing events around the three stages of the next steps
  static constexpr int kTrace = 64;
  int trace_left, trace_n;
  cudaEvent_t tr_base, tr_ev[kTrace][6];  // copy begin/end, plan begin/end, step begin/end
};

using namespace esr;

static bool ok(const EsrPipeline* p) { return p != nullptr && p->magic == kMagic; }

extern "C" int esr_pipeline_create(int32_t depth, int32_t main_high_priority, EsrPipeline** out) {
  ESR_REQUIRE(out != nullptr && depth >= 1 && depth <= kMaxDepth);
  EsrPipeline* p = new (std::nothrow) EsrPipeline();
  if (p == nullptr) return ESR_ENOMEM;
  p->magic = kMagic;
  p->depth = depth;
  p->capturing = -1;
  p->trace_left = p->trace_n = 0;
  p->tr_base = nullptr;
  int lo = 0, hi = 0;
  ESR_CUDA(cudaDeviceGetStreamPriorityRange(&lo, &hi));
  ESR_CUDA(cudaStreamCreateWithPriority(&p->copy, cudaStreamNonBlocking, lo));
  static const bool side_hi = getenv("ESR_PIPE_SIDE_HI") != nullptr;  // A/B probe: plan stream above the step stream
  ESR_CUDA(cudaStreamCreateWithPriority(&p->side, cudaStreamNonBlocking, side_hi ? hi : lo));
  ESR_CUDA(cudaStreamCreateWithPriority(&p->main, cudaStreamNonBlocking, main_high_priority ? hi : lo));
  ESR_CUDA(cudaStreamCreateWithPriority(&p->d2h, cudaStreamNonBlocking, lo));
  ESR_CUDA(cudaEventCreateWithFlags(&p->ev_in, cudaEventDisableTiming));
  for (int k = 0; k < kMaxDepth; ++k) {
    ESR_CUDA(cudaEventCreateWithFlags(&p->ev_copy[k], cudaEventDisableTiming));
    ESR_CUDA(cudaEventCreateWithFlags(&p->ev_plan[k], cudaEventDisableTiming));
    ESR_CUDA(cudaEventCreateWithFlags(&p->ev_done[k], cudaEventDisableTiming));
    p->g_plan[k] = p->g_step[k] = nullptr;
    p->ids_dev[k] = p->counts_dev[k] = nullptr;
  }
  *out = p;
  return ESR_OK;
}

extern "C" int esr_pipeline_streams(const EsrPipeline* p, esr_stream_t* copy, esr_stream_t* side, esr_stream_t* main_) {
  ESR_REQUIRE(ok(p));
  if (copy) *copy = p->copy;
  if (side) *side = p->side;
  if (main_) *main_ = p->main;
  return ESR_OK;
}

extern "C" int esr_pipeline_set_buffers(EsrPipeline* p, void* const* ids_dev, void* const* counts_dev, size_t ids_bytes,
                                        size_t counts_bytes, const float* loss_src, float* loss_log, int64_t loss_len,
                                        float* loss_host) {
  ESR_REQUIRE(ok(p) && ids_dev && counts_dev && loss_log && loss_len > 0);
  for (int k = 0; k < p->depth; ++k) {
    ESR_REQUIRE(ids_dev[k] != nullptr && counts_dev[k] != nullptr);
    p->ids_dev[k] = ids_dev[k];
    p->counts_dev[k] = counts_dev[k];
  }
  p->ids_bytes = ids_bytes;
  p->counts_bytes = counts_bytes;
  p->loss_src = loss_src;
  p->loss_log = loss_log;
  p->loss_len = loss_len;
  p->loss_host = loss_host;
  return ESR_OK;
}

// which: 0 = the plan stage (side stream), 1 = the step stage (main stream).  Everything the calling thread launches on
// that stream until esr_pipeline_capture_end becomes the stage's graph for buffer parity k.
extern "C" int esr_pipeline_capture_begin(EsrPipeline* p, int32_t which) {
  ESR_REQUIRE(ok(p) && (which == 0 || which == 1) && p->capturing < 0);
  ESR_CUDA(cudaStreamBeginCapture(which == 0 ? p->side : p->main, cudaStreamCaptureModeThreadLocal));
  p->capturing = which;
  return ESR_OK;
}

extern "C" int esr_pipeline_capture_end(EsrPipeline* p, int32_t which, int32_t k) {
  ESR_REQUIRE(ok(p) && which == p->capturing && k >= 0 && k < p->depth);
  cudaGraph_t g = nullptr;
  p->capturing = -1;
  ESR_CUDA(cudaStreamEndCapture(which == 0 ? p->side : p->main, &g));
  cudaGraphExec_t* slot = which == 0 ? &p->g_plan[k] : &p->g_step[k];
  if (*slot) {
    cudaGraphExecDestroy(*slot);
    *slot = nullptr;
  }
  const cudaError_t e = cudaGraphInstantiate(slot, g, 0);
  cudaGraphDestroy(g);
  ESR_CUDA(e);
  // make the executable graph resident on the device now: a launch then has nothing left to fetch but its trigger
  ESR_CUDA(cudaGraphUpload(*slot, which == 0 ? p->side : p->main));
  return ESR_OK;
}

// One training step.  ids / counts: pinned host or device memory of ids_bytes / counts_bytes.  flags: bit 0 = also copy
// the step's loss to the pinned host log; bit 1 = the inputs were produced on caller_stream (NULL = the legacy default
// stream): stage them behind it; bit 2 = ids and counts are adjacent views of ONE allocation (uploaded with one copy);
// bit 3 = the inputs are PINNED HOST memory: staged by a small kernel over PCIe instead of the copy engine.
// Returns the step number in *step_out.
extern "C" int esr_pipeline_submit(EsrPipeline* p, const void* ids, const void* counts, esr_stream_t caller_stream,
                                   int32_t flags, int64_t* step_out) {
  ESR_RANGE("esr_pipeline_submit");
  const bool read_loss = (flags & 1) != 0;
  ESR_REQUIRE(ok(p) && ids && counts && p->capturing < 0);
  const int k = (int)(p->t % p->depth);
  ESR_REQUIRE(p->g_plan[k] != nullptr && p->g_step[k] != nullptr && p->ids_dev[k] != nullptr);
  if (flags & 2) {
    ESR_CUDA(cudaEventRecord(p->ev_in, static_cast<cudaStream_t>(caller_stream)));
    ESR_CUDA(cudaStreamWaitEvent(p->copy, p->ev_in, 0));
  }
  ESR_CUDA(cudaStreamWaitEvent(p->copy, p->ev_done[k], 0));  // staging / plan buffers k are free (step t - depth done)
  const int tr = p->trace_left > 0 ? p->trace_n : -1;
  if (tr >= 0) ESR_CUDA(cudaEventRecord(p->tr_ev[tr][0], p->copy));
  // one copy when ids and counts are adjacent on both sides (GloveTrainer.pinned_batch() hands out such batches): a
  // small host->device copy costs ~40 us of fixed latency on the copy engine whatever its size
  static const bool split_copy = getenv("ESR_PIPE_SPLIT_COPY") != nullptr;   // A/B probes
  static const bool d2h_main = getenv("ESR_PIPE_D2H_MAIN") != nullptr;
  // flags bit 2: the caller vouches that ids and counts live in ONE allocation (two adjacent allocations must not be
  // spanned by one copy: separately pinned blocks fail with cudaErrorInvalidValue)
  const bool adjacent = !split_copy && (flags & 4) != 0 &&
                        static_cast<const char*>(ids) + p->ids_bytes == static_cast<const char*>(counts) &&
                        static_cast<char*>(p->ids_dev[k]) + p->ids_bytes == static_cast<char*>(p->counts_dev[k]);
  static const bool use_dma = getenv("ESR_PIPE_DMA") != nullptr;   // A/B probe: copy-engine staging of host batches
  const bool by_kernel = (flags & 8) != 0 && !use_dma;              // bit 3: the source is pinned HOST memory
  int rc;
  if (adjacent) {
    rc = stage(p->ids_dev[k], ids, p->ids_bytes + p->counts_bytes, by_kernel, p->copy);
  } else {
    rc = stage(p->ids_dev[k], ids, p->ids_bytes, by_kernel, p->copy);
    if (rc == ESR_OK) rc = stage(p->counts_dev[k], counts, p->counts_bytes, by_kernel, p->copy);
  }
  if (rc != ESR_OK) return rc;
  if (tr >= 0) ESR_CUDA(cudaEventRecord(p->tr_ev[tr][1], p->copy));
  ESR_CUDA(cudaEventRecord(p->ev_copy[k], p->copy));
  ESR_CUDA(cudaStreamWaitEvent(p->side, p->ev_copy[k], 0));
  if (tr >= 0) ESR_CUDA(cudaEventRecord(p->tr_ev[tr][2], p->side));
  ESR_CUDA(cudaGraphLaunch(p->g_plan[k], p->side));
  if (tr >= 0) ESR_CUDA(cudaEventRecord(p->tr_ev[tr][3], p->side));
  ESR_CUDA(cudaEventRecord(p->ev_plan[k], p->side));
  ESR_CUDA(cudaStreamWaitEvent(p->main, p->ev_plan[k], 0));
  if (tr >= 0) ESR_CUDA(cudaEventRecord(p->tr_ev[tr][4], p->main));
  ESR_CUDA(cudaGraphLaunch(p->g_step[k], p->main));
  if (tr >= 0) {
    ESR_CUDA(cudaEventRecord(p->tr_ev[tr][5], p->main));
    p->trace_n += 1;
    p->trace_left -= 1;
  }
  // the step's finish kernel logs its loss to loss_log[t % loss_len] itself (EsrGloveCfg.loss_log: device-side slot), so
  // nothing but the graph sits on the main stream; the optional host read runs on its own stream behind ev_done
  if (read_loss && p->loss_host && d2h_main) {
    const int64_t slot = p->t % p->loss_len;
    ESR_CUDA(cudaMemcpyAsync(p->loss_host + slot, p->loss_log + slot, sizeof(float), cudaMemcpyDeviceToHost, p->main));
  }
  ESR_CUDA(cudaEventRecord(p->ev_done[k], p->main));
  if (read_loss && p->loss_host && !d2h_main) {
    const int64_t slot = p->t % p->loss_len;
    ESR_CUDA(cudaStreamWaitEvent(p->d2h, p->ev_done[k], 0));
    ESR_CUDA(cudaMemcpyAsync(p->loss_host + slot, p->loss_log + slot, sizeof(float), cudaMemcpyDeviceToHost, p->d2h));
  }
  if (step_out) *step_out = p->t;
  p->t += 1;
  return ESR_OK;
}

// Developer aid: record a timeline of the next n steps (n <= 64); esr_pipeline_trace_read returns, per traced step, the six
// stage boundaries {copy begin, copy end, plan begin, plan end, step begin, step end} in microseconds since the arming call.
extern "C" int esr_pipeline_trace(EsrPipeline* p, int32_t n) {
  ESR_REQUIRE(ok(p) && n >= 0 && n <= EsrPipeline::kTrace);
  if (p->tr_base == nullptr) {
    ESR_CUDA(cudaEventCreate(&p->tr_base));
    for (int i = 0; i < EsrPipeline::kTrace; ++i)
      for (int j = 0; j < 6; ++j) ESR_CUDA(cudaEventCreate(&p->tr_ev[i][j]));
  }
  ESR_CUDA(cudaEventRecord(p->tr_base, p->main));
  ESR_CUDA(cudaStreamWaitEvent(p->copy, p->tr_base, 0));
  p->trace_left = n;
  p->trace_n = 0;
  return ESR_OK;
}

extern "C" int esr_pipeline_trace_read(EsrPipeline* p, float* out_us /* [n][6] */, int32_t* n_out) {
  ESR_REQUIRE(ok(p) && out_us && n_out);
  ESR_CUDA(cudaDeviceSynchronize());
  for (int i = 0; i < p->trace_n; ++i)
    for (int j = 0; j < 6; ++j) {
      float ms = 0.f;
      ESR_CUDA(cudaEventElapsedTime(&ms, p->tr_base, p->tr_ev[i][j]));
      out_us[i * 6 + j] = ms * 1e3f;
    }
  *n_out = p->trace_n;
  return ESR_OK;
}

// Host-side wait until the staging copy of step `step` (already submitted) has completed: its host batch may be rewritten.
// Conservative: waits for the latest staging copy of that buffer parity.
extern "C" int esr_pipeline_wait_staged(const EsrPipeline* p, int64_t step) {
  ESR_REQUIRE(ok(p) && step >= 0 && step < p->t);
  ESR_CUDA(cudaEventSynchronize(p->ev_copy[step % p->depth]));
  return ESR_OK;
}

extern "C" int esr_pipeline_sync(const EsrPipeline* p) {
  ESR_REQUIRE(ok(p));
  ESR_CUDA(cudaStreamSynchronize(p->copy));
  ESR_CUDA(cudaStreamSynchronize(p->side));
  ESR_CUDA(cudaStreamSynchronize(p->main));
  ESR_CUDA(cudaStreamSynchronize(p->d2h));
  return ESR_OK;
}

extern "C" int esr_pipeline_destroy(EsrPipeline* p) {
  if (p == nullptr) return ESR_OK;
  ESR_REQUIRE(ok(p));
  cudaStreamSynchronize(p->copy);
  cudaStreamSynchronize(p->side);
  cudaStreamSynchronize(p->main);
  for (int k = 0; k < kMaxDepth; ++k) {
    if (p->g_plan[k]) cudaGraphExecDestroy(p->g_plan[k]);
    if (p->g_step[k]) cudaGraphExecDestroy(p->g_step[k]);
    cudaEventDestroy(p->ev_copy[k]);
    cudaEventDestroy(p->ev_plan[k]);
    cudaEventDestroy(p->ev_done[k]);
  }
  cudaEventDestroy(p->ev_in);
  cudaStreamDestroy(p->copy);
  cudaStreamDestroy(p->side);
  cudaStreamDestroy(p->main);
  cudaStreamDestroy(p->d2h);
  p->magic = 0;
  delete p;
  return ESR_OK;
}
