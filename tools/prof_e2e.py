"""Developer probe: end-to-end (pinned host batches) vs device-resident step time of GloveTrainer, host time per submit."""
import json, os, sys, time
import numpy as np
import torch
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from esrecsys_b200 import engine, synth
from esrecsys_b200.trainer import GloveTrainer
V, D, B = 1_000_000, 128, 262144
steps = int(os.environ.get("STEPS", 100))
table = engine.EmbeddingTable(V, D)
table.rows0.normal_(0.0, 1.0 / np.sqrt(D))
ids, counts = synth.glove_batches(V, B, 8, 0)
tr = GloveTrainer(table, B)
dev = [(torch.from_numpy(ids[k].reshape(-1)).cuda(), torch.from_numpy(counts[k]).cuda()) for k in range(8)]
pin_adj, pin_sep = [], []
for k in range(8):
    hi, hc = tr.pinned_batch(); hi.copy_(torch.from_numpy(ids[k])); hc.copy_(torch.from_numpy(counts[k])); pin_adj.append((hi, hc))
    pin_sep.append((torch.from_numpy(ids[k]).pin_memory(), torch.from_numpy(counts[k]).pin_memory()))
def run(batches, read_loss):
    for k in range(5):
        tr.submit(*batches[k % 8])
    tr.synchronize(); torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    tr.s_side.wait_stream(tr.s_main); e0.record(tr.s_main); tr.s_side.wait_event(e0); tr.s_copy.wait_event(e0)
    t0 = time.perf_counter()
    for k in range(steps):
        tr.submit(*batches[k % 8], read_loss=read_loss)
    host = (time.perf_counter() - t0) / steps * 1e6
    e1.record(tr.s_main); tr.synchronize(); torch.cuda.synchronize()
    return round(e0.elapsed_time(e1) / steps * 1e3, 1), round(host, 1)
out = {"env": {k: v for k, v in os.environ.items() if k.startswith("ESR_PIPE")}, "steps": steps}
for name, b, rl in (("device", dev, False), ("pinned_adjacent", pin_adj, False), ("pinned_separate", pin_sep, False)):
    out[name] = dict(zip(("gpu_us_per_step", "host_us_per_submit"), run(b, rl)))
print(json.dumps(out))

# timeline of 12 steady-state steps, device vs pinned
import ctypes as C
from esrecsys_b200 import _lib as L
for name, b in ((("pinned_adjacent", pin_adj),) if os.environ.get("TIMELINE") else ()):
    for k in range(8):
        tr.submit(*b[k % 8])
    L.check(L.lib().esr_pipeline_trace(tr.pipe, 12), "trace")
    for k in range(12):
        tr.submit(*b[k % 8])
    buf = (C.c_float * (64 * 6))(); n = C.c_int32(0)
    L.check(L.lib().esr_pipeline_trace_read(tr.pipe, buf, C.byref(n)), "trace_read")
    print("timeline", name, "(us since arming: copy b/e, plan b/e, step b/e)")
    for i in range(n.value):
        print("  step %2d: " % i + "  ".join("%7.1f" % buf[i * 6 + j] for j in range(6)))
