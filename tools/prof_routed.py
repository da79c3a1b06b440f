"""Developer probe: device time per libesr entry point of one sharded GloVe step (OwnerRoutedGloveTrainer by default),
eager launches on ONE stream so every call can be bracketed with CUDA events.  Run under torchrun (any world size):

    python -m torch.distributed.run --nproc-per-node N --master-addr 127.0.0.1 --master-port 29650 tools/prof_routed.py
"""
import json
import os
import sys
from collections import OrderedDict

import numpy as np
import torch
import torch.distributed as dist

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from esrecsys_b200 import _lib as L, synth  # noqa: E402
from esrecsys_b200 import sharded  # noqa: E402


class TimedLib:
    """Proxy of the ctypes library: every esr_* call is bracketed with CUDA events on the current stream."""

    def __init__(self, real):
        self.real, self.acc, self.count, self.pending, self.on = real, OrderedDict(), {}, [], False

    def __getattr__(self, name):
        fn = getattr(self.real, name)
        if not name.startswith("esr_") or name.endswith("_bytes"):
            return fn

        def wrapped(*a):
            if not self.on:
                return fn(*a)
            e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            e0.record()
            rc = fn(*a)
            e1.record()
            self.pending.append((name, e0, e1))
            return rc
        return wrapped

    def flush(self):
        torch.cuda.synchronize()
        for name, e0, e1 in self.pending:
            self.acc[name] = self.acc.get(name, 0.0) + e0.elapsed_time(e1)
            self.count[name] = self.count.get(name, 0) + 1
        self.pending = []


def main():
    rank = int(os.environ["RANK"]); world = int(os.environ["WORLD_SIZE"]); local = int(os.environ["LOCAL_RANK"])
    torch.cuda.set_device(local)
    dist.init_process_group("nccl", device_id=torch.device("cuda", local))
    V, D, B = int(os.environ.get("V", 1000000)), 128, int(os.environ.get("B", 262144))
    cls = getattr(sharded, os.environ.get("TRAINER", "OwnerRoutedGloveTrainer"))
    tr = cls(V, D, B, graphs=False)
    tr.shard.rows0.normal_(0, 1 / np.sqrt(D))
    ids, counts = synth.glove_batches(V, B, 4, 17 * rank)
    d_ids = [torch.from_numpy(ids[k].reshape(-1)).cuda() for k in range(4)]
    d_cnt = [torch.from_numpy(counts[k]).cuda() for k in range(4)]
    proxy = TimedLib(L.lib())
    L.lib = lambda: proxy                      # every module resolves L.lib() at call time
    cur = torch.cuda.current_stream()
    tr.s_side = cur                            # one stream: phases run back to back
    for name in ("s_main", "s_ids", "s_gather"):
        if hasattr(tr, name):
            setattr(tr, name, cur)
    steps = 10
    tot = 0.0
    for it in range(steps + 3):
        proxy.on = it >= 3
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        tr.step(d_ids[it % 4], d_cnt[it % 4])
        e1.record()
        proxy.flush()
        if it >= 3:
            tot += e0.elapsed_time(e1)
    if rank == 0:
        us = {k: round(v / steps * 1e3, 1) for k, v in proxy.acc.items()}
        print(json.dumps(dict(trainer=cls.__name__, world=world, V=V, B=B, serial_step_us=round(tot / steps * 1e3, 1),
                              sum_us=round(sum(us.values()), 1), us=us, calls_per_step={k: v // steps for k, v in proxy.count.items()})))
    dist.barrier()
    dist.destroy_process_group()


if __name__ == "__main__":
    main()
