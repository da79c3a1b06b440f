"""Host-side throughput of the input pipeline (no GPU): pairs/s of (1) the reference-format reader -- bz2 + base64 + protobuf,
W decoder threads, file order preserved --, (2) the decoded-corpus cache (memory-mapped flat arrays) and (3) batch assembly
into a ring of (2,B) int32 + (B,) f32 blocks (pinned tensors in the product, plain arrays here), next to the ~1.7 G pairs/s
the CUDA step consumes.   python tools/bench_loader.py [--triples 20000000]"""
import argparse
import base64
import bz2
import json
import os
import sys
import tempfile
import time

import numpy as np

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from esrecsys_b200.wikipedia import cooccurrence_matrix as cm  # noqa: E402
from esrecsys_b200.wikipedia import input_pipeline as ip  # noqa: E402


def write_corpus(d, triples, parts, seed=0):
    rng = np.random.default_rng(seed)
    per_row = 500                                   # make_cooccurrence.py caps a row message at 1001 entries
    rows = triples // per_row // parts
    for p in range(parts):
        lines = []
        for _ in range(rows):
            idx = int(rng.integers(2000, 1_000_000))
            lines.append(base64.b64encode(cm.encode_row(idx, rng.integers(1, idx, per_row).tolist(),
                                                        (rng.random(per_row) * 5).astype(np.float32).tolist())) + b"\n")
        with bz2.open(os.path.join(d, "part-%05d.bz2" % p), "wb") as f:
            f.write(b"".join(lines))
    return rows * per_row * parts


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--triples", type=int, default=8_000_000)
    ap.add_argument("--parts", type=int, default=8)
    ap.add_argument("--batch", type=int, default=262144)
    a = ap.parse_args()
    cores = len(os.sched_getaffinity(0))
    out = {"cores": cores, "batch": a.batch}
    with tempfile.TemporaryDirectory() as d:
        n = write_corpus(d, a.triples, a.parts)
        out["triples"] = n
        pat = os.path.join(d, "part-*.bz2")
        for w in (1, min(4, cores), cores):
            t0 = time.perf_counter()
            rd = ip.ParallelPartReader(pat, workers=w)
            got = sum(i.size for i, _, _ in rd)
            dt = time.perf_counter() - t0
            assert got == n
            out["bz2_b64_protobuf_reader_%d_threads_Mpairs_s" % w] = round(n / dt / 1e6, 2)
        t0 = time.perf_counter()
        cache = ip.TripleCache.build(pat, os.path.join(d, "cache"), workers=cores)
        out["cache_build_Mpairs_s"] = round(n / (time.perf_counter() - t0) / 1e6, 2)
        B = a.batch

        def make_block():
            buf = np.empty(3 * B, np.int32)
            return buf[:2 * B].reshape(2, B), buf[2 * B:].view(np.float32)
        for shuffle in (0, 5_000_000 if n >= 6_000_000 else n // 2):
            t0 = time.perf_counter()
            ld = ip.PinnedBatchLoader(cache.blocks(loop=True), B, make_block, ring=4, shuffle_size=shuffle,
                                      rng=np.random.default_rng(0))
            steps = 0
            for _ in ld:
                steps += 1
                if time.perf_counter() - t0 > 3.0:
                    break
            ld.close()
            dt = time.perf_counter() - t0
            out["cache_to_batch_ring_shuffle%d_Mpairs_s" % shuffle] = round(steps * B / dt / 1e6, 1)
    print(json.dumps(out))


if __name__ == "__main__":
    main()
