"""Sanitizer fuzz of the host record decoders (csrc/record_decode.cu is plain host C++): the file is compiled with
g++ -fsanitize=address,undefined next to tests/fuzz/decode_fuzz.cpp and fed valid and mutated CooccurrenceRow / TFRecord
streams through exact-size heap buffers.  CPU only."""
import ctypes as C
import os
import shutil
import struct
import subprocess

import numpy as np
import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


@pytest.mark.timeout(300)
def test_decoders_under_asan_ubsan(tmp_path):
    gxx = shutil.which("g++")
    if gxx is None:
        pytest.skip("g++ not on PATH")
    exe = str(tmp_path / "decode_fuzz")
    cmd = [gxx, "-std=c++17", "-O1", "-g", "-fsanitize=address,undefined", "-fno-sanitize-recover=undefined",
           "-I" + os.path.join(ROOT, "include"), "-x", "c++", os.path.join(ROOT, "esrecsys_b200", "csrc", "record_decode.cu"),
           os.path.join(ROOT, "tests", "fuzz", "decode_fuzz.cpp"), "-o", exe]
    r = subprocess.run(cmd, capture_output=True, text=True)
    if r.returncode != 0 and "asan" in (r.stderr + r.stdout).lower():
        pytest.skip("sanitizer runtime not installed")
    assert r.returncode == 0, r.stderr
    for seed in (1, 2):
        r = subprocess.run([exe, "4000", str(seed)], capture_output=True, text=True)
        assert r.returncode == 0, (r.stdout + r.stderr)[-3000:]
        assert "decode_fuzz ok" in r.stdout


def test_tfrecord_hostile_length_does_not_wrap():
    """A record header whose length makes 16 + len wrap around 2^64 is an incomplete record, not max_records empty ones."""
    from esrecsys_b200 import _lib
    h = _lib.lib()
    for hostile in (0xFFFFFFFFFFFFFFF0, 0xFFFFFFFFFFFFFFF8, 0xFFFFFFFFFFFFFFFF, 1 << 63):
        data = struct.pack("<Q", hostile) + b"\0" * 24
        buf = (C.c_char * len(data)).from_buffer_copy(data)
        vals = np.zeros(8, np.int64)
        offs = np.full(5, -1, np.int64)
        keys = (C.c_char_p * 1)(b"k")
        cvals = (C.c_void_p * 1)(vals.ctypes.data)
        coffs = (C.c_void_p * 1)(offs.ctypes.data)
        cap = (C.c_int64 * 1)(8)
        used = C.c_size_t(99)
        rec = h.esr_decode_tfrecord_int64(C.addressof(buf), len(data), 1, keys, cvals, cap, coffs, 4, C.byref(used))
        assert rec == 0 and used.value == 0, (hex(hostile), rec, used.value)
