"""Host-only ceiling of the 2-bit transport: elx_unpack2_rows over a [rows][sites] result (no GPU involved).

  python tools/bench_unpack.py [rows] [sites]"""
import ctypes as C
import os
import sys
import time

import numpy as np

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from elynx_b200 import _lib as L  # noqa: E402


def main():
    rows = int(sys.argv[1]) if len(sys.argv) > 1 else 1000
    n = int(sys.argv[2]) if len(sys.argv) > 2 else 4_000_000
    packed = np.random.default_rng(0).integers(0, 256, size=(rows, n // 4), dtype=np.uint8)
    dst = np.zeros((rows, n), dtype=np.uint8)
    for th in (1, 2, 4, 8, 12, 14, 16):
        best = 1e9
        for _ in range(2):
            t0 = time.perf_counter()
            L.check(L.lib().elx_unpack2_rows(C.c_void_p(packed.ctypes.data), n // 4, C.c_void_p(dst.ctypes.data), n, rows, n, None, th))
            best = min(best, time.perf_counter() - t0)
        print("%2d threads: %.1f GB/s written" % (th, rows * n / best / 1e9), flush=True)
    t0 = time.perf_counter()
    dst2 = dst.copy()
    print("numpy copy (1 thread): %.1f GB/s" % (rows * n / (time.perf_counter() - t0) / 1e9))


if __name__ == "__main__":
    main()
