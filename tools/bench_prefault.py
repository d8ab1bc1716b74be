"""How fast can a FRESH pageable 10 GB buffer (what a Haskell caller's mallocForeignPtrBytes hands elx_simulate) be made
resident?  First-touch faults by N threads, with and without MADV_HUGEPAGE, and MADV_POPULATE_WRITE in N threads.
Host only: run on the GPU box (its kernel / THP settings are what matters).  python tools/bench_prefault.py [GB] [threads]"""
import ctypes as C
import mmap
import sys
import threading
import time

import numpy as np

libc = C.CDLL("libc.so.6", use_errno=True)
libc.madvise.argtypes = [C.c_void_p, C.c_size_t, C.c_int]
MADV_HUGEPAGE, MADV_POPULATE_WRITE = 14, 23
gb = float(sys.argv[1]) if len(sys.argv) > 1 else 10.0
nt = int(sys.argv[2]) if len(sys.argv) > 2 else 13
n = int(gb * 1e9) // (1 << 21) * (1 << 21)
for f in ("enabled", "defrag", "shmem_enabled"):
    try:
        print("thp", f, open("/sys/kernel/mm/transparent_hugepage/" + f).read().strip())
    except OSError as e:
        print("thp", f, e)


def fresh():
    m = mmap.mmap(-1, n + (1 << 21), flags=mmap.MAP_PRIVATE | mmap.MAP_ANONYMOUS)
    a = np.frombuffer(m, dtype=np.uint8)
    base = a.ctypes.data
    off = (-base) % (1 << 21)
    return m, a[off:off + n]


def par(fn):
    ths = [threading.Thread(target=fn, args=(i,)) for i in range(nt)]
    t0 = time.perf_counter()
    [t.start() for t in ths]
    [t.join() for t in ths]
    return time.perf_counter() - t0


def touch(a):
    step = n // nt

    def w(i):
        a[i * step:(i + 1) * step:4096] = 1  # one store per 4 KB page (numpy releases the GIL in the strided fill)
    return par(w)


def populate(a):
    step = n // nt // 4096 * 4096
    base = a.ctypes.data
    rcs = []

    def w(i):
        rcs.append(libc.madvise(C.c_void_p(base + i * step), step, MADV_POPULATE_WRITE))
    dt = par(w)
    return dt, rcs


for label in ("touch", "hugepage+touch", "populate_write", "hugepage+populate_write"):
    m, a = fresh()
    if label.startswith("hugepage"):
        rc = libc.madvise(C.c_void_p(a.ctypes.data), n, MADV_HUGEPAGE)
        if rc:
            print(label, "madvise(MADV_HUGEPAGE) failed, errno", C.get_errno())
    if "populate" in label:
        dt, rcs = populate(a)
        print("%-26s %.1f ms  (%.1f GB/s) rcs=%s" % (label, dt * 1e3, n / dt / 1e9, set(rcs)))
    else:
        dt = touch(a)
        print("%-26s %.1f ms  (%.1f GB/s)" % (label, dt * 1e3, n / dt / 1e9))
    t0 = time.perf_counter()
    dt2 = touch(a)
    print("%-26s second pass %.1f ms" % ("", dt2 * 1e3))
    del a
    m.close()
