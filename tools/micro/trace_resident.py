"""Phase timestamps of one evaluation by the RESIDENT kernel (development build, tools/micro/build_trace.sh).

    python tools/micro/trace_resident.py [n] [m] [threads] [lpg]

GPU stamps are globaltimer ns of CTA 0 (command recognised, command seen by the CTA, tile loop done,
block sum done) and of the last CTA (ticket taken, totals published); host stamps are steady_clock.
The two clocks are not synchronised: only differences within one clock are meaningful, the host's
command-written -> result-seen interval minus the GPU's recognised -> published interval is the
PCIe share (command read + result write)."""
import ctypes, os, sys, time
import numpy as np
ROOT = os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, ROOT)
from fitnoise_b200 import build
build.LIB_PATH = os.path.join(ROOT, "tools", "micro", "libfn_trace.so")
build.needs_build = lambda: False
os.environ["FN_SKIP_SOURCE_HASH"] = "1"       # development build: whatever the sources were
import fitnoise_b200 as fitnoise
from fitnoise_b200 import _native
n = int(sys.argv[1]) if len(sys.argv) > 1 else 20000
m = int(sys.argv[2]) if len(sys.argv) > 2 else 6
threads = int(sys.argv[3]) if len(sys.argv) > 3 else 0
lpg = int(sys.argv[4]) if len(sys.argv) > 4 else 0
rng = np.random.default_rng(0)
y = rng.normal(size=(n, m))
design = np.array([[1.0, 0.0]] * (m // 2) + [[1.0, 1.0]] * (m - m // 2))
e = _native.Engine(0)
e.set_tuning(lpg=lpg, threads=threads)
e.upload(y, None, fitnoise.Model_t()._family(m), None, design)
lib = _native.load_library()
lib.fn_trace_dump.argtypes = [ctypes.c_void_p, ctypes.c_void_p]
lib.fn_host_trace.argtypes = [ctypes.c_void_p]
theta = np.array([8.0, 1.1])
names = {10: "CTA 0: command recognised", 11: "CTA 0: command in shared memory", 3: "CTA 0: tile loop done",
         4: "CTA 0: finalize done", 5: "CTA 0: block sum done", 8: "last CTA: ticket taken", 9: "last CTA: published"}
for mode in ("resident", "launch"):
    if mode == "resident":
        e.resident_begin()
    walls = []
    for it in range(300):
        t0 = time.perf_counter_ns()
        e.nll_grad(theta)
        walls.append(time.perf_counter_ns() - t0)
    ht = np.zeros(4)
    lib.fn_host_trace(ht.ctypes.data)
    buf = np.zeros(64, dtype=np.uint64)
    lib.fn_trace_dump(e._h, buf.ctypes.data)          # ends the resident phase
    buf = buf.astype(np.int64)
    print("%s n=%d m=%d: python call median %.2f us (min %.2f)" % (mode, n, m, np.median(walls) / 1000, min(walls) / 1000))
    base = buf[10] if mode == "resident" else buf[0]
    keys = sorted(names) if mode == "resident" else [0, 1, 2, 3, 4, 5, 8, 9]
    for k in keys:
        print("  gpu  %-34s +%7.2f us" % (names.get(k, "stamp %d" % k), (buf[k] - base) / 1000.0))
    print("  host: prepare %.2f us, send/launch %.2f us, wait %.2f us (C total %.2f us)" % (
        ht[1] - ht[0], ht[2] - ht[1], ht[3] - ht[2], ht[3] - ht[0]))
