"""Phase timestamps inside the nll+grad kernel (development build with -DFN_TRACE).

    nvcc ... -DFN_TRACE -DFN_SINGLE_TU -o tools/micro/libfn_trace.so fitnoise_b200/csrc/fn_api.cu
    python tools/micro/trace_eval.py [n] [m]
"""
import ctypes, os, sys, time
import numpy as np
ROOT = os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, ROOT)
from fitnoise_b200 import build
build.LIB_PATH = os.path.join(ROOT, "tools", "micro", "libfn_trace.so")
import fitnoise_b200 as fitnoise
from fitnoise_b200 import _native
n = int(sys.argv[1]) if len(sys.argv) > 1 else 1024
m = int(sys.argv[2]) if len(sys.argv) > 2 else 96
rng = np.random.default_rng(0)
y = rng.normal(size=(n, m))
design = np.array([[1.0, 0.0]] * (m // 2) + [[1.0, 1.0]] * (m - m // 2))
e = _native.Engine(0)
e.upload(y, None, fitnoise.Model_t()._family(m), None, design)
lib = _native.load_library()
lib.fn_trace_dump.argtypes = [ctypes.c_void_p, ctypes.c_void_p]
theta = np.array([8.0, 1.1])
names = {0: "kernel start (CTA 0)", 1: "TMA issued", 2: "tables done", 3: "tile loop done", 4: "finalize done",
         5: "block sum done", 8: "last CTA: ticket taken", 9: "last CTA: published"}
rows = []
for it in range(8):
    t0 = time.perf_counter_ns()
    e.nll_grad(theta)
    t1 = time.perf_counter_ns()
    buf = np.zeros(64, dtype=np.uint64)
    lib.fn_trace_dump(e._h, buf.ctypes.data)
    rows.append((buf.astype(np.int64), t1 - t0))
buf, wall = rows[-1]
base = buf[0]
print("n=%d m=%d: host call %.1f us" % (n, m, wall / 1000))
for k in sorted(names):
    print("  %-28s +%7.2f us" % (names[k], (buf[k] - base) / 1000.0))
ht = np.zeros(4)
lib.fn_host_trace.argtypes = [ctypes.c_void_p]
lib.fn_host_trace(ht.ctypes.data)
print("  host: stage_theta %.2f us, launch_eval %.2f us, wait_result %.2f us (C++ total %.2f us)" % (
    ht[1] - ht[0], ht[2] - ht[1], ht[3] - ht[2], ht[3] - ht[0]))
