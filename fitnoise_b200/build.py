"""Build libfitnoise_b200.so (sm_100a) in-tree with nvcc.  Used by __graft_entry__.build().

The per-gene kernels are templates over the design width; each width is its own translation unit
(csrc/fn_width.cu with -DFN_W=<width>), compiled in parallel with csrc/fn_api.cu and linked into one
shared library.  `nvcc -DFN_SINGLE_TU ... fn_api.cu` builds the same library from one command.
"""
import os
import shutil
import subprocess
import sys
from concurrent.futures import ThreadPoolExecutor

HERE = os.path.dirname(os.path.abspath(__file__))
CSRC = os.path.join(HERE, "csrc")
LIB_DIR = os.path.join(HERE, "_lib")
LIB_PATH = os.path.join(LIB_DIR, "libfitnoise_b200.so")

OBJ_DIR = os.path.join(LIB_DIR, "obj")
WIDTHS = (1, 2, 3, 4, 6, 8, 12, 16)         # fn_host.hpp round_width()

NVCC_FLAGS = [
    "-O3", "-std=c++17",
    "-gencode", "arch=compute_100a,code=sm_100a",
    "-lineinfo",
    "-Xcompiler", "-fPIC",
    "-diag-suppress", "550",
]
LINK_FLAGS = ["-shared", "--cudart", "static", "-gencode", "arch=compute_100a,code=sm_100a"]


def sources():
    return sorted(os.path.join(CSRC, f) for f in os.listdir(CSRC) if f.endswith((".cu", ".cuh", ".hpp", ".h", ".c")))


HEADER = os.path.join(os.path.dirname(HERE), "include", "fitnoise_b200.h")
HASH_PATH = os.path.join(LIB_DIR, "source_hash.txt")


def source_hash():
    """sha1 over the contents of csrc/* and the public header: compiled into the library
    (fn_source_hash) so that a binary built from other sources is recognised whatever the mtimes."""
    import hashlib
    digest = hashlib.sha1()
    for path in sources() + [HEADER]:
        digest.update(os.path.basename(path).encode())
        with open(path, "rb") as f:
            digest.update(f.read())
    return digest.hexdigest()


def built_hash():
    try:
        with open(HASH_PATH) as f:
            return f.read().strip()
    except OSError:
        return None


def needs_build():
    return not os.path.exists(LIB_PATH) or built_hash() != source_hash()


def pycall_path():
    """The optional CPython entry for fn_nll_grad (csrc/fn_pycall.c), built for the running interpreter."""
    import sysconfig
    return os.path.join(LIB_DIR, "_fn_pycall" + (sysconfig.get_config_var("EXT_SUFFIX") or ".so"))


def build_pycall():
    """gcc build of csrc/fn_pycall.c; returns its path or None (no compiler / no Python.h: the ctypes
    path is used instead -- this module only marshals arguments, it computes nothing)."""
    import sysconfig
    include = sysconfig.get_paths().get("include")
    if not include or not os.path.exists(os.path.join(include, "Python.h")):
        return None
    out = pycall_path()
    cmd = [os.environ.get("CC", "gcc"), "-O2", "-shared", "-fPIC", "-I", include,
           os.path.join(CSRC, "fn_pycall.c"), "-o", out]
    try:
        proc = subprocess.run(cmd, capture_output=True, text=True)
    except OSError:
        return None
    return out if proc.returncode == 0 else None


def build(force=False, verbose=False):
    """Compile csrc/fn_api.cu (which includes every kernel header) into _lib/libfitnoise_b200.so."""
    if not force and not needs_build():
        if not os.path.exists(pycall_path()):
            build_pycall()
        return LIB_PATH
    nvcc = os.environ.get("NVCC", "nvcc")
    os.makedirs(OBJ_DIR, exist_ok=True)
    digest = source_hash()
    extra = ["-Xptxas", "-v"] if verbose else []
    jobs = []
    for w in sorted(WIDTHS, reverse=True):          # widest first: they take longest
        jobs.append([nvcc] + NVCC_FLAGS + extra + ["-DFN_W=%d" % w, "-c", os.path.join(CSRC, "fn_width.cu"),
                                                   "-o", os.path.join(OBJ_DIR, "fn_width_%d.o" % w)])

    def run(cmd):
        proc = subprocess.run(cmd, capture_output=True, text=True)
        if proc.returncode != 0:
            raise RuntimeError("nvcc failed:\n%s\n%s" % (" ".join(cmd), proc.stderr))
        return proc.stderr

    jobs.insert(1, [nvcc] + NVCC_FLAGS + extra + ["-DFN_SOURCE_HASH=\"%s\"" % digest, "-c", os.path.join(CSRC, "fn_api.cu"),
                    "-o", os.path.join(OBJ_DIR, "fn_api.o")])
    with ThreadPoolExecutor(max_workers=max(1, min(len(jobs), os.cpu_count() or 1))) as pool:
        logs = list(pool.map(run, jobs))
    objs = [c[-1] for c in jobs]
    log = run([nvcc] + LINK_FLAGS + ["-o", LIB_PATH] + objs)
    shutil.rmtree(OBJ_DIR, ignore_errors=True)     # only the .so travels to the GPU box
    build_pycall()
    with open(HASH_PATH, "w") as f:
        f.write(digest + "\n")
    if verbose:
        sys.stderr.write("".join(logs) + log)
    return LIB_PATH


if __name__ == "__main__":
    print(build(force="--force" in sys.argv, verbose="-v" in sys.argv))
