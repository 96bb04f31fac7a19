"""Builds lib/libcdb200.so from csrc/*.cu with nvcc for sm_100a (cross-compiles without a GPU).

Every translation unit is compiled to its own object file (in parallel, re-used while the source and the
headers are older than it) and the objects are linked into one shared library."""
import os
import subprocess
import sys
from concurrent.futures import ThreadPoolExecutor

HERE = os.path.dirname(os.path.abspath(__file__))
CSRC = os.path.join(HERE, "csrc")
LIBDIR = os.path.join(HERE, "lib")
OBJDIR = os.path.join(HERE, "build")
LIB = os.path.join(LIBDIR, "libcdb200.so")

NVCC_FLAGS = [
    "-gencode", "arch=compute_100a,code=sm_100a", "-lineinfo", "-O3", "-std=c++17",
    "-Xcompiler", "-fPIC",
]


# experiments: extra nvcc flags (e.g. CDB_EXTRA_NVCC_FLAGS="-DCDB_WAIT_SPIN=1"); objects are rebuilt when it changes
EXTRA = os.environ.get("CDB_EXTRA_NVCC_FLAGS", "").split()


def sources():
    return sorted(os.path.join(CSRC, f) for f in os.listdir(CSRC) if f.endswith(".cu"))


def _headers():
    hs = [os.path.join(CSRC, f) for f in os.listdir(CSRC) if f.endswith((".cuh", ".h"))]
    hs.append(os.path.join(HERE, "..", "include", "cdb200.h"))
    return hs


def needs_build():
    if not os.path.exists(LIB):
        return True
    t = os.path.getmtime(LIB)
    deps = [os.path.join(CSRC, f) for f in os.listdir(CSRC)]
    deps.append(os.path.join(HERE, "..", "include", "cdb200.h"))
    return any(os.path.getmtime(d) > t for d in deps)


def _compile(src, force, verbose):
    obj = os.path.join(OBJDIR, os.path.basename(src)[:-3] + ".o")
    newest = max([os.path.getmtime(src)] + [os.path.getmtime(h) for h in _headers()])
    if not force and os.path.exists(obj) and os.path.getmtime(obj) > newest:
        return obj, ""
    nvcc = os.environ.get("NVCC", "nvcc")
    cmd = [nvcc] + NVCC_FLAGS + EXTRA + (["-Xptxas", "-v"] if verbose else []) + ["-c", "-o", obj, src]
    proc = subprocess.run(cmd, capture_output=True, text=True)
    if proc.returncode != 0:
        raise RuntimeError("nvcc failed on %s:\n%s%s" % (src, proc.stdout, proc.stderr))
    return obj, proc.stderr


def build(force=False, verbose=False):
    if not force and not needs_build():
        return LIB
    os.makedirs(LIBDIR, exist_ok=True)
    os.makedirs(OBJDIR, exist_ok=True)
    with ThreadPoolExecutor(max_workers=max(1, min(8, os.cpu_count() or 1))) as pool:
        results = list(pool.map(lambda s: _compile(s, force, verbose), sources()))
    nvcc = os.environ.get("NVCC", "nvcc")
    cmd = [nvcc, "-shared", "-gencode", "arch=compute_100a,code=sm_100a", "-o", LIB] + [o for o, _ in results]
    proc = subprocess.run(cmd, capture_output=True, text=True)
    if proc.returncode != 0:
        raise RuntimeError("nvcc link failed:\n" + proc.stdout + proc.stderr)
    if verbose:
        for _, log in results:
            print(log)
    return LIB


if __name__ == "__main__":
    print(build(force="--force" in sys.argv, verbose="-v" in sys.argv))
