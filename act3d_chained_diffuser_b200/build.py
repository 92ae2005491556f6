"""Build libact3d_b200.so in-tree with nvcc for sm_100a (cross-compiles without a GPU).

Each csrc/*.cu is compiled to an object under build/obj/ (in parallel, only when stale against its own source, the
shared headers and include/*.h) and the objects are linked into one shared library."""
import concurrent.futures
import glob
import os
import shutil
import subprocess
import sys

HERE = os.path.dirname(os.path.abspath(__file__))
ROOT = os.path.abspath(os.path.join(HERE, ".."))
CSRC = os.path.join(HERE, "csrc")
OBJ = os.path.join(ROOT, "build", "obj")
LIB = os.path.join(HERE, "libact3d_b200.so")
NVCC_FLAGS = [
    "-gencode", "arch=compute_100a,code=sm_100a", "-O3", "-lineinfo", "-std=c++17",
    "-Xcompiler", "-fPIC", "--expt-relaxed-constexpr",
]


def sources():
    return sorted(glob.glob(os.path.join(CSRC, "*.cu")))


def _headers():
    return glob.glob(os.path.join(CSRC, "*.cuh")) + glob.glob(os.path.join(ROOT, "include", "*.h"))


def _nvcc():
    nvcc = shutil.which("nvcc") or "/usr/local/cuda/bin/nvcc"
    if not os.path.exists(nvcc):
        raise RuntimeError("nvcc not found: libact3d_b200.so cannot be built (and there is no fallback path)")
    return nvcc


def _obj_of(src):
    return os.path.join(OBJ, os.path.splitext(os.path.basename(src))[0] + ".o")


def _compile(nvcc, src, verbose):
    extra = os.environ.get("A3D_NVCC_EXTRA", "").split()          # study builds (e.g. -DA3D_X6_TRACE)
    cmd = [nvcc] + NVCC_FLAGS + extra + (["-Xptxas", "-v"] if verbose else []) + ["-c", "-o", _obj_of(src), src]
    res = subprocess.run(cmd, capture_output=True, text=True)
    return src, res


def build(force=False, verbose=False):
    """Compile every .cu under csrc/ into one shared library. Raises on failure."""
    nvcc = _nvcc()
    os.makedirs(OBJ, exist_ok=True)
    hdr_t = max([os.path.getmtime(p) for p in _headers()] + [os.path.getmtime(__file__)])
    todo = []
    for src in sources():
        o = _obj_of(src)
        if force or not os.path.exists(o) or os.path.getmtime(o) < max(os.path.getmtime(src), hdr_t):
            todo.append(src)
    for o in glob.glob(os.path.join(OBJ, "*.o")):          # objects of deleted sources
        if o not in [_obj_of(s) for s in sources()]:
            os.remove(o)
            force = True
    if todo:
        with concurrent.futures.ThreadPoolExecutor(max_workers=min(len(todo), os.cpu_count() or 4)) as ex:
            for src, res in ex.map(lambda s: _compile(nvcc, s, verbose), todo):
                if verbose:
                    sys.stderr.write(res.stderr)
                if res.returncode != 0:
                    raise RuntimeError("nvcc failed on " + src + ":\n" + res.stdout + res.stderr)
    objs = [_obj_of(s) for s in sources()]
    if force or todo or not os.path.exists(LIB) or any(os.path.getmtime(o) > os.path.getmtime(LIB) for o in objs):
        res = subprocess.run([nvcc, "-shared", "-gencode", "arch=compute_100a,code=sm_100a", "-o", LIB] + objs,
                             capture_output=True, text=True)
        if res.returncode != 0:
            raise RuntimeError("link failed:\n" + res.stdout + res.stderr)
    return LIB


def build_tools(verbose=False):
    """Compile the standalone micro-benchmarks under tools/micro/ (one executable each, into build/)."""
    nvcc = _nvcc()
    out_dir = os.path.join(ROOT, "build")
    os.makedirs(out_dir, exist_ok=True)
    built = []
    for src in sorted(glob.glob(os.path.join(ROOT, "tools", "micro", "*.cu"))):
        exe = os.path.join(out_dir, os.path.splitext(os.path.basename(src))[0])
        if os.path.exists(exe) and os.path.getmtime(exe) >= os.path.getmtime(src):
            built.append(exe)
            continue
        res = subprocess.run([nvcc, "-gencode", "arch=compute_100a,code=sm_100a", "-O3", "-lineinfo", "-o", exe, src],
                             capture_output=True, text=True)
        if res.returncode != 0:
            raise RuntimeError("nvcc failed on " + src + ":\n" + res.stdout + res.stderr)
        built.append(exe)
    return built


if __name__ == "__main__":
    print(build(force="--force" in sys.argv, verbose="-v" in sys.argv))
