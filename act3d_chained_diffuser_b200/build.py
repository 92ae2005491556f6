"""Build libact3d_b200.so in-tree with nvcc for sm_100a (cross-compiles without a GPU)."""
import glob
import os
import shutil
import subprocess
import sys

HERE = os.path.dirname(os.path.abspath(__file__))
CSRC = os.path.join(HERE, "csrc")
LIB = os.path.join(HERE, "libact3d_b200.so")
NVCC_FLAGS = [
    "-gencode", "arch=compute_100a,code=sm_100a", "-O3", "-lineinfo", "-std=c++17",
    "-Xcompiler", "-fPIC", "-shared", "--expt-relaxed-constexpr",
]


def sources():
    return sorted(glob.glob(os.path.join(CSRC, "*.cu")))


def _stale():
    if not os.path.exists(LIB):
        return True
    t = os.path.getmtime(LIB)
    deps = sources() + glob.glob(os.path.join(CSRC, "*.cuh")) + \
        glob.glob(os.path.join(HERE, "..", "include", "*.h"))
    return any(os.path.getmtime(p) > t for p in deps)


def build(force=False, verbose=False):
    """Compile every .cu under csrc/ into one shared library. Raises on failure."""
    if not force and not _stale():
        return LIB
    nvcc = shutil.which("nvcc") or "/usr/local/cuda/bin/nvcc"
    if not os.path.exists(nvcc):
        raise RuntimeError("nvcc not found: libact3d_b200.so cannot be built (and there is no fallback path)")
    cmd = [nvcc] + NVCC_FLAGS + (["-Xptxas", "-v"] if verbose else []) + ["-o", LIB] + sources()
    res = subprocess.run(cmd, capture_output=True, text=True)
    if verbose:
        sys.stderr.write(res.stderr)
    if res.returncode != 0:
        raise RuntimeError("nvcc failed:\n" + res.stdout + res.stderr)
    return LIB


def build_tools(verbose=False):
    """Compile the standalone micro-benchmarks under tools/micro/ (one executable each, into build/)."""
    nvcc = shutil.which("nvcc") or "/usr/local/cuda/bin/nvcc"
    root = os.path.join(HERE, "..")
    out_dir = os.path.join(root, "build")
    os.makedirs(out_dir, exist_ok=True)
    built = []
    for src in sorted(glob.glob(os.path.join(root, "tools", "micro", "*.cu"))):
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
