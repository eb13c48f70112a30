"""In-tree build of libcgpt.so (sm_100a only) with nvcc.

The built library lives at certifiedgpt_b200/lib/libcgpt.so: git-ignored, but it travels
to the GPU box with the gpurun snapshot.  nvcc cross-compiles without a GPU.
"""
import hashlib
import os
import subprocess
import sys

HERE = os.path.dirname(os.path.abspath(__file__))
CSRC = os.path.join(HERE, "csrc")
LIBDIR = os.path.join(HERE, "lib")
LIB = os.path.join(LIBDIR, "libcgpt.so")
STAMP = os.path.join(LIBDIR, "libcgpt.stamp")

# no --use_fast_math: kernels pick approximate intrinsics (ex2.approx, rcp.approx, MUFU sin/cos/lg2)
# explicitly where the error budget allows, everything else keeps IEEE behaviour
NVCC_FLAGS = [
    "-gencode", "arch=compute_100a,code=sm_100a", "-lineinfo", "-O3", "-std=c++17",
    "-Xcompiler", "-fPIC", "-Xcompiler", "-O2",
]


def sources():
    return sorted(os.path.join(CSRC, f) for f in os.listdir(CSRC) if f.endswith(".cu"))


def _digest():
    h = hashlib.sha256()
    for f in sorted(os.listdir(CSRC)):
        if f.endswith((".cu", ".cuh", ".h")):
            h.update(f.encode())
            h.update(open(os.path.join(CSRC, f), "rb").read())
    h.update(open(os.path.join(HERE, "..", "include", "cgpt.h"), "rb").read())
    h.update(" ".join(NVCC_FLAGS).encode())
    return h.hexdigest()


def is_current():
    return os.path.exists(LIB) and os.path.exists(STAMP) and open(STAMP).read().strip() == _digest()


def build(force=False, verbose=False):
    """Compile every .cu under csrc/ into one shared library."""
    os.makedirs(LIBDIR, exist_ok=True)
    if not force and is_current():
        return LIB
    nvcc = os.environ.get("NVCC", "nvcc")
    objs = []
    procs = []
    for src in sources():
        obj = os.path.join(LIBDIR, os.path.basename(src)[:-3] + ".o")
        cmd = [nvcc, *NVCC_FLAGS, "-c", src, "-o", obj]
        if verbose:
            cmd.insert(1, "-Xptxas")
            cmd.insert(2, "-v")
        procs.append((src, subprocess.Popen(cmd, stdout=subprocess.PIPE, stderr=subprocess.STDOUT)))
        objs.append(obj)
    failed = False
    for src, p in procs:
        out, _ = p.communicate()
        if p.returncode != 0 or verbose:
            sys.stderr.write(out.decode())
        if p.returncode != 0:
            failed = True
    if failed:
        raise RuntimeError("nvcc failed building libcgpt.so")
    subprocess.check_call([nvcc, "-shared", "-o", LIB, *objs,
                           "-gencode", "arch=compute_100a,code=sm_100a"])
    open(STAMP, "w").write(_digest())
    return LIB


if __name__ == "__main__":
    print(build(force="--force" in sys.argv, verbose="-v" in sys.argv))
