"""Builds ``csrc/libbhmc.so`` in-tree with nvcc for sm_100a (B200).  No torch dependency:
the library is a plain C-ABI shared object (include/bhmc.h) loaded through ctypes."""
import os
import shutil
import subprocess
import sys
from concurrent.futures import ThreadPoolExecutor

HERE = os.path.dirname(os.path.abspath(__file__))
CSRC = os.path.join(HERE, "csrc")
LIB = os.path.join(CSRC, "libbhmc.so")
SOURCES = ["capi.cu", "update.cu", "softmax_simt.cu", "softmax_tc.cu", "mlp.cu", "tc_bgemm.cu", "comm.cu"]
NVCC_FLAGS = ["-gencode", "arch=compute_100a,code=sm_100a", "-O3", "-std=c++17", "-lineinfo",
              "-Xcompiler", "-fPIC", "--expt-relaxed-constexpr", "-Xptxas", "-v"]


def _nvcc():
    exe = shutil.which("nvcc") or "/usr/local/cuda/bin/nvcc"
    if not os.path.exists(exe):
        raise RuntimeError("nvcc not found: libbhmc.so cannot be built")
    return exe


def _stale(target, deps):
    if not os.path.exists(target):
        return True
    t = os.path.getmtime(target)
    return any(os.path.getmtime(d) > t for d in deps)


def build(force=False, verbose=False):
    """Compile every CUDA translation unit for sm_100a and link libbhmc.so. Returns its path."""
    nvcc = _nvcc()
    srcs = [s for s in SOURCES if os.path.exists(os.path.join(CSRC, s))]
    headers = [os.path.join(CSRC, f) for f in os.listdir(CSRC) if f.endswith(".cuh")]
    headers.append(os.path.join(os.path.dirname(HERE), "include", "bhmc.h"))
    objs = []
    jobs = []
    for s in srcs:
        src = os.path.join(CSRC, s)
        obj = os.path.join(CSRC, s[:-3] + ".o")
        objs.append(obj)
        if force or _stale(obj, [src] + headers):
            jobs.append((src, obj))

    def compile_one(job):
        src, obj = job
        r = subprocess.run([nvcc] + NVCC_FLAGS + ["-c", src, "-o", obj], capture_output=True, text=True)
        with open(obj[:-2] + ".ptxas.log", "w") as f:
            # tracked next to the sources (registers / spills per kernel); compile times would change with every build
            f.write("".join(l for l in r.stderr.splitlines(True) if "Compile time" not in l))
        if r.returncode != 0:
            raise RuntimeError("nvcc failed for %s:\n%s" % (src, r.stderr[-4000:]))
        if verbose:
            print(r.stderr)
        return obj

    with ThreadPoolExecutor(max_workers=max(1, min(len(jobs), os.cpu_count() or 1))) as ex:
        list(ex.map(compile_one, jobs))
    if jobs or not os.path.exists(LIB):
        r = subprocess.run([nvcc, "-shared", "-o", LIB] + objs + ["-gencode", "arch=compute_100a,code=sm_100a",
                                                                   "-cudart", "static", "-ldl"],
                           capture_output=True, text=True)
        if r.returncode != 0:
            raise RuntimeError("link failed:\n" + r.stderr[-4000:])
    return LIB


if __name__ == "__main__":
    print(build(force="--force" in sys.argv, verbose="-v" in sys.argv))
