"""Build recipe of the CUDA library (sm_100a only, in-tree .so so that it travels to the GPU box)."""
import os
import subprocess
import sys

HERE = os.path.dirname(os.path.abspath(__file__))
CSRC = os.path.join(HERE, "csrc")
SOURCES = ["api.cu", "voxel.cu", "knn.cu", "loam.cu", "ndt.cu", "vgicp.cu", "scancontext.cu"]
HEADERS = ["common.cuh", "voxel.cuh", "loam.cuh", "ndt.cuh", "vgicp.cuh", "knn.cuh", "scancontext.cuh", "dev_linalg.cuh", "host_math.hpp", "ndt_logic.cuh", "vgicp_logic.cuh", "hostpack.hpp", "../../include/pcr_cuda.h"]
LIB = os.environ.get("PCR_LIB_OUT") or os.path.join(CSRC, "libpcr_cuda.so")  # PCR_LIB_OUT + PCR_NVCC_EXTRA: tuning variants
NVCC_FLAGS = ["-gencode", "arch=compute_100a,code=sm_100a", "-lineinfo", "-O3", "-std=c++17", "-Xcompiler", "-fPIC",
              "-Wno-deprecated-declarations", "-Xcompiler", "-Wno-deprecated-declarations"]


def _nvcc():
    for cand in (os.environ.get("NVCC"), "/usr/local/cuda/bin/nvcc", "nvcc"):
        if cand and (os.path.isabs(cand) and os.path.exists(cand) or not os.path.isabs(cand)):
            return cand
    return "nvcc"


def stale():
    if not os.path.exists(LIB):
        return True
    t = os.path.getmtime(LIB)
    return any(os.path.getmtime(os.path.join(CSRC, f)) > t for f in SOURCES + HEADERS)


def build(force=False, verbose=False):
    """Compile every CUDA source for sm_100a into csrc/libpcr_cuda.so (objects compiled in parallel)."""
    if not force and not stale():
        return LIB
    host_cxx = "/usr/bin/g++" if os.path.exists("/usr/bin/g++") else None
    objs, procs = [], []
    for src in SOURCES:
        obj = os.path.join(CSRC, src.replace(".cu", ".o")) if LIB.endswith("libpcr_cuda.so") else LIB + "." + src.replace(".cu", ".o")
        cmd = [_nvcc()] + NVCC_FLAGS + os.environ.get("PCR_NVCC_EXTRA", "").split() + (["-ccbin", host_cxx] if host_cxx else []) + ["-c", os.path.join(CSRC, src), "-o", obj]
        if verbose:
            cmd.insert(1, "-Xptxas=-v")
        procs.append((src, subprocess.Popen(cmd, stdout=subprocess.PIPE, stderr=subprocess.STDOUT, text=True)))
        objs.append(obj)
    failed = False
    for src, p in procs:
        out, _ = p.communicate()
        if p.returncode != 0:
            failed = True
            sys.stderr.write("nvcc failed on %s:\n%s\n" % (src, out))
        elif verbose:
            sys.stderr.write(out)
    if failed:
        raise RuntimeError("CUDA build failed")
    link = [_nvcc(), "-gencode", "arch=compute_100a,code=sm_100a", "-shared", "-o", LIB] + objs + (["-ccbin", host_cxx] if host_cxx else [])
    subprocess.check_call(link)
    return LIB


if __name__ == "__main__":
    print(build(force="--force" in sys.argv, verbose="-v" in sys.argv))
