"""Builds semantic_slam_mapping_b200/libssm.so in-tree with nvcc for sm_100a (cross-compiles without a GPU)."""
from __future__ import annotations

import os
import subprocess
import sys
from concurrent.futures import ThreadPoolExecutor

HERE = os.path.dirname(os.path.abspath(__file__))
CSRC = os.path.join(HERE, "csrc")
LIB = os.path.join(HERE, "libssm.so")
SOURCES = ["api.cu", "sgbm_cost.cu", "sgbm_aggregate.cu", "sgbm_vertical.cu", "sgbm_hsweep.cu", "sgbm_hsweep2.cu", "sgbm_select.cu", "mapper.cu", "voxel_table.cu", "comm.cu", "cues.cu", "labels.cu",
           "ingest.cu"]
NVCC_FLAGS = ["-gencode", "arch=compute_100a,code=sm_100a", "-O3", "-lineinfo", "-std=c++17", "-Xcompiler", "-fPIC",
              "-Xcompiler", "-Wall", "--fmad=false"]


def _stale() -> bool:
    if not os.path.exists(LIB):
        return True
    t = os.path.getmtime(LIB)
    deps = [os.path.join(CSRC, f) for f in os.listdir(CSRC)] + [os.path.join(HERE, "..", "include", "ssm.h"), __file__]
    return any(os.path.getmtime(d) > t for d in deps)


def build(force: bool = False, verbose: bool = False) -> str:
    if not (force or _stale()):
        return LIB
    nvcc = os.environ.get("NVCC", "/usr/local/cuda/bin/nvcc")
    objdir = os.path.join(HERE, "build")
    os.makedirs(objdir, exist_ok=True)

    def compile_one(src):
        obj = os.path.join(objdir, src.replace(".cu", ".o"))
        cmd = [nvcc, *NVCC_FLAGS, "-c", os.path.join(CSRC, src), "-o", obj]
        if verbose:
            cmd.insert(1, "-Xptxas=-v")
        r = subprocess.run(cmd, capture_output=True, text=True)
        if r.returncode != 0:
            raise RuntimeError(f"nvcc failed for {src}:\n{r.stdout}\n{r.stderr}")
        if verbose:
            sys.stderr.write(r.stderr)
        return obj

    with ThreadPoolExecutor(max_workers=6) as ex:
        objs = list(ex.map(compile_one, SOURCES))
    # (the arch flags at link time keep nvcc from adding an empty device-link cubin for its default architecture)
    r = subprocess.run([nvcc, "-gencode", "arch=compute_100a,code=sm_100a", "-shared", "-o", LIB, *objs, "-lcudart", "-ldl", "-lz"],
                       capture_output=True, text=True)
    if r.returncode != 0:
        raise RuntimeError(f"link failed:\n{r.stdout}\n{r.stderr}")
    return LIB



def build_host_driver(out: str | None = None) -> str:
    """g++ the C++ host layer (host/host.cpp) + tests/cpp/host_driver.cpp against libssm.so (no GPU needed to build)."""
    build()
    root = os.path.dirname(HERE)
    out = out or os.path.join(HERE, "build", "host_driver")
    os.makedirs(os.path.dirname(out), exist_ok=True)
    cuda_lib = os.environ.get("CUDA_LIB", "/usr/local/cuda/lib64")
    cmd = ["g++", "-std=c++17", "-O2", "-Wall", "-pthread", os.path.join(root, "tests", "cpp", "host_driver.cpp"),
           os.path.join(HERE, "host", "host.cpp"), "-L" + HERE, "-lssm", "-Wl,-rpath," + HERE, "-L" + cuda_lib,
           "-Wl,-rpath," + cuda_lib, "-o", out]
    r = subprocess.run(cmd, capture_output=True, text=True)
    if r.returncode != 0:
        raise RuntimeError(f"g++ failed:\n{r.stdout}\n{r.stderr}")
    return out


if __name__ == "__main__":
    print(build(force="--force" in sys.argv, verbose="-v" in sys.argv))
