"""In-tree build of the sm_100a CUDA libraries (nvcc cross-compiles without a GPU).

    python -m pegainfer_b200.build            # build everything that is stale
    python -m pegainfer_b200.build --force

Outputs (git-ignored, shipped to the GPU box by gpurun):
    pegainfer_b200/libpegainfer_kernels_b200.so   -- the C-ABI kernel library (include/pegainfer_kernels.h)
    pegainfer_b200/libpegainfer_qwen3_host.so     -- C++ mirror of the reference's Rust host side
"""
from __future__ import annotations

import os
import subprocess
import sys
from concurrent.futures import ThreadPoolExecutor

HERE = os.path.dirname(os.path.abspath(__file__))
CSRC = os.path.join(HERE, "csrc")
OBJ = os.path.join(HERE, "build")
NVCC = os.environ.get("NVCC", "/usr/local/cuda/bin/nvcc")
ARCH = ["-gencode", "arch=compute_100a,code=sm_100a"]
NVCC_FLAGS = ["-O3", "--std=c++17", "-lineinfo", "-Xcompiler", "-fPIC", "--expt-relaxed-constexpr",
              "-ccbin", "/usr/bin/g++"]

KERNEL_LIB = os.path.join(HERE, "libpegainfer_kernels_b200.so")
HOST_LIB = os.path.join(HERE, "libpegainfer_qwen3_host.so")


def _stale(target: str, deps: list[str]) -> bool:
    if not os.path.exists(target):
        return True
    t = os.path.getmtime(target)
    return any(os.path.getmtime(d) > t for d in deps if os.path.exists(d))


def _run(cmd: list[str]) -> None:
    r = subprocess.run(cmd, capture_output=True, text=True)
    if r.returncode != 0:
        sys.stderr.write(" ".join(cmd) + "\n" + r.stdout + r.stderr)
        raise RuntimeError(f"build failed: {cmd[0]} ... {cmd[-1]}")


def kernel_sources() -> list[str]:
    return sorted(os.path.join(CSRC, f) for f in os.listdir(CSRC) if f.endswith(".cu"))


def host_sources() -> list[str]:
    d = os.path.join(CSRC, "host")
    if not os.path.isdir(d):
        return []
    return sorted(os.path.join(d, f) for f in os.listdir(d) if f.endswith(".cpp"))


def build(force: bool = False, verbose: bool = False) -> None:
    os.makedirs(OBJ, exist_ok=True)
    headers = [os.path.join(CSRC, f) for f in os.listdir(CSRC) if f.endswith((".cuh", ".h"))]
    headers.append(os.path.join(HERE, "..", "include", "pegainfer_kernels.h"))
    hdir = os.path.join(CSRC, "host")
    if os.path.isdir(hdir):
        headers += [os.path.join(hdir, f) for f in os.listdir(hdir) if f.endswith((".hpp", ".h"))]

    jobs, objs = [], []
    for src in kernel_sources():
        obj = os.path.join(OBJ, os.path.basename(src)[:-3] + ".o")
        objs.append(obj)
        if force or _stale(obj, [src] + headers):
            jobs.append([NVCC, *ARCH, *NVCC_FLAGS, "-c", src, "-o", obj])
    if verbose and jobs:
        print(f"[pegainfer_b200.build] compiling {len(jobs)} CUDA file(s) for sm_100a", flush=True)
    if jobs:
        with ThreadPoolExecutor(max_workers=min(8, len(jobs))) as ex:
            list(ex.map(_run, jobs))
    if force or jobs or _stale(KERNEL_LIB, objs):
        _run([NVCC, *ARCH, "-shared", "-o", KERNEL_LIB, *objs, "-lcudart", "-ccbin", "/usr/bin/g++"])

    hsrcs = host_sources()
    if hsrcs and (force or _stale(HOST_LIB, hsrcs + headers)):
        if verbose:
            print("[pegainfer_b200.build] compiling the C++ host mirror", flush=True)
        _run(["/usr/bin/g++", "-O2", "-std=c++17", "-fPIC", "-shared", "-pthread",
              "-I/usr/local/cuda/include", "-I" + os.path.join(HERE, "..", "include"), *hsrcs,
              "-o", HOST_LIB, "-L/usr/local/cuda/lib64", "-lcudart", "-ldl",
              "-Wl,-rpath,/usr/local/cuda/lib64"])


if __name__ == "__main__":
    build(force="--force" in sys.argv, verbose=True)
    print("ok:", KERNEL_LIB)
