"""Builds libplinopt_b200.so (hand-written sm_100a kernels + the extern "C" ABI of
include/plinopt_b200.h) in-tree with nvcc.  No torch dependency: the library is
a plain C-ABI shared object (cudart linked statically)."""
import os
import shutil
import subprocess
import sys
from concurrent.futures import ThreadPoolExecutor

HERE = os.path.dirname(os.path.abspath(__file__))
CSRC = os.path.join(HERE, "csrc")
BUILD = os.path.join(HERE, "build")
LIB = os.path.join(HERE, "libplinopt_b200.so")
SOURCES = ["common.cu", "orbit_sweep.cu", "lincomb_search.cu", "lincomb_quad.cu", "mmcheck.cu", "factor_sweep.cu", "dependency_explore.cu", "multi_device.cu", "nccl_comm.cu", "peaks.cu", "host/host_api.cpp"]
NVCC_FLAGS = ["-gencode", "arch=compute_100a,code=sm_100a", "-lineinfo", "-O3", "-std=c++17",
              "-Xcompiler", "-fPIC", "--fmad=true", "-Xptxas", "-v",
              "-split-compile", "0"]  # ptxas of the many template instantiations in parallel (orbit_sweep.cu: 180 s -> 60 s)


def _nvcc():
    nv = shutil.which("nvcc") or "/usr/local/cuda/bin/nvcc"
    if not os.path.exists(nv):
        raise RuntimeError("nvcc not found")
    return nv


HOST_HEADERS = ["exact.hpp", "sparsify_host.hpp", "matrix_io.hpp", "slp.hpp", "factor_host.hpp", "dependency_host.hpp", "negate_host.hpp"]
# headers each source really includes (a change to the host layer does not recompile the 2000-line orbit kernels)
EXTRA_DEPS = {
    "lincomb_search.cu": ["lincomb_common.cuh", "host/exact.hpp"],
    "lincomb_quad.cu": ["lincomb_common.cuh", "host/exact.hpp"],
    "factor_sweep.cu": ["host/exact.hpp"],
    "dependency_explore.cu": ["host/exact.hpp"],
    "host/host_api.cpp": ["host/" + h for h in HOST_HEADERS],
}


def _deps(src):
    deps = [os.path.join(CSRC, src), os.path.join(CSRC, "plo_device.cuh"), os.path.join(os.path.dirname(HERE), "include", "plinopt_b200.h"), os.path.abspath(__file__)]
    deps += [os.path.join(CSRC, h) for h in EXTRA_DEPS.get(src, [])]
    return max(os.path.getmtime(d) for d in deps if os.path.exists(d))


def build_library(force=False, verbose=False):
    os.makedirs(BUILD, exist_ok=True)
    nvcc = _nvcc()
    host_cxx = "/usr/bin/g++" if os.path.exists("/usr/bin/g++") else None
    ccbin = ["-ccbin", host_cxx] if host_cxx else []
    objs, jobs = [], []
    for s in SOURCES:
        o = os.path.join(BUILD, os.path.basename(s).replace(".cu", ".o").replace(".cpp", ".o"))
        objs.append(o)
        if force or not os.path.exists(o) or os.path.getmtime(o) < _deps(s):
            jobs.append((s, o))

    def compile_one(job):
        s, o = job
        extra = os.environ.get("PLO_NVCC_EXTRA", "").split()  # tuning aid, e.g. -DPLO_MM_WARPS=24
        cmd = [nvcc] + ccbin + NVCC_FLAGS + extra + (["-x", "cu"] if s.endswith(".cpp") else []) + ["-c", os.path.join(CSRC, s), "-o", o]
        p = subprocess.run(cmd, capture_output=True, text=True)
        with open(o + ".log", "w") as f:
            f.write(" ".join(cmd) + "\n" + p.stdout + p.stderr)
        if p.returncode != 0:
            raise RuntimeError(f"nvcc failed on {s}:\n{p.stdout}\n{p.stderr}")
        return s

    if jobs:
        with ThreadPoolExecutor(max_workers=min(len(jobs), os.cpu_count() or 4)) as ex:
            for s in ex.map(compile_one, jobs):
                if verbose:
                    print("compiled", s, file=sys.stderr)
    if jobs or force or not os.path.exists(LIB):
        cmd = [nvcc] + ccbin + ["-shared", "-cudart", "static", "-o", LIB] + objs + ["-ldl"]
        p = subprocess.run(cmd, capture_output=True, text=True)
        if p.returncode != 0:
            raise RuntimeError(f"link failed:\n{p.stdout}\n{p.stderr}")
    return LIB


CLI_DIR = os.path.join(HERE, "cli")
BIN_DIR = os.path.join(os.path.dirname(HERE), "bin")
CLIS = ["sparsifier", "orbiter", "MMchecker", "factorizer", "dependency", "negater", "rotater", "growthfactor"]


def build_clis(force=False):
    """bin/sparsifier, bin/orbiter, bin/MMchecker: host C++ drivers linked against the library."""
    os.makedirs(BIN_DIR, exist_ok=True)
    cxx = "/usr/bin/g++" if os.path.exists("/usr/bin/g++") else "g++"
    out = []
    for name in CLIS:
        src = os.path.join(CLI_DIR, name + ".cpp")
        exe = os.path.join(BIN_DIR, name)
        deps = [src, os.path.join(CLI_DIR, "cli_common.hpp"), os.path.join(CSRC, "host", "matrix_io.hpp"), os.path.join(CSRC, "host", "exact.hpp"), LIB]
        if force or not os.path.exists(exe) or os.path.getmtime(exe) < max(os.path.getmtime(d) for d in deps):
            cmd = [cxx, "-O2", "-std=c++17", "-o", exe, src, "-L" + HERE, "-lplinopt_b200", "-Wl,-rpath,$ORIGIN/../plinopt_b200"]
            p = subprocess.run(cmd, capture_output=True, text=True)
            if p.returncode != 0:
                raise RuntimeError(f"g++ failed on {name}:\n{p.stdout}\n{p.stderr}")
        out.append(exe)
    return out


if __name__ == "__main__":
    print(build_library(force="--force" in sys.argv, verbose=True))
    print(build_clis(force="--force" in sys.argv))
