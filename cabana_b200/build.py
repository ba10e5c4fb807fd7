"""Build the C-ABI shared library (hand-written CUDA, sm_100a) in-tree with nvcc.

    python -m cabana_b200.build [--force]

Output: cabana_b200/lib/libcabana_b200.so (git-ignored; travels to the GPU box).
"""
from __future__ import annotations

import hashlib
import os
import subprocess
import sys

HERE = os.path.dirname(os.path.abspath(__file__))
ROOT = os.path.dirname(HERE)
CSRC = os.path.join(HERE, "csrc")
LIB_DIR = os.path.join(HERE, "lib")
LIB_PATH = os.path.join(LIB_DIR, "libcabana_b200.so")
STAMP = os.path.join(LIB_DIR, "build.stamp")

SOURCES = ["cb_core.cu", "cb_scan.cu", "cb_lcl.cu", "cb_verlet.cu", "cb_verlet_fine.cu", "cb_verlet_tile.cu", "cb_traverse.cu", "cb_comm.cu"]
HEADERS = ["cb_common.cuh", "cb_internal.h", "cb_verlet_fine.h", "cb_verlet_tile.h", os.path.join(ROOT, "include", "cabana_b200.h")]

NVCC = os.environ.get("CB_NVCC", "/usr/local/cuda/bin/nvcc")
HOST_CXX = "/usr/bin/g++"  # the image's $CXX wrapper lacks libgomp.spec; use the system g++

NVCC_FLAGS = [
    "-gencode", "arch=compute_100a,code=sm_100a",
    "-O3", "-std=c++17", "-lineinfo",
    "-Xcompiler", "-fPIC,-ffp-contract=off,-Wall,-Wno-unused-function",
    "-ccbin", HOST_CXX,
    "--expt-relaxed-constexpr",
    "-Xptxas", "-v",
] + os.environ.get("CB_NVCC_EXTRA", "").split()   # e.g. -DCB_ABLATE_EMIT for profiling ablations


def _digest() -> str:
    h = hashlib.sha256()
    for name in SOURCES + HEADERS:
        path = name if os.path.isabs(name) else os.path.join(CSRC, name)
        with open(path, "rb") as f:
            h.update(f.read())
    h.update(" ".join(NVCC_FLAGS).encode())
    return h.hexdigest()


def build(force: bool = False, verbose: bool = False) -> str:
    os.makedirs(LIB_DIR, exist_ok=True)
    digest = _digest()
    if not force and os.path.exists(LIB_PATH) and os.path.exists(STAMP):
        with open(STAMP) as f:
            if f.read().strip() == digest:
                return LIB_PATH
    if not os.path.exists(NVCC):
        raise RuntimeError(f"nvcc not found at {NVCC}; cannot build libcabana_b200.so")
    objs = []
    procs = []
    for src in SOURCES:
        obj = os.path.join(LIB_DIR, src.replace(".cu", ".o"))
        objs.append(obj)
        cmd = [NVCC, *NVCC_FLAGS, "-c", os.path.join(CSRC, src), "-o", obj]
        procs.append((src, subprocess.Popen(cmd, stdout=subprocess.PIPE, stderr=subprocess.STDOUT, text=True)))
    log = []
    failed = False
    for src, p in procs:
        out, _ = p.communicate()
        log.append(f"==== {src}\n{out}")
        if p.returncode != 0:
            failed = True
    with open(os.path.join(LIB_DIR, "build.log"), "w") as f:
        f.write("\n".join(log))
    if failed:
        sys.stderr.write("\n".join(log))
        raise RuntimeError("nvcc failed; see cabana_b200/lib/build.log")
    link = [NVCC, "-shared", "-ccbin", HOST_CXX, "-o", LIB_PATH, *objs, "-cudart", "shared", "-ldl"]
    r = subprocess.run(link, capture_output=True, text=True)
    if r.returncode != 0:
        sys.stderr.write(r.stdout + r.stderr)
        raise RuntimeError("link failed")
    with open(STAMP, "w") as f:
        f.write(digest)
    if verbose:
        sys.stdout.write("\n".join(log))
    return LIB_PATH


CPP_TEST_SRC = os.path.join(ROOT, "tests", "cpp", "test_cabana_api.cu")
CPP_TEST_BIN = os.path.join(ROOT, "tests", "cpp", "test_cabana_api.bin")


def build_cpp_test(force: bool = False) -> str:
    """Compile the C++ test of include/Cabana_B200.hpp against the in-tree library."""
    build()
    deps = [CPP_TEST_SRC, os.path.join(ROOT, "include", "Cabana_B200.hpp"),
            os.path.join(ROOT, "include", "cabana_b200.h"), LIB_PATH]
    if (not force and os.path.exists(CPP_TEST_BIN)
            and all(os.path.getmtime(CPP_TEST_BIN) >= os.path.getmtime(d) for d in deps)):
        return CPP_TEST_BIN
    cmd = [NVCC, "-gencode", "arch=compute_100a,code=sm_100a", "-O2", "-std=c++17",
           "-ccbin", HOST_CXX, "--extended-lambda", "-I", os.path.join(ROOT, "include"),
           CPP_TEST_SRC, "-o", CPP_TEST_BIN, "-L", LIB_DIR, "-lcabana_b200", "-ldl",
           "-Xlinker", "-rpath", "-Xlinker", "$ORIGIN/../../cabana_b200/lib", "-cudart", "shared"]
    r = subprocess.run(cmd, capture_output=True, text=True)
    if r.returncode != 0:
        sys.stderr.write(r.stdout + r.stderr)
        raise RuntimeError("C++ shim test failed to compile")
    return CPP_TEST_BIN


CPP_COMM_SRC = os.path.join(ROOT, "tests", "cpp", "test_cabana_comm.cu")
CPP_COMM_BIN = os.path.join(ROOT, "tests", "cpp", "test_cabana_comm.bin")


def build_cpp_comm_test(force: bool = False) -> str:
    """Compile the C++ test of include/Cabana_B200_Comm.hpp (Halo / Distributor over NCCL) against
    the in-tree library and the system NCCL."""
    build()
    deps = [CPP_COMM_SRC, os.path.join(ROOT, "include", "Cabana_B200_Comm.hpp"),
            os.path.join(ROOT, "include", "Cabana_B200.hpp"),
            os.path.join(ROOT, "include", "cabana_b200.h"), LIB_PATH]
    if (not force and os.path.exists(CPP_COMM_BIN)
            and all(os.path.getmtime(CPP_COMM_BIN) >= os.path.getmtime(d) for d in deps)):
        return CPP_COMM_BIN
    cmd = [NVCC, "-gencode", "arch=compute_100a,code=sm_100a", "-O2", "-std=c++17",
           "-ccbin", HOST_CXX, "--extended-lambda", "-I", os.path.join(ROOT, "include"),
           CPP_COMM_SRC, "-o", CPP_COMM_BIN, "-L", LIB_DIR, "-lcabana_b200", "-lnccl", "-ldl",
           "-Xlinker", "-rpath", "-Xlinker", "$ORIGIN/../../cabana_b200/lib", "-cudart", "shared"]
    r = subprocess.run(cmd, capture_output=True, text=True)
    if r.returncode != 0:
        sys.stderr.write(r.stdout + r.stderr)
        raise RuntimeError("C++ comm test failed to compile")
    return CPP_COMM_BIN


if __name__ == "__main__":
    path = build(force="--force" in sys.argv, verbose="-v" in sys.argv)
    print(path)
