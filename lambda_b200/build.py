"""Builds lambda_b200/liblambda_b200.so (CUDA kernels + C ABI) for sm_100a with nvcc, in-tree."""
import os
import subprocess
import sys

HERE = os.path.dirname(os.path.abspath(__file__))
SRC = os.path.join(HERE, "csrc")
OUT = os.path.join(HERE, "liblambda_b200.so")
NVCC_FLAGS = ["-gencode", "arch=compute_100a,code=sm_100a", "-O3", "-lineinfo", "-std=c++17",
              "-Xcompiler", "-fPIC,-Wall,-Wextra,-Wno-unused-parameter", "-shared"]


def _newest(paths):
    return max(os.path.getmtime(p) for p in paths)


def sources():
    deps = [os.path.join(SRC, f) for f in sorted(os.listdir(SRC))]
    deps.append(os.path.join(os.path.dirname(HERE), "include", "lambda_b200.h"))
    return deps


CLI = os.path.join(os.path.dirname(HERE), "bin", "lambda3_b200")


def build(force=False, verbose=False):
    deps = sources()
    if force or not os.path.exists(OUT) or os.path.getmtime(OUT) < _newest(deps):
        nvcc = os.environ.get("NVCC", "/usr/local/cuda/bin/nvcc")
        cmd = [nvcc, *NVCC_FLAGS, os.path.join(SRC, "engine.cu"), "-o", OUT]
        if verbose:
            print(" ".join(cmd), file=sys.stderr)
        subprocess.check_call(cmd)
    # host program with the lambda3 search command line, linked against the C ABI only
    main = os.path.join(SRC, "main.cpp")
    if force or not os.path.exists(CLI) or os.path.getmtime(CLI) < max(os.path.getmtime(main), os.path.getmtime(OUT)):
        os.makedirs(os.path.dirname(CLI), exist_ok=True)
        cmd = ["/usr/bin/g++", "-std=c++17", "-O2", "-Wall", "-Wextra", main, "-o", CLI, "-L" + HERE, "-llambda_b200",
               "-Wl,-rpath,$ORIGIN/../lambda_b200", "-pthread", "-lz"]
        if verbose:
            print(" ".join(cmd), file=sys.stderr)
        subprocess.check_call(cmd)
    return OUT


if __name__ == "__main__":
    print(build(force="--force" in sys.argv, verbose=True))
