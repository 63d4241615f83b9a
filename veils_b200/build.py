"""Builds veils_b200/libveils_cuda.so (CUDA kernels + C ABI) for sm_100a with nvcc, in-tree.

    python -m veils_b200.build            # rebuild if any source is newer than the .so
"""
import os
import subprocess
import sys

HERE = os.path.dirname(os.path.abspath(__file__))
CSRC = os.path.join(HERE, "csrc")
LIB = os.path.join(HERE, "libveils_cuda.so")
CLI = os.path.join(HERE, "veils_comparison_helper")          # csrc/cli/comparison_helper.cpp
SOURCES = ["api.cu", "plan.cpp", "dft_generic.cu", "pow2_tile.cu", "pow2_pk.cu", "pad.cu", "detrend.cu",
           "wide_fwd.cu", "wide_inv.cu", "smooth.cu"]
NVCC_FLAGS = ["-gencode", "arch=compute_100a,code=sm_100a", "-lineinfo", "-O3", "-std=c++17",
              "-Xcompiler", "-fPIC,-O3,-Wall,-Wno-unknown-pragmas", "--use_fast_math=false"]


def _nvcc():
    for c in (os.environ.get("NVCC"), "/usr/local/cuda/bin/nvcc", "nvcc"):
        if c and (os.path.isabs(c) and os.path.exists(c) or not os.path.isabs(c)):
            return c
    return "nvcc"


def needs_build():
    if not os.path.exists(LIB):
        return True
    t = os.path.getmtime(LIB)
    if not os.path.exists(CLI):
        return True
    deps = [os.path.join(CSRC, f) for f in os.listdir(CSRC) if os.path.isfile(os.path.join(CSRC, f))]
    deps.append(os.path.join(CSRC, "cli", "comparison_helper.cpp"))
    deps.append(os.path.join(HERE, "..", "include", "veils_cuda.h"))
    deps.append(os.path.abspath(__file__))
    return any(os.path.getmtime(d) > t for d in deps)


def build(force=False, verbose=False):
    if not force and not needs_build():
        return LIB
    objs = []
    procs = []
    os.makedirs(os.path.join(HERE, "build"), exist_ok=True)
    flags = [f for f in NVCC_FLAGS if not f.startswith("--use_fast_math")]
    flags += os.environ.get("VEILS_EXTRA_NVCC_FLAGS", "").split()
    for src in SOURCES:
        obj = os.path.join(HERE, "build", src + ".o")
        cmd = [_nvcc(), *flags, "-Xptxas", "-v", "-c", os.path.join(CSRC, src), "-o", obj]
        procs.append((src, subprocess.Popen(cmd, stdout=subprocess.PIPE, stderr=subprocess.STDOUT, text=True)))
        objs.append(obj)
    log = []
    for src, p in procs:
        out, _ = p.communicate()
        log.append(f"==== {src}\n{out}")
        if p.returncode != 0:
            sys.stderr.write(out)
            raise RuntimeError(f"nvcc failed on {src}")
    with open(os.path.join(HERE, "build", "ptxas.log"), "w") as f:
        f.write("\n".join(log))
    cmd = [_nvcc(), "-shared", "-o", LIB, *objs, "-gencode", "arch=compute_100a,code=sm_100a"]
    r = subprocess.run(cmd, stdout=subprocess.PIPE, stderr=subprocess.STDOUT, text=True)
    if r.returncode != 0:
        sys.stderr.write(r.stdout)
        raise RuntimeError("link failed")
    # the comparison-helper CLI (reference: src/bin/python_rust_comparison_helper.rs) links the C ABI
    cmd = ["g++", "-O2", "-std=c++17", "-Wall", "-o", CLI, os.path.join(CSRC, "cli", "comparison_helper.cpp"),
           "-L" + HERE, "-l:libveils_cuda.so", "-Wl,-rpath,$ORIGIN"]
    r = subprocess.run(cmd, stdout=subprocess.PIPE, stderr=subprocess.STDOUT, text=True)
    if r.returncode != 0:
        sys.stderr.write(r.stdout)
        raise RuntimeError("CLI build failed")
    if verbose:
        print("\n".join(log))
    return LIB


if __name__ == "__main__":
    print(build(force="--force" in sys.argv, verbose="-v" in sys.argv))
