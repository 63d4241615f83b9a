"""Builds tests/emu/libsmooth_emu.so -- the host emulation of the mixed-radix kernel family (test-only)."""
import os
import subprocess

HERE = os.path.dirname(os.path.abspath(__file__))
ROOT = os.path.dirname(os.path.dirname(HERE))
LIB = os.path.join(HERE, "libsmooth_emu.so")
SRCS = [os.path.join(HERE, "smooth_emu.cu")]
DEPS = SRCS + [os.path.join(ROOT, "veils_b200", "csrc", f) for f in
               ("smooth_impl.cuh", "smooth_dft.cuh", "pow2_tile_impl.cuh", "pow2_cfg.hpp", "kernels.cuh", "f32x2.cuh")]


def build():
    if os.path.exists(LIB) and all(os.path.getmtime(d) <= os.path.getmtime(LIB) for d in DEPS):
        return LIB
    cmd = ["g++", "-O1", "-std=c++17", "-shared", "-fPIC", "-x", "c++", "-I/usr/local/cuda/include",
           "-o", LIB, *SRCS]
    r = subprocess.run(cmd, stdout=subprocess.PIPE, stderr=subprocess.STDOUT, text=True)
    if r.returncode != 0:
        raise RuntimeError("smooth emu build failed:\n" + r.stdout[-4000:])
    return LIB


if __name__ == "__main__":
    print(build())
