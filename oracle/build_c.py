"""Builds oracle/_build/libstft_ref.so from oracle/stft_ref.c (gcc -O3 -march=native -fopenmp).
TEST INFRASTRUCTURE: the C restatement of the reference CPU path (checker + CPU baseline)."""
import os
import subprocess

HERE = os.path.dirname(os.path.abspath(__file__))
SRC = os.path.join(HERE, "stft_ref.c")
OUT_DIR = os.path.join(HERE, "_build")
LIB = os.path.join(OUT_DIR, "libstft_ref.so")


def build(force=False):
    os.makedirs(OUT_DIR, exist_ok=True)
    if not force and os.path.exists(LIB) and os.path.getmtime(LIB) >= os.path.getmtime(SRC):
        return LIB
    # -march=native is deliberately NOT used: the .so is built in the CPU container and travels
    # to the GPU box, whose host CPU may differ; x86-64-v3 (AVX2/FMA) is safe on both.
    cmd = ["gcc", "-O3", "-march=x86-64-v3", "-fopenmp", "-shared", "-fPIC", "-o", LIB, SRC, "-lm"]
    r = subprocess.run(cmd, stdout=subprocess.PIPE, stderr=subprocess.STDOUT, text=True)
    if r.returncode != 0:
        raise RuntimeError("gcc failed:\n" + r.stdout)
    return LIB


if __name__ == "__main__":
    print(build(force=True))
