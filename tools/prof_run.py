"""Small driver for ncu captures: runs the c2-shaped forward and inverse kernels a few times.
    python tools/prof_run.py [--batch 128] [--config c2] [--reps 3]
"""
import argparse
import os
import sys

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch

import bench
from veils_b200 import StandaloneSTFT, lib

ap = argparse.ArgumentParser()
ap.add_argument("--batch", type=int, default=128)
ap.add_argument("--config", default="c2")
ap.add_argument("--reps", type=int, default=3)
ap.add_argument("--ldp-align", type=int, default=1)
ap.add_argument("--m", type=int, default=0, help="override m_num = mfft (hop = m / 4)")
ap.add_argument("--n", type=int, default=0)
ap.add_argument("--mode", default="")
ap.add_argument("--hop", type=int, default=0)
ap.add_argument("--prec", default="")
ap.add_argument("--spec", action="store_true", help="spectrogram (forward only)")
args = ap.parse_args()
cfg = dict(bench.CONFIGS[args.config])
cfg["batch"] = args.batch
if args.m:
    cfg["m"], cfg["hop"] = args.m, args.m // 4
if args.n:
    cfg["n"] = args.n
if args.mode:
    cfg["mode"] = args.mode
if args.hop:
    cfg["hop"] = args.hop
if args.prec:
    cfg["prec"] = args.prec
dev = torch.device("cuda", 0)
st = StandaloneSTFT(bench.make_window(cfg), cfg["hop"], cfg["fs"], fft_mode=cfg["mode"], precision=cfg["prec"],
                    scale_to=cfg.get("scale"))
x = bench.synth_device(cfg, torch, dev, 1)
B, n = x.shape
P, F = st.p_num(n), st.f_pts
ldp = (P + args.ldp_align - 1) // args.ldp_align * args.ldp_align
k0, k1 = st.istft_range(P)
S = torch.empty((B, F, ldp), dtype=torch.complex64 if cfg["prec"] == "f32" else torch.complex128, device=dev)
y = torch.empty((B, k1 - k0), dtype=x.dtype, device=dev)
L = lib()
if args.spec:
    sz = 4 if cfg["prec"] == "f32" else 8
    ldp = (P + 128 // sz - 1) // (128 // sz) * (128 // sz)
    Sx = torch.empty((B, F, ldp), dtype=x.dtype, device=dev)
    for _ in range(args.reps):
        assert L.veils_spectrogram_dev(st._h, x.data_ptr(), n, B, n, None, None, 0, Sx.data_ptr(), ldp, None) == 0
    torch.cuda.synchronize()
    print("ok spectrogram", B, n, P, F, ldp)
    sys.exit(0)
for _ in range(args.reps):
    assert L.veils_stft_dev(st._h, x.data_ptr(), n, B, n, None, None, 0, S.data_ptr(), ldp, None) == 0
    assert L.veils_istft_dev(st._h, S.data_ptr(), P, B, ldp, None, None, y.data_ptr(), k1 - k0, None) == 0
torch.cuda.synchronize()
print("ok", B, n, P, F, ldp, float((y[:, :n] - x).abs().max()))
