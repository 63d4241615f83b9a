"""Runs the c2-shaped forward and inverse kernels separately with a sync after each (debug aid)."""
import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np, torch
from veils_b200 import StandaloneSTFT, lib
L = lib()
dev = torch.device("cuda", 0)
win = 0.5 * (1 - np.cos(2 * np.pi * np.arange(512) / 512))
st = StandaloneSTFT(win, 128, 16000.0, precision="f32")
B, n = 64, 160000
x = torch.rand((B, n), device=dev) * 2 - 1
P, F = st.p_num(n), st.f_pts
ldp = (P + 3) // 4 * 4
S = torch.zeros((B, F, ldp), dtype=torch.complex64, device=dev)
k0, k1 = st.istft_range(P)
y = torch.empty((B, k1 - k0), device=dev)
print("stft rc", L.veils_stft_dev(st._h, x.data_ptr(), n, B, n, None, None, 0, S.data_ptr(), ldp, None), flush=True)
torch.cuda.synchronize(); print("stft ok", flush=True)
print("istft rc", L.veils_istft_dev(st._h, S.data_ptr(), P, B, ldp, None, None, y.data_ptr(), k1 - k0, None), flush=True)
torch.cuda.synchronize(); print("istft ok, err", float((y[:, :n] - x).abs().max()), flush=True)
