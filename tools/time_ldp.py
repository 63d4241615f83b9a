"""Times the c2 forward / inverse kernels for dense (ldp = P) and aligned row pitches."""
import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
import bench
from veils_b200 import StandaloneSTFT, lib

cfg = dict(bench.CONFIGS["c2"])
cfg["batch"] = int(sys.argv[1]) if len(sys.argv) > 1 else 1024
dev = torch.device("cuda", 0)
st = StandaloneSTFT(bench.make_window(cfg), cfg["hop"], cfg["fs"], fft_mode=cfg["mode"], precision=cfg["prec"])
x = bench.synth_device(cfg, torch, dev, 1)
B, n = x.shape
P, F = st.p_num(n), st.f_pts
k0, k1 = st.istft_range(P)
L = lib()
y = torch.empty((B, k1 - k0), dtype=x.dtype, device=dev)
for align in (1, 2, 4, 16, 32):
    ldp = (P + align - 1) // align * align
    S = torch.empty((B, F, ldp), dtype=torch.complex64, device=dev)
    def fwd():
        assert L.veils_stft_dev(st._h, x.data_ptr(), n, B, n, None, None, 0, S.data_ptr(), ldp, None) == 0
    def inv():
        assert L.veils_istft_dev(st._h, S.data_ptr(), P, B, ldp, None, None, y.data_ptr(), k1 - k0, None) == 0
    res = []
    for fn in (fwd, inv):
        for _ in range(3): fn()
        torch.cuda.synchronize()
        a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        a.record()
        for _ in range(10): fn()
        b.record(); torch.cuda.synchronize()
        res.append(a.elapsed_time(b) / 10)
    err = float((y[:, :n] - x).abs().max())
    print(f"ldp={ldp} (align {align}): fwd {res[0]:.3f} ms  inv {res[1]:.3f} ms  recon {err:.2e}", flush=True)
    del S
