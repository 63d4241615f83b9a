#!/usr/bin/env python
"""bench.py -- BASELINE.json's metric: STFT+iSTFT Msamples/s (input samples per second for one
forward + one inverse transform), % of the HBM roofline, next to the reference's CPU path timed on
this box's host cores.

    python bench.py [--gpus N] [--steps K] [--warmup W] [--impl ours|reference] [--config c2]
    torchrun --nproc-per-node N ... bench.py --gpus N ...          (one rank per GPU, NCCL)

Headline (`value`, `roofline`, `e2e`): BASELINE.json configs[1] ("c2"): batch 1024 x 10 s of 16 kHz
speech-shaped noise, periodic Hann 512 / hop 128, onesided, f32.  A step = one stft + one istft over
the whole batch through the C ABI (veils_stft_dev / veils_istft_dev), inputs resident in HBM.  With
N > 1 every rank owns its own batch of that shape (rows are independent -> no data-path collective,
"weak" scaling); time = max over ranks.

The same JSON line carries, under `configs`, the other BASELINE shapes, each with its own roofline
and an oracle parity check on first / last / interior slices:
  c1  1 s @ 44.1 kHz, Hann 2048 / 512, f64: single-signal latency in microseconds;
  c3  the FIXED batch of 16 384 rows x 30 s @ 48 kHz (8192 stereo items), Hann 4096 / 1024, round trip:
      rows sharded over the ranks (sharding.row_range), streamed through fixed device buffers in
      sub-batches -> `scaling: "strong"`;
  c4  ONE long signal, Gaussian 1024 / 256, twosided: 2^32 samples at N = 8 (2^29 per GPU below that),
      sharded by slice range with an (m_num - hop)-sample halo; the inverse exchanges the halo
      columns of S with the neighbour rank over NCCL send/recv (the one data-path exchange, timed);
  c5  Hann 8192 / 64 onesided2X psd spectrogram, f64 (FP64-bound shape): a sub-batch.
One JSON line is printed by rank 0.
"""
import argparse
import json
import os
import sys
import threading
import time

import numpy as np

ROOT = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, ROOT)

CONFIGS = {
    # name: batch, n, fs, m_num, hop, mode, precision, window
    "c1": dict(batch=1, n=44100, fs=44100.0, m=2048, hop=512, mode="onesided", prec="f64", win="hann"),
    "c2": dict(batch=1024, n=160000, fs=16000.0, m=512, hop=128, mode="onesided", prec="f32", win="hann"),
    "c3": dict(batch=16384, n=1440000, fs=48000.0, m=4096, hop=1024, mode="onesided", prec="f32", win="hann"),
    "c4": dict(batch=1, n=1 << 32, fs=48000.0, m=1024, hop=256, mode="twosided", prec="f32", win="gauss"),
    "c5": dict(batch=4096, n=160000, fs=16000.0, m=8192, hop=64, mode="onesided2X", prec="f64", win="hann",
               scale="psd", spectrogram=True),
}
FP64_PEAK_TFLOPS = 37.0       # B200 DFMA peak used for the c5 compute roof (measured: profiles/r1_microbench.txt 33.9)


def make_window(cfg):
    m = cfg["m"]
    if cfg["win"] == "hann":
        return 0.5 * (1.0 - np.cos(2.0 * np.pi * np.arange(m) / m))
    k = np.arange(m) - m / 2.0
    return np.exp(-0.5 * (k / (m / 8.0)) ** 2)


def synth_host(cfg, rows, seed):
    """Speech-shaped noise (AR(1), a = 0.95, unit peak) for `rows` signals, host float64."""
    rng = np.random.default_rng(seed)
    from scipy.signal import lfilter
    e = rng.uniform(-1.0, 1.0, (rows, cfg["n"]))
    y = lfilter([1.0], [1.0, -0.95], e, axis=1)
    return y / np.max(np.abs(y), axis=1, keepdims=True)


def synth_device(cfg, torch, device, seed):
    """Same shape of signal generated on the device: U(-1,1) noise through a truncated AR(1)
    impulse response (128 taps of 0.95^k), scaled to unit peak per row."""
    g = torch.Generator(device=device)
    g.manual_seed(seed)
    dt = torch.float32 if cfg["prec"] == "f32" else torch.float64
    B, n = cfg["batch"], cfg["n"]
    x = torch.empty((B, n), dtype=dt, device=device)
    taps = (0.95 ** torch.arange(127, -1, -1, dtype=torch.float32, device=device)).view(1, 1, -1)
    rows = max(1, min(B, (1 << 26) // n))
    for a in range(0, B, rows):
        b = min(B, a + rows)
        e = torch.rand((b - a, 1, n + 127), generator=g, device=device, dtype=torch.float32) * 2 - 1
        y = torch.nn.functional.conv1d(e, taps).squeeze(1)
        y = y / y.abs().amax(dim=1, keepdim=True)
        x[a:b] = y.to(dt)
    return x


# ---- counter-based noise: sample i of the long c4 signal depends on i only, so every rank (and the
#      host-side oracle check) generates exactly the samples of its own range
def hash_noise_np(i0, i1, seed=12345):
    i = np.arange(i0, i1, dtype=np.uint64)
    h = ((i & np.uint64(0xFFFFFFFF)) * np.uint64(2654435761) + (i >> np.uint64(32)) * np.uint64(40503)
         + np.uint64(seed)) & np.uint64(0xFFFFFFFF)
    h ^= h >> np.uint64(16)
    h = (h * np.uint64(2246822519)) & np.uint64(0xFFFFFFFF)
    h ^= h >> np.uint64(13)
    h = (h * np.uint64(3266489917)) & np.uint64(0xFFFFFFFF)
    h ^= h >> np.uint64(16)
    return (h.astype(np.float64) / 2147483648.0 - 1.0).astype(np.float32)


def hash_noise_torch(torch, dev, i0, i1, seed=12345, out=None):
    n = i1 - i0
    x = out if out is not None else torch.empty((n,), dtype=torch.float32, device=dev)
    step = 1 << 26
    for a in range(0, n, step):
        b = min(n, a + step)
        i = torch.arange(i0 + a, i0 + b, device=dev, dtype=torch.int64)
        h = ((i & 0xFFFFFFFF) * 2654435761 + (i >> 32) * 40503 + seed) & 0xFFFFFFFF
        h = h ^ (h >> 16)
        h = (h * 2246822519) & 0xFFFFFFFF
        h = h ^ (h >> 13)
        h = (h * 3266489917) & 0xFFFFFFFF
        h = h ^ (h >> 16)
        x[a:b] = (h.to(torch.float64) / 2147483648.0 - 1.0).to(torch.float32)
    return x


class ClockSampler:
    """Samples SM clock / power / throttle reasons DURING the timed region through NVML
    (the same counters as the nvidia-smi line of B200_PROFILING.md, at ~2 ms period)."""
    REASONS = {0x8: "hw_slowdown", 0x40: "hw_thermal_slowdown", 0x20: "sw_thermal_slowdown",
               0x4: "sw_power_cap"}

    def __init__(self, index):
        self.index = index
        self.rows = []
        self.ok = False
        self._stop = False

    def start(self):
        try:
            import pynvml
            pynvml.nvmlInit()
            self.nv = pynvml
            self.h = pynvml.nvmlDeviceGetHandleByIndex(self.index)
            self.max_sm = pynvml.nvmlDeviceGetMaxClockInfo(self.h, pynvml.NVML_CLOCK_SM)
            self.ok = True
            self.t = threading.Thread(target=self._run, daemon=True)
            self.t.start()
        except Exception as e:                      # noqa: BLE001
            self.err = str(e)

    def _run(self):
        nv = self.nv
        while not self._stop:
            try:
                sm = nv.nvmlDeviceGetClockInfo(self.h, nv.NVML_CLOCK_SM)
                try:
                    rs = nv.nvmlDeviceGetCurrentClocksEventReasons(self.h)
                except Exception:                   # noqa: BLE001
                    rs = nv.nvmlDeviceGetCurrentClocksThrottleReasons(self.h)
                pw = nv.nvmlDeviceGetPowerUsage(self.h) / 1000.0
                self.rows.append((sm, rs, pw))
            except Exception:                       # noqa: BLE001
                pass
            time.sleep(0.002)

    def mark(self):
        return len(self.rows)

    def stop(self, lo=0, hi=None):
        self._stop = True
        if not self.ok:
            return {"sm_mhz": None, "sm_max_mhz": None, "reasons": ["nvml unavailable"]}
        self.t.join(1.0)
        rows = self.rows[lo:hi] or self.rows
        reasons = set()
        for _, rs, _ in rows:
            for bit, name in self.REASONS.items():
                if rs & bit:
                    reasons.add(name)
        sm = [r[0] for r in rows]
        return {"sm_mhz": float(np.median(sm)) if sm else None, "sm_max_mhz": float(self.max_sm),
                "reasons": sorted(reasons), "samples": len(sm),
                "power_w_max": max((r[2] for r in rows), default=None)}


def measured_peak():
    p = os.path.join(ROOT, "MEASURED_PEAKS.json")
    try:
        return float(json.load(open(p))["hbm_gbs"]), "measured (MEASURED_PEAKS.json hbm_gbs)"
    except Exception:
        return 6650.0, "fallback (B200_PROFILING.md)"


def algorithmic_bytes(cfg, B, n, P, F, n_out):
    """forward B(n sz + F P 2 sz), spectrogram B(n sz + F P sz), inverse B(F P 2 sz + n_out sz)."""
    sz = 4 if cfg["prec"] == "f32" else 8
    if cfg.get("spectrogram"):
        return B * (n * sz + F * P * sz), 0
    return B * (n * sz + F * P * 2 * sz), B * (F * P * 2 * sz + n_out * sz)


def host_threads():
    """Threads the CPU arm uses: every core this process may run on -- set explicitly, so that
    OMP_NUM_THREADS=1 exported by torchrun does not change the baseline between N = 1 and N > 1."""
    try:
        return max(1, len(os.sched_getaffinity(0)))
    except Exception:
        return max(1, os.cpu_count() or 1)


def cpu_model():
    try:
        for line in open("/proc/cpuinfo"):
            if line.startswith("model name"):
                return line.split(":", 1)[1].strip()
    except Exception:
        pass
    return "unknown"


def oracle_of(cfg):
    from oracle.stft_oracle import OracleSTFT
    return OracleSTFT(make_window(cfg), cfg["hop"], cfg["fs"], fft_mode=cfg["mode"], scale_to=cfg.get("scale"))


def cpu_reference_run(cfg, rows, threads, seed=1234, reps=1):
    """Round trip of `rows` signals with the C restatement of the reference (oracle/stft_ref.c)."""
    from oracle.ref_c import RefC
    x = synth_host(cfg, rows, seed)
    c = RefC(oracle_of(cfg))
    best = None
    for _ in range(reps):
        t0 = time.perf_counter()
        S = c.stft(x, threads=threads)
        y = c.istft(S, threads=threads)
        dt = time.perf_counter() - t0
        best = dt if best is None else min(best, dt)
    return x, S, y, best


def cpu_context(cfg, budget_s=4.0):
    """Context numbers for the CPU side, each on a bounded sample (Msamples/s, stft + istft):
    the C port single-threaded (the reference IS single-threaded), scipy.signal.ShortTimeFFT on one
    thread (the reference's Python twin restates exactly this class; bit-identical)
    and torch.stft / istft on all host threads (best readily available CPU library)."""
    out = {}
    n = cfg["n"]
    try:
        _, _, _, dt = cpu_reference_run(cfg, 2, 1, seed=5)
        out["port_1_thread"] = 2 * n / dt / 1e6
    except Exception as e:                          # noqa: BLE001
        out["port_1_thread_error"] = str(e)[:100]
    try:
        from scipy.signal import ShortTimeFFT
        x = synth_host(cfg, 1, 6)[0]
        sft = ShortTimeFFT(make_window(cfg), cfg["hop"], cfg["fs"], fft_mode=cfg["mode"])
        t0 = time.perf_counter()
        S = sft.stft(x)
        sft.istft(S)
        out["scipy_twin_1_thread"] = n / (time.perf_counter() - t0) / 1e6
    except Exception as e:                          # noqa: BLE001
        out["scipy_twin_error"] = str(e)[:100]
    try:
        import torch
        nt = host_threads()
        torch.set_num_threads(nt)
        rows = 64
        x = torch.from_numpy(synth_host(cfg, rows, 7).astype(np.float32))
        w = torch.from_numpy(make_window(cfg).astype(np.float32))
        best = None
        for _ in range(3):
            t0 = time.perf_counter()
            S = torch.stft(x, cfg["m"], cfg["hop"], window=w, center=True, return_complex=True)
            torch.istft(S, cfg["m"], cfg["hop"], window=w, center=True)
            dt = time.perf_counter() - t0
            best = dt if best is None else min(best, dt)
        out["torch_stft_cpu_all_threads"] = rows * n / best / 1e6
        out["torch_threads"] = nt
    except Exception as e:                          # noqa: BLE001
        out["torch_stft_error"] = str(e)[:100]
    return out


def workload_config(cfg, name):
    return {"workload": f"{name}: batch {cfg['batch']} x {cfg['n']} samples @ {cfg['fs']:.0f} Hz, "
                        f"{cfg['win']} {cfg['m']} / hop {cfg['hop']}, {cfg['mode']}, {cfg['prec']}, stft+istft",
            "batch": cfg["batch"], "n": cfg["n"], "m_num": cfg["m"], "hop": cfg["hop"],
            "fft_mode": cfg["mode"], "l2_policy": "working set (input + S) >> 126 MB L2, no flush needed",
            "s_row_pitch": "rows padded to 128-byte lines (ldp = ceil(P/16)*16 for f32); --dense-pitch for ldp = P"}


def run_reference(args, cfg, rank, world):
    """--impl reference: the reference's CPU path (C restatement; the Rust crate cannot be built in
    this image) on ALL host threads -- the same explicit thread count at every N -- on a bounded sample
    per step.  Rank 0 only; the context numbers (single thread, scipy twin, torch.stft) ride along."""
    if rank != 0:
        return
    threads = host_threads()
    rows = max(threads, 8)
    _, _, _, t1 = cpu_reference_run(cfg, rows, threads, reps=2)
    rows = int(max(rows, min(cfg["batch"], rows * 2.0 / max(t1, 1e-3))))      # ~2 s per step
    times = []
    for i in range(args.warmup + args.steps):
        _, _, _, dt = cpu_reference_run(cfg, rows, threads, seed=100 + i)
        if i >= args.warmup:
            times.append(dt)
    ms = 1e3 * float(np.mean(times))
    val = rows * cfg["n"] / (ms * 1e-3) / 1e6
    sample = f"{rows} of {cfg['batch']} rows of {args.config} per step (stft+istft, f64, C restatement of src/lib.rs)"
    print(json.dumps({
        "impl": "reference", "metric": "stft_istft_throughput", "value": val, "unit": "Msamples/s",
        "n_gpus": args.gpus, "steps": args.steps, "warmup": args.warmup, "ms_per_step": ms,
        "higher_is_better": True, "scaling": "weak", "vs_baseline": None, "dtype": "f64",
        "data": "synthetic", "config": workload_config(cfg, args.config),
        "cpu_baseline": {"value": val, "unit": "Msamples/s", "cores": threads, "kind": "port",
                         "sample": sample, "context_msamples_per_s": cpu_context(cfg),
                         "note": "the reference itself is single-threaded (port_1_thread is its shape); "
                                 "all host threads are used here because the arm may use them"},
        "e2e": {"value": val, "unit": "Msamples/s", "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
        "host": {"nproc": os.cpu_count(), "threads_used": threads, "cpu": cpu_model()}}))


# =========================================================================== GPU side helpers
class Env:
    pass


def gpu_env(local, world):
    import torch
    import torch.distributed as dist
    from veils_b200 import StandaloneSTFT, lib, sharding
    e = Env()
    e.torch, e.dist, e.STFT, e.L, e.sharding = torch, dist, StandaloneSTFT, lib(), sharding
    assert torch.cuda.is_available(), "bench.py needs a CUDA device (no CPU fallback)"
    torch.cuda.set_device(local)
    e.dev = torch.device("cuda", local)
    e.local, e.world = local, world
    e.rank = int(os.environ.get("RANK", "0"))
    if world > 1:
        os.environ.setdefault("MASTER_ADDR", "127.0.0.1")
        dist.init_process_group("nccl", device_id=e.dev)
    e.peak, e.peak_src = measured_peak()
    return e


def barrier(e):
    if e.world > 1:
        e.dist.barrier()
    e.torch.cuda.synchronize()


def max_over_ranks(e, v):
    t = e.torch.tensor([float(v)], device=e.dev, dtype=e.torch.float64)
    if e.world > 1:
        e.dist.all_reduce(t, op=e.dist.ReduceOp.MAX)
    return float(t.item())


def plan_of(e, cfg):
    return e.STFT(make_window(cfg), cfg["hop"], cfg["fs"], fft_mode=cfg["mode"], scale_to=cfg.get("scale"),
                  precision=cfg["prec"], device=e.local)


def rel_err(a, b):
    return float(np.max(np.abs(a - b)) / max(np.max(np.abs(b)), 1e-300))


def parity_rows(e, cfg, st, x, S, y, rows, n_slices=24):
    """Oracle check on the SAME device buffers: first / interior / last slices of the given rows of
    S (and the matching stretches of the reconstruction) against the numpy oracle evaluated on the
    samples those slices cover.  Returns the worst relative errors."""
    o = oracle_of(cfg)
    n = x.shape[1]
    P = o.p_num(n)
    N, H, mid = cfg["m"], cfg["hop"], cfg["m"] // 2
    worst_s = worst_y = 0.0
    for r in rows:
        for pa in sorted({o.p_min, o.p_min + (P - n_slices) // 2, o.p_min + P - n_slices}):
            pb = pa + n_slices
            s_a, s_b = max(0, pa * H - mid), min(n, (pb - 1) * H - mid + N)
            xs = x[r, s_a:s_b].double().cpu().numpy()
            ref = o.stft(xs, pa, pb, k_offset=-s_a)
            got = S[r, :, pa - o.p_min:pb - o.p_min].cpu().numpy()
            worst_s = max(worst_s, rel_err(got, ref))
            if y is not None:
                # samples whose contributing slices all lie inside [pa, pb): oracle istft of the got columns
                k_lo, k_hi = (pa - 1) * H + N - mid, pb * H - mid
                k_lo, k_hi = max(k_lo, 0), min(k_hi, n)
                if k_hi - k_lo >= N - mid:
                    shift = (pa - o.p_min) * H
                    yr = o.istft(got.astype(np.complex128), k_lo - shift, k_hi - shift)
                    worst_y = max(worst_y, rel_err(y[r, k_lo:k_hi].cpu().numpy(), yr))
    return worst_s, worst_y


def time_events(e, fn, reps):
    a, b = e.torch.cuda.Event(enable_timing=True), e.torch.cuda.Event(enable_timing=True)
    fn()
    e.torch.cuda.synchronize()
    a.record()
    for _ in range(reps):
        fn()
    b.record()
    e.torch.cuda.synchronize()
    return a.elapsed_time(b) / reps


def roof_entry(e, nbytes, ms):
    return {"ms": ms, "GBps": nbytes / (ms * 1e-3) / 1e9, "frac": nbytes / (ms * 1e-3) / 1e9 / e.peak}


# ---- c1: one second of audio, f64 -- latency-bound (1.8 MB): microseconds per call
def side_c1(e):
    cfg = CONFIGS["c1"]
    st = plan_of(e, cfg)
    torch, L = e.torch, e.L
    rng = np.random.default_rng(42)
    n = cfg["n"]
    xh = np.sin(2 * np.pi * 440 * np.arange(n) / cfg["fs"]) + 0.1 * rng.standard_normal(n)
    x = torch.from_numpy(xh).to(e.dev).view(1, n)
    P, F = st.p_num(n), st.f_pts
    k0, k1 = st.istft_range(P)
    ldp = -(-P // st.pitch_align) * st.pitch_align
    S = torch.empty((1, F, ldp), dtype=torch.complex128, device=e.dev)
    y = torch.empty((1, k1 - k0), dtype=torch.float64, device=e.dev)
    stream = torch.cuda.current_stream(e.dev).cuda_stream

    def fwd():
        assert L.veils_stft_dev(st._h, x.data_ptr(), n, 1, n, None, None, 0, S.data_ptr(), ldp, stream) == 0

    def inv():
        assert L.veils_istft_dev(st._h, S.data_ptr(), P, 1, ldp, None, None, y.data_ptr(), k1 - k0, stream) == 0
    us_f, us_i = 1e3 * time_events(e, fwd, 200), 1e3 * time_events(e, inv, 200)
    # host cost per call: 1000 back-to-back launches, wall clock / launch (the stream never runs dry)
    torch.cuda.synchronize()
    t0 = time.perf_counter()
    for _ in range(1000):
        fwd()
    host_us = (time.perf_counter() - t0) / 1000 * 1e6
    torch.cuda.synchronize()
    import ctypes
    us_c = ctypes.c_double(0.0)
    assert L.veils_debug_launch_cost_us(st._h, x.data_ptr(), n, 1, n, S.data_ptr(), ldp, stream, 1000, ctypes.byref(us_c)) == 0
    torch.cuda.synchronize()
    # the round trip as a CUDA graph (both launches captured once, replayed): what a latency-bound caller would do
    graph_us = None
    try:
        gs = torch.cuda.Stream(e.dev)
        with torch.cuda.stream(gs):
            gstream = gs.cuda_stream
            g = torch.cuda.CUDAGraph()
            with torch.cuda.graph(g, stream=gs):
                assert L.veils_stft_dev(st._h, x.data_ptr(), n, 1, n, None, None, 0, S.data_ptr(), ldp, gstream) == 0
                assert L.veils_istft_dev(st._h, S.data_ptr(), P, 1, ldp, None, None, y.data_ptr(), k1 - k0, gstream) == 0
            for _ in range(5):
                g.replay()
            gs.synchronize()
            gl = []
            for _ in range(50):
                t0 = time.perf_counter()
                g.replay()
                gs.synchronize()
                gl.append(time.perf_counter() - t0)
            graph_us = {"median": 1e6 * float(np.median(gl)), "min": 1e6 * float(np.min(gl))}
    except Exception as ex:                         # noqa: BLE001
        graph_us = {"error": str(ex)[:120]}
    torch.cuda.synchronize()
    # wall latency of one synchronous round trip
    lat = []
    for _ in range(50):
        t0 = time.perf_counter()
        fwd()
        inv()
        torch.cuda.synchronize()
        lat.append(time.perf_counter() - t0)
    o = oracle_of(cfg)
    Sref = o.stft(xh)
    es = rel_err(S[0, :, :P].cpu().numpy(), Sref)
    ey = rel_err(y[0].cpu().numpy(), o.istft(Sref))
    bf, bi = algorithmic_bytes(cfg, 1, n, P, F, k1 - k0)
    return {"workload": "c1: 1 x 44100 @ 44.1 kHz, Hann 2048 / 512 onesided, f64 (latency-bound, 1.8 MB)",
            "kernel_us": {"stft": us_f, "istft": us_i},
            "host_us_per_launch": {"inside_library": us_c.value, "through_python_ctypes": host_us,
                                   "how": "1000 back-to-back veils_stft_dev calls, wall clock / 1000, no sync"},
            "roundtrip_wall_us": {"median": 1e6 * float(np.median(lat)), "min": 1e6 * float(np.min(lat))},
            "roundtrip_cuda_graph_wall_us": graph_us,
            "roofline": {"bound": "latency", "stft": roof_entry(e, bf, us_f * 1e-3), "istft": roof_entry(e, bi, us_i * 1e-3)},
            "parity": {"stft_rel_err_vs_oracle": es, "istft_rel_err_vs_oracle": ey, "tolerance": 1e-12,
                       "recon_max_abs_err": float(np.max(np.abs(y[0, :n].cpu().numpy() - xh)))},
            "kernels": {"stft": st.kernel_name(False), "istft": st.kernel_name(True)}}


# ---- c3: the fixed 16 384-row batch, rows sharded over the ranks, streamed in sub-batches
def side_c3(e, args):
    cfg = dict(CONFIGS["c3"])
    if args.c3_rows:
        cfg["batch"] = args.c3_rows
    torch, L = e.torch, e.L
    st = plan_of(e, cfg)
    n = cfg["n"]
    ra, rb = e.sharding.row_range(cfg["batch"], e.rank, e.world)
    rows = rb - ra
    sub = min(args.c3_sub, max(rows, 1))
    P, F = st.p_num(n), st.f_pts
    k0, k1 = st.istft_range(P)
    n_out = k1 - k0
    ldp = -(-P // st.pitch_align) * st.pitch_align
    free, _ = torch.cuda.mem_get_info()
    need = rows * n * 4 + sub * (F * ldp * 8 + n_out * 4) + (2 << 30)
    if need > free:
        return {"skipped": f"needs {need / 2**30:.0f} GiB of device memory, {free / 2**30:.0f} free"}
    # this rank's rows, resident in HBM (U(-1,1) noise, BASELINE: synthetic)
    g = torch.Generator(device=e.dev)
    g.manual_seed(3000 + e.rank)
    x = torch.empty((rows, n), dtype=torch.float32, device=e.dev)
    for a in range(0, rows, 64):
        b = min(rows, a + 64)
        x[a:b] = torch.rand((b - a, n), generator=g, device=e.dev, dtype=torch.float32) * 2 - 1
    S = torch.empty((sub, F, ldp), dtype=torch.complex64, device=e.dev)
    y = torch.empty((sub, n_out), dtype=torch.float32, device=e.dev)
    stream = torch.cuda.current_stream(e.dev).cuda_stream

    def fwd(a, b):
        assert L.veils_stft_dev(st._h, x[a:b].data_ptr(), n, b - a, n, None, None, 0, S.data_ptr(), ldp, stream) == 0

    def inv(a, b):
        assert L.veils_istft_dev(st._h, S.data_ptr(), P, b - a, ldp, None, None, y.data_ptr(), n_out, stream) == 0

    def whole():
        for a in range(0, rows, sub):
            b = min(rows, a + sub)
            fwd(a, b)
            inv(a, b)
    whole()
    barrier(e)
    steps = max(1, min(args.steps, args.c3_steps))
    launches0 = L.veils_kernel_launch_count()
    ev = [torch.cuda.Event(enable_timing=True) for _ in range(2)]
    ev[0].record()
    for _ in range(steps):
        whole()
    ev[1].record()
    barrier(e)
    launches = L.veils_kernel_launch_count() - launches0
    ms = max_over_ranks(e, ev[0].elapsed_time(ev[1])) / steps
    # kernel split on one sub-batch
    b0 = min(rows, sub)
    ms_f = time_events(e, lambda: fwd(0, b0), 3)
    ms_i = time_events(e, lambda: inv(0, b0), 3)
    bf, bi = algorithmic_bytes(cfg, b0, n, P, F, n_out)
    # parity: first and last row of this rank (the last sub-batch is what S / y hold now)
    last_a = ((rows - 1) // sub) * sub
    fwd(last_a, rows)
    inv(last_a, rows)
    torch.cuda.synchronize()
    es1, ey1 = parity_rows(e, cfg, st, x[last_a:rows], S, y, [rows - 1 - last_a])
    recon = float((y[:rows - last_a, :n] - x[last_a:rows]).abs().max().item())
    fwd(0, b0)
    inv(0, b0)
    torch.cuda.synchronize()
    es0, ey0 = parity_rows(e, cfg, st, x[0:b0], S, y, [0])
    recon = max(recon, float((y[:b0, :n] - x[0:b0]).abs().max().item()))
    es, ey, recon = max_over_ranks(e, max(es0, es1)), max_over_ranks(e, max(ey0, ey1)), max_over_ranks(e, recon)
    total = cfg["batch"] * n
    bfa, bia = algorithmic_bytes(cfg, cfg["batch"], n, P, F, n_out)
    return {"workload": f"c3: FIXED batch {cfg['batch']} rows (8192 stereo items) x {n} @ 48 kHz, Hann 4096 / hop 1024 "
                        f"onesided f32, stft+istft; rows sharded over {e.world} rank(s) (sharding.row_range), "
                        f"streamed through fixed device buffers in sub-batches of {sub} rows; inputs resident in HBM",
            "scaling": "strong", "value": total / (ms * 1e-3) / 1e6, "unit": "Msamples/s", "ms_per_pass": ms,
            "steps": steps, "rows_per_rank": rows, "gpu_launches": int(launches),
            "roofline": {"bound": "hbm", "peak": e.peak, "unit": "GB/s",
                         "stft": roof_entry(e, bf, ms_f), "istft": roof_entry(e, bi, ms_i),
                         "roundtrip_frac_per_gpu": (bfa + bia) / e.world / (ms * 1e-3) / 1e9 / e.peak},
            "parity": {"stft_rel_err_vs_oracle": es, "istft_rel_err_vs_oracle": ey, "tolerance": 1e-5,
                       "recon_max_abs_err": recon,
                       "checked": "first / interior / last 24 slices of the first and last row of every rank, max over ranks"},
            "kernels": {"stft": st.kernel_name(False), "istft": st.kernel_name(True)}}


# ---- c4: ONE long signal sharded by slice range (halo of m_num - hop samples); inverse with the
#      halo-column exchange over NCCL send / recv
def side_c4(e, args):
    cfg = dict(CONFIGS["c4"])
    torch, L, sh = e.torch, e.L, e.sharding
    n = cfg["n"] if e.world == 8 else (1 << args.c4_log2_per_gpu) * e.world
    if args.c4_log2:
        n = 1 << args.c4_log2
    cfg["n"] = n
    st = plan_of(e, cfg)
    o = oracle_of(cfg)
    N, H, mid = cfg["m"], cfg["hop"], cfg["m"] // 2
    p_min, p_max = st.p_min, st.p_max(n)
    P_tot, F = p_max - p_min, st.f_pts
    fs_ = sh.frame_shard(n, N, H, p_min, p_max, e.rank, e.world)
    shards = [sh.inverse_shard(P_tot, N, H, p_min, r, e.world, halo_align=2) for r in range(e.world)]
    me = shards[e.rank]
    halo_next = shards[e.rank + 1].halo_cols if e.rank + 1 < e.world else 0
    P_r = fs_.p_b - fs_.p_a
    off = 16 if me.halo_cols else 0       # own columns start on a 128-byte line; the halo block sits right before them
    ldp = -(-(off + P_r) // 16) * 16
    free, _ = torch.cuda.mem_get_info()
    need = (fs_.s_b - fs_.s_a) * 4 + F * ldp * 8 + (me.k_b - me.k_a) * 4 + (3 << 30)
    if need > free:
        return {"skipped": f"needs {need / 2**30:.0f} GiB of device memory, {free / 2**30:.0f} free"}
    x = hash_noise_torch(torch, e.dev, fs_.s_a, fs_.s_b)                 # this rank's samples + halo, generated in place
    Sbuf = torch.empty((F, ldp), dtype=torch.complex64, device=e.dev)
    S = Sbuf[:, off - me.halo_cols:]                                     # [halo | own] columns
    y = torch.empty((me.k_b - me.k_a,), dtype=torch.float32, device=e.dev)
    stream = torch.cuda.current_stream(e.dev).cuda_stream
    n_loc = fs_.s_b - fs_.s_a
    S_own = S[:, me.halo_cols:]
    pa, pb = C_i64(fs_.p_a), C_i64(fs_.p_b)
    k0l, k1l = C_i64(me.k0_local), C_i64(me.k1_local)

    def fwd():
        rc = L.veils_stft_dev(st._h, x.data_ptr(), n_loc, 1, n_loc, pa, pb, fs_.k_offset, S_own.data_ptr(), ldp, stream)
        assert rc == 0, L.veils_last_error()

    send_buf = torch.empty((F, max(halo_next, 1)), dtype=torch.complex64, device=e.dev)
    recv_buf = torch.empty((F, max(me.halo_cols, 1)), dtype=torch.complex64, device=e.dev)

    def exchange():
        """my last halo_next columns -> rank + 1; rank - 1's last columns -> my halo block (NVLink, NCCL p2p)"""
        if e.world == 1:
            return
        ops = []
        if halo_next > 0:
            send_buf.copy_(S_own[:, P_r - halo_next:P_r])
            ops.append(e.dist.P2POp(e.dist.isend, send_buf, e.rank + 1))
        if me.halo_cols > 0:
            ops.append(e.dist.P2POp(e.dist.irecv, recv_buf, e.rank - 1))
        if ops:
            for r in e.dist.batch_isend_irecv(ops):
                r.wait()
        if me.halo_cols > 0:
            S[:, :me.halo_cols].copy_(recv_buf)

    def inv():
        rc = L.veils_istft_dev(st._h, S.data_ptr(), me.halo_cols + P_r, 1, ldp, k0l, k1l, y.data_ptr(), y.numel(), stream)
        assert rc == 0, L.veils_last_error()

    for _ in range(2):
        fwd()
        exchange()
        inv()
    barrier(e)
    steps = max(1, min(args.steps, args.c4_steps))
    ev = [torch.cuda.Event(enable_timing=True) for _ in range(4)]
    t_f = t_x = t_i = 0.0
    ev_all = [torch.cuda.Event(enable_timing=True) for _ in range(2)]
    launches0 = L.veils_kernel_launch_count()
    ev_all[0].record()
    for _ in range(steps):
        ev[0].record()
        fwd()
        ev[1].record()
        exchange()
        ev[2].record()
        inv()
        ev[3].record()
        torch.cuda.synchronize()
        t_f += ev[0].elapsed_time(ev[1])
        t_x += ev[1].elapsed_time(ev[2])
        t_i += ev[2].elapsed_time(ev[3])
    ev_all[1].record()
    barrier(e)
    launches = L.veils_kernel_launch_count() - launches0
    ms = max_over_ranks(e, ev_all[0].elapsed_time(ev_all[1])) / steps
    ms_f, ms_x, ms_i = (max_over_ranks(e, t / steps) for t in (t_f, t_x, t_i))
    # ---- parity at the shard boundaries (every rank, max over ranks)
    # (1) oracle: the first / last 12 own slices from the hash-generated samples they cover
    worst_s = worst_y = 0.0
    bit_ok = True
    for qa in sorted({fs_.p_a, max(fs_.p_a, fs_.p_b - 12)}):
        qb = min(qa + 12, fs_.p_b)
        s_a, s_b = max(0, qa * H - mid), min(n, (qb - 1) * H - mid + N)
        ref = o.stft(hash_noise_np(s_a, s_b).astype(np.float64), qa, qb, k_offset=-s_a)
        got = S_own[:, qa - fs_.p_a:qb - fs_.p_a].cpu().numpy()
        worst_s = max(worst_s, rel_err(got, ref))
    # (2) the boundary neighbourhood transformed as ONE unsharded call on this GPU: slices
    #     [p_a - 8, p_a + 8) -> the halo columns received from rank - 1 and my first columns must be
    #     the same bits, and so must my first output samples
    if e.rank > 0:
        qa, qb = fs_.p_a - 8, fs_.p_a + 8
        s_a, s_b = max(0, qa * H - mid), min(n, (qb - 1) * H - mid + N)
        xn = hash_noise_torch(torch, e.dev, s_a, s_b)
        Sn = st.stft(xn, qa, qb, k_offset=-s_a)                           # [F, 16]
        h = me.halo_cols
        bit_ok &= bool(torch.equal(Sn[:, 8 - h:8], S[:, :h])) and bool(torch.equal(Sn[:, 8:16], S_own[:, :8]))
        # istft of the 16 neighbourhood columns: the samples all of whose slices lie inside
        shift = (qa - p_min) * H
        k_lo, k_hi = me.k_a, me.k_a + 4 * H
        yn = st.istft(Sn, k_lo - shift, k_hi - shift)
        bit_ok &= bool(torch.equal(yn, y[:4 * H]))
        yo = o.istft(Sn.cpu().numpy().astype(np.complex128), k_lo - shift, k_hi - shift)
        worst_y = max(worst_y, rel_err(y[:4 * H].cpu().numpy(), yo))
    # (3) perfect reconstruction of this rank's output range
    ya, yb = max(me.k_a, 0), min(me.k_b, n)
    xr = hash_noise_torch(torch, e.dev, ya, yb)
    recon = float((y[ya - me.k_a:yb - me.k_a] - xr).abs().max().item())
    worst_s, worst_y, recon = max_over_ranks(e, worst_s), max_over_ranks(e, worst_y), max_over_ranks(e, recon)
    bit_all = max_over_ranks(e, 0.0 if bit_ok else 1.0) == 0.0
    bf, bi = algorithmic_bytes(cfg, 1, n_loc, P_r, F, y.numel())
    return {"workload": f"c4: ONE signal of 2^{int(np.log2(n))} samples (hash noise, generated per rank for its own range + halo), "
                        f"Gaussian 1024 (std 128) / hop 256, twosided, f32, stft+istft; sharded by slice range over {e.world} "
                        f"rank(s), halo {fs_.halo} samples forward, {me.halo_cols if e.rank else shards[-1].halo_cols} S columns inverse",
            "scaling": "weak below 8 GPUs (2^%d samples per GPU), BASELINE size 2^32 at 8" % args.c4_log2_per_gpu,
            "value": n / (ms * 1e-3) / 1e6, "unit": "Msamples/s", "ms_per_step": ms, "steps": steps,
            "gpu_launches": int(launches),
            "exchange": {"ms": ms_x, "bytes_per_boundary": int(F * max(s.halo_cols for s in shards) * 8),
                         "how": "NCCL send/recv (torch.distributed.batch_isend_irecv) of the halo columns + 2 small copies",
                         "share_of_step": ms_x / ms if ms > 0 else None},
            "roofline": {"bound": "hbm", "peak": e.peak, "unit": "GB/s", "stft": roof_entry(e, bf, ms_f),
                         "istft": roof_entry(e, bi, ms_i),
                         "roundtrip_frac_per_gpu": (bf + bi) / ((ms_f + ms_i) * 1e-3) / 1e9 / e.peak},
            "parity": {"stft_rel_err_vs_oracle": worst_s, "istft_rel_err_vs_oracle": worst_y, "tolerance": 1e-5,
                       "recon_max_abs_err": recon, "boundary_bit_identical_to_unsharded": bit_all,
                       "checked": "first / last 12 slices of every rank vs the oracle; halo columns, first own columns and "
                                  "first 4 hops of output of every rank bit-compared with an unsharded transform of the boundary"},
            "kernels": {"stft": st.kernel_name(False), "istft": st.kernel_name(True)}}


def C_i64(v):
    import ctypes
    return ctypes.byref(ctypes.c_int64(int(v)))


# ---- c5: high-overlap psd spectrogram in f64 (FP64-bound shape): a sub-batch
def side_c5(e, args):
    cfg = dict(CONFIGS["c5"])
    torch, L = e.torch, e.L
    st = plan_of(e, cfg)
    B, n = args.c5_rows, cfg["n"]
    g = torch.Generator(device=e.dev)
    g.manual_seed(5000 + e.rank)
    x = torch.rand((B, n), generator=g, device=e.dev, dtype=torch.float64) * 2 - 1
    P, F = st.p_num(n), st.f_pts
    ldp = -(-P // 16) * 16
    Sx = torch.empty((B, F, ldp), dtype=torch.float64, device=e.dev)
    stream = torch.cuda.current_stream(e.dev).cuda_stream

    def spec():
        assert L.veils_spectrogram_dev(st._h, x.data_ptr(), n, B, n, None, None, 0, Sx.data_ptr(), ldp, stream) == 0
    ms = max_over_ranks(e, time_events(e, spec, 3))
    o = oracle_of(cfg)
    xr = x[B - 1, :40000].cpu().numpy()
    ref = o.spectrogram(xr)                                               # leading slices depend on the prefix only
    ncol = 400
    es = rel_err(Sx[B - 1, :, :ncol].cpu().numpy(), ref[:, :ncol])
    nbytes, _ = algorithmic_bytes(cfg, B, n, P, F, 0)
    M = cfg["m"]
    flops = B * P * (2.5 * M * np.log2(M) + 4 * M)                        # real FFT of M points + window + |X|^2
    return {"workload": f"c5: {B} of 4096 rows x {n} @ 16 kHz, Hann 8192 / hop 64 (128x overlap), onesided2X, "
                        "scale_to psd, spectrogram, f64",
            "value": e.world * B * n / (ms * 1e-3) / 1e6, "unit": "Msamples/s", "ms_per_launch": ms,
            "roofline": {"bound": "fp64", "hbm": roof_entry(e, nbytes, ms),
                         "fp64": {"TFLOPs": flops / (ms * 1e-3) / 1e12, "peak": FP64_PEAK_TFLOPS,
                                  "frac": flops / (ms * 1e-3) / 1e12 / FP64_PEAK_TFLOPS,
                                  "flop_model": "2.5 M log2 M + 4 M per slice"}},
            "parity": {"spectrogram_rel_err_vs_oracle": es, "tolerance": 1e-12,
                       "checked": f"first {ncol} slices of the last row"},
            "kernels": {"spectrogram": st.kernel_name(False)}}


# ---- neighbouring shapes the round-1 review asked for next to the BASELINE configs: the c2 geometry in f64 (the
#      reference's own precision), the non-power-of-two speech front end Hann 400 / hop 160 (mixed-radix family),
#      and the 1024 / 256 and 2048 / 512 front ends (wide family): round trip on a sub-batch, roofline, oracle parity
EXTRA = {
    "f64_c2": dict(batch=256, n=160000, fs=16000.0, m=512, hop=128, mode="onesided", prec="f64", win="hann"),
    "hann400": dict(batch=1024, n=160000, fs=16000.0, m=400, hop=160, mode="onesided", prec="f32", win="hann"),
    "m1024": dict(batch=512, n=160000, fs=16000.0, m=1024, hop=256, mode="onesided", prec="f32", win="hann"),
    "m2048": dict(batch=256, n=480000, fs=48000.0, m=2048, hop=512, mode="onesided", prec="f32", win="hann"),
}


def side_extra(e, name):
    cfg = dict(EXTRA[name])
    torch, L = e.torch, e.L
    st = plan_of(e, cfg)
    x = synth_device(cfg, torch, e.dev, 7000 + e.rank)
    B, n = x.shape
    P, F = st.p_num(n), st.f_pts
    k0, k1 = st.istft_range(P)
    ldp = -(-P // st.pitch_align) * st.pitch_align
    S = torch.empty((B, F, ldp), dtype=torch.complex64 if cfg["prec"] == "f32" else torch.complex128, device=e.dev)
    y = torch.empty((B, k1 - k0), dtype=x.dtype, device=e.dev)
    stream = torch.cuda.current_stream(e.dev).cuda_stream

    def fwd():
        assert L.veils_stft_dev(st._h, x.data_ptr(), n, B, n, None, None, 0, S.data_ptr(), ldp, stream) == 0

    def inv():
        assert L.veils_istft_dev(st._h, S.data_ptr(), P, B, ldp, None, None, y.data_ptr(), k1 - k0, stream) == 0
    ms_f = max_over_ranks(e, time_events(e, fwd, 5))
    ms_i = max_over_ranks(e, time_events(e, inv, 5))
    bf, bi = algorithmic_bytes(cfg, B, n, P, F, k1 - k0)
    es, ey = parity_rows(e, cfg, st, x, S[:, :, :P], y, [0, B - 1])
    tol = 1e-5 if cfg["prec"] == "f32" else 1e-12
    return {"workload": f"{name}: {B} x {n}, {cfg['win']} {cfg['m']} / hop {cfg['hop']}, {cfg['mode']}, {cfg['prec']}, stft+istft",
            "value": e.world * B * n / ((ms_f + ms_i) * 1e-3) / 1e6, "unit": "Msamples/s",
            "roofline": {"bound": "hbm", "peak": e.peak, "unit": "GB/s", "stft": roof_entry(e, bf, ms_f),
                         "istft": roof_entry(e, bi, ms_i),
                         "roundtrip_frac": (bf + bi) / ((ms_f + ms_i) * 1e-3) / 1e9 / e.peak},
            "parity": {"stft_rel_err_vs_oracle": es, "istft_rel_err_vs_oracle": ey, "tolerance": tol,
                       "recon_max_abs_err": float((y[:, :n] - x).abs().max()),
                       "checked": "first / interior / last 24 slices of the first and last row"},
            "kernels": {"stft": st.kernel_name(False), "istft": st.kernel_name(True)}}


# =========================================================================== main
def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=50)
    ap.add_argument("--warmup", type=int, default=5)
    ap.add_argument("--impl", default="ours", choices=["ours", "reference"])
    ap.add_argument("--config", default="c2", choices=["c2", "c3"],
                    help="headline workload (the other shapes ride in `configs`; c3 as headline: a 128-row sub-batch)")
    ap.add_argument("--batch", type=int, default=0, help="override the batch size (debug)")
    ap.add_argument("--no-cpu", action="store_true", help="skip the cpu_baseline leg")
    ap.add_argument("--no-e2e", action="store_true")
    ap.add_argument("--no-sides", action="store_true", help="skip the c1 / c3 / c4 / c5 legs")
    ap.add_argument("--sides", default="c1,c3,c4,c5,f64_c2,hann400,m1024,m2048")
    ap.add_argument("--dense-pitch", action="store_true", help="S row pitch ldp = P (reference-dense)")
    ap.add_argument("--e2e-chunks", type=int, default=8, help="sub-batches per e2e step")
    ap.add_argument("--e2e-streams", type=int, default=3)
    ap.add_argument("--c3-rows", type=int, default=0, help="override the fixed c3 batch (debug)")
    ap.add_argument("--c3-sub", type=int, default=256, help="rows per streamed sub-batch")
    ap.add_argument("--c3-steps", type=int, default=2)
    ap.add_argument("--c4-log2", type=int, default=0, help="override the c4 signal length (debug)")
    ap.add_argument("--c4-log2-per-gpu", type=int, default=29)
    ap.add_argument("--c4-steps", type=int, default=3)
    ap.add_argument("--c5-rows", type=int, default=16)
    args = ap.parse_args()
    cfg = dict(CONFIGS[args.config])
    if args.batch:
        cfg["batch"] = args.batch
    elif args.config == "c3":
        cfg["batch"] = 128
    rank = int(os.environ.get("RANK", "0"))
    world = int(os.environ.get("WORLD_SIZE", "1"))
    local = int(os.environ.get("LOCAL_RANK", "0"))
    if args.impl == "reference":
        run_reference(args, cfg, rank, world)
        return

    e = gpu_env(local, world)
    torch, dist, L, dev = e.torch, e.dist, e.L, e.dev
    st = plan_of(e, cfg)
    x = synth_device(cfg, torch, dev, seed=1000 + rank)
    B, n = x.shape
    P, F = st.p_num(n), st.f_pts
    k0, k1 = st.istft_range(P)
    n_out = k1 - k0
    cdt = torch.complex64 if cfg["prec"] == "f32" else torch.complex128
    # row pitch of S: sector-aligned rows (what StandaloneSTFT.stft allocates for device tensors);
    # --dense-pitch times the reference-dense layout ldp = P instead
    ldp = P if args.dense_pitch else -(-P // st.pitch_align) * st.pitch_align
    S = torch.empty((B, F, ldp), dtype=cdt, device=dev)
    y = torch.empty((B, n_out), dtype=x.dtype, device=dev)
    stream = torch.cuda.current_stream(dev).cuda_stream

    def fwd(Sbuf=None, pitch=None):
        Sb, lp = (S, ldp) if Sbuf is None else (Sbuf, pitch)
        rc = L.veils_stft_dev(st._h, x.data_ptr(), n, B, n, None, None, 0, Sb.data_ptr(), lp, stream)
        assert rc == 0, L.veils_last_error()

    def inv(Sbuf=None, pitch=None):
        Sb, lp = (S, ldp) if Sbuf is None else (Sbuf, pitch)
        rc = L.veils_istft_dev(st._h, Sb.data_ptr(), P, B, lp, None, None, y.data_ptr(), n_out, stream)
        assert rc == 0, L.veils_last_error()

    def step():
        fwd()
        inv()

    sampler = ClockSampler(local)
    if rank == 0:
        sampler.start()
    for _ in range(max(args.warmup, 3)):
        step()
    barrier(e)
    launches0 = L.veils_kernel_launch_count()
    ev0, ev1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    m0 = sampler.mark()
    # per-kernel CUDA events INSIDE the timed region (same stream as the launches): the roofline's
    # launch durations are those of the timed steps themselves, and stft + istft add up to the step
    kev = [[torch.cuda.Event(enable_timing=True) for _ in range(3)] for _ in range(args.steps)]
    ev0.record()
    for k in range(args.steps):
        kev[k][0].record()
        fwd()
        kev[k][1].record()
        inv()
        kev[k][2].record()
    ev1.record()
    barrier(e)
    m1 = sampler.mark()
    launches = L.veils_kernel_launch_count() - launches0
    ms_step = max_over_ranks(e, ev0.elapsed_time(ev1)) / args.steps
    value = world * B * n / (ms_step * 1e-3) / 1e6

    # perfect-reconstruction error of the device round trip (north_star: reported)
    recon = float((y[:, :n] - x).abs().max().item())

    reps = max(5, args.steps)
    ms_f_alone, ms_i_alone = time_events(e, fwd, reps), time_events(e, inv, reps)
    ms_f = sum(ev[0].elapsed_time(ev[1]) for ev in kev) / args.steps
    ms_i = sum(ev[1].elapsed_time(ev[2]) for ev in kev) / args.steps
    clocks = sampler.stop(m0, m1) if rank == 0 else None
    bytes_f, bytes_i = algorithmic_bytes(cfg, B, n, P, F, n_out)
    peak, peak_src = e.peak, e.peak_src
    dom = ("stft", ms_f, bytes_f) if ms_f >= ms_i else ("istft", ms_i, bytes_i)
    traffic = None
    try:
        traffic = json.load(open(os.path.join(ROOT, "profiles", "traffic.json"))).get(f"{args.config}:{dom[0]}")
    except Exception:
        pass
    roof = {"bound": "hbm", "kernel": f"pow2 {dom[0]} ({st.kernel_name(dom[0] == 'istft')})",
            "achieved": dom[2] / (dom[1] * 1e-3) / 1e9, "peak": peak, "unit": "GB/s",
            "frac": dom[2] / (dom[1] * 1e-3) / 1e9 / peak, "traffic": traffic,
            "traffic_source": "dram__bytes_read.sum + dram__bytes_write.sum of the committed ncu --set full capture "
                              "(profiles/traffic.json), not re-measured in this run",
            "peak_source": peak_src, "bytes_per_launch": dom[2], "ms_per_launch": dom[1],
            "stft": roof_entry(e, bytes_f, ms_f), "istft": roof_entry(e, bytes_i, ms_i),
            "roundtrip_frac": (bytes_f + bytes_i) / (ms_step * 1e-3) / 1e9 / peak, "ldp": ldp,
            "timing": "per-launch CUDA events inside the timed steps (stft.ms + istft.ms = ms_per_step); "
                      "back_to_back_ms = the same kernel launched alone in a loop after the timed region",
            "back_to_back_ms": {"stft": ms_f_alone, "istft": ms_i_alone}}
    if ldp != P:                                   # the same kernels on the dense layout, for the record
        Sd = torch.empty((B, F, P), dtype=cdt, device=dev)
        ms_fd, ms_id = time_events(e, lambda: fwd(Sd, P), reps), time_events(e, lambda: inv(Sd, P), reps)
        roof["dense_pitch"] = {"ldp": P, "stft_ms": ms_fd, "istft_ms": ms_id,
                               "stft_frac": bytes_f / (ms_fd * 1e-3) / 1e9 / peak,
                               "istft_frac": bytes_i / (ms_id * 1e-3) / 1e9 / peak}
        del Sd

    # ---- e2e: host (pinned) -> device -> stft -> istft -> host (pinned), public tensor API; S resident
    e2e = None
    if not args.no_e2e:
        chunk = max(1, -(-B // args.e2e_chunks))
        xh = torch.empty((B, n), dtype=x.dtype, pin_memory=True)
        xh.copy_(x)
        yh = torch.empty((B, n_out), dtype=x.dtype, pin_memory=True)
        streams = [torch.cuda.Stream(dev) for _ in range(args.e2e_streams)]
        ydev = [torch.empty((chunk, n_out), dtype=x.dtype, device=dev) for _ in range(args.e2e_streams)]

        def e2e_step(compute=True):
            for i, a in enumerate(range(0, B, chunk)):
                b = min(B, a + chunk)
                s = streams[i % len(streams)]
                with torch.cuda.stream(s):
                    xd = xh[a:b].to(dev, non_blocking=True)
                    if compute:
                        yd = st.istft(st.stft(xd))
                    else:                          # copy-only leg: same chunks, same streams, kernels removed
                        yd = ydev[i % len(streams)][:b - a]
                    yh[a:b].copy_(yd, non_blocking=True)

        def drain():
            for s in streams:
                s.synchronize()

        def timed(compute):
            # the K steps are issued back to back (a streaming caller does not stop the pipeline between batches: the
            # uploads of step i + 1 overlap the downloads of step i); every step still uploads its own inputs and
            # downloads its own result, and the clock stops only when the last download has landed
            for _ in range(2):
                e2e_step(compute)
            drain()
            barrier(e)
            k = max(3, min(args.steps, 10))
            t0 = time.perf_counter()
            for _ in range(k):
                e2e_step(compute)
            drain()
            barrier(e)
            return max_over_ranks(e, (time.perf_counter() - t0) / k), k
        dt, k = timed(True)
        err_e2e = float((yh[:, :n] - xh).abs().max().item())
        dt_copy, _ = timed(False)
        e2e = {"value": world * B * n / dt / 1e6, "unit": "Msamples/s",
               "h2d_bytes_per_step": int(xh.numel() * xh.element_size()),
               "d2h_bytes_per_step": int(yh.numel() * yh.element_size()),
               "ms_per_step": dt * 1e3, "steps": k,
               "what": "S-resident pipeline through the public tensor API: x in pinned host memory -> H2D -> stft -> "
                       f"istft -> D2H of the reconstructed signal; S never leaves HBM; {args.e2e_chunks} sub-batches on "
                       f"{args.e2e_streams} streams; the timed steps are issued back to back and the clock stops when "
                       "the last download has landed",
               "copy_only": {"value": world * B * n / dt_copy / 1e6, "ms_per_step": dt_copy * 1e3,
                             "what": "the same chunks / streams / pinned buffers with the kernels removed: the PCIe + "
                                     "host-memory ceiling of this box for these transfers"},
               "fraction_of_copy_only": dt_copy / dt,
               "recon_max_abs_err": err_e2e}
        # the same round trip through the HOST-pointer flavour of the C ABI (the drop-in for the
        # reference's Vec API: veils_stft_host returns S to the host, veils_istft_host takes it back ->
        # S crosses PCIe twice), on a bounded sample of rows, for the record
        if rank == 0:
            rows_h = min(B, 256)
            host_abi = {}
            for kind in ("pinned", "pageable"):
                pin = kind == "pinned"
                xs = torch.empty((rows_h, n), dtype=xh.dtype, pin_memory=pin)
                xs.copy_(xh[:rows_h])
                Sh = torch.empty((rows_h, F, P), dtype=cdt, pin_memory=pin)
                y2 = torch.empty((rows_h, n_out), dtype=xh.dtype, pin_memory=pin)
                best = None
                for _ in range(3):
                    t0 = time.perf_counter()
                    rc = L.veils_stft_host(st._h, xs.data_ptr(), n, rows_h, n, None, None, 0, Sh.data_ptr())
                    assert rc == 0, L.veils_last_error()
                    rc = L.veils_istft_host(st._h, Sh.data_ptr(), P, rows_h, None, None, y2.data_ptr())
                    assert rc == 0, L.veils_last_error()
                    dth = time.perf_counter() - t0
                    best = dth if best is None else min(best, dth)
                host_abi[kind] = rows_h * n / best / 1e6
                err_h = float((y2[:, :n] - xs).abs().max().item())
                del xs, Sh, y2
            e2e["host_abi"] = {"value": host_abi["pinned"], "pageable": host_abi["pageable"], "unit": "Msamples/s",
                               "rows": rows_h, "recon_max_abs_err": err_h,
                               "note": "veils_stft_host + veils_istft_host (chunked two-stream pipeline inside the "
                                       "library): x in, S out, S in, y out -- S, 16x the signal bytes, crosses PCIe twice"}
        del xh, yh, ydev

    # ---- parity of the headline buffers against the oracle: every rank, max over ranks
    es, ey = parity_rows(e, cfg, st, x, S, y, [0, B - 1])
    parity = {"stft_rel_err_vs_oracle": max_over_ranks(e, es), "istft_rel_err_vs_oracle": max_over_ranks(e, ey),
              "tolerance": 1e-5 if cfg["prec"] == "f32" else 1e-12,
              "recon_max_abs_err": max_over_ranks(e, recon),
              "checked": "first / interior / last 24 slices of the first and last row of every rank's S and y, max over ranks"}

    # ---- cpu_baseline: the reference's CPU path on a bounded sample (rank 0, N = 1 only)
    cpu = None
    if rank == 0 and world == 1 and not args.no_cpu:
        threads = host_threads()
        rows = max(threads, 8)
        _, _, _, t1 = cpu_reference_run(cfg, rows, threads, reps=2)
        rows = int(max(rows, min(B, rows * 12.0 / max(t1, 1e-3))))
        reps_c = int(max(1, min(10, round(10.0 / max(t1 * rows / max(threads, 8), 1e-3)))))
        xs, Sref, yref, dt = cpu_reference_run(cfg, rows, threads, seed=77, reps=reps_c)
        cpu = {"value": rows * n / dt / 1e6, "unit": "Msamples/s", "cores": threads, "kind": "port",
               "sample": f"{rows} of {B} rows, stft+istft in f64, C restatement of src/lib.rs "
                         f"(oracle/stft_ref.c, OpenMP over rows, {threads} threads set explicitly), best of {reps_c}",
               "seconds": dt, "context_msamples_per_s": cpu_context(cfg),
               "host": {"nproc": os.cpu_count(), "threads_used": threads, "cpu": cpu_model()}}
        # the CUDA path on the same rows against that CPU result
        r8 = min(8, rows)
        xd = torch.from_numpy(xs[:r8]).to(dev, x.dtype)
        Sd = st.stft(xd)
        torch.cuda.synchronize()
        parity["stft_rel_err_vs_cpu_port"] = rel_err(Sd.cpu().numpy(), Sref[:r8].astype(np.complex64).astype(np.complex128))
        del xs, Sref, yref

    del S, y, x
    torch.cuda.empty_cache()
    sides = {}
    if not args.no_sides:
        for name in args.sides.split(","):
            fn = {"c1": lambda: side_c1(e), "c3": lambda: side_c3(e, args), "c4": lambda: side_c4(e, args),
                  "c5": lambda: side_c5(e, args)}.get(name.strip())
            if fn is None and name.strip() in EXTRA:
                fn = (lambda nm: lambda: side_extra(e, nm))(name.strip())
            if fn is None:
                continue
            try:
                sides[name.strip()] = fn()
            except Exception as ex:                 # noqa: BLE001  (a side leg must not lose the headline)
                sides[name.strip()] = {"error": f"{type(ex).__name__}: {str(ex)[:300]}"}
            torch.cuda.empty_cache()
            barrier(e)

    if rank == 0:
        out = {"metric": "stft_istft_throughput", "value": value, "unit": "Msamples/s",
               "n_gpus": world, "steps": args.steps, "warmup": max(args.warmup, 3),
               "ms_per_step": ms_step, "higher_is_better": True, "scaling": "weak",
               "vs_baseline": None, "dtype": cfg["prec"], "data": "synthetic",
               "config": workload_config(cfg, args.config), "clocks": clocks, "e2e": e2e,
               "gpu_launches": int(launches), "roofline": roof, "cpu_baseline": cpu,
               "parity": parity, "recon_max_abs_err": recon,
               "kernels": {"stft": st.kernel_name(False), "istft": st.kernel_name(True)},
               "configs": sides}
        print(json.dumps(out))
    if world > 1:
        dist.destroy_process_group()


if __name__ == "__main__":
    main()
