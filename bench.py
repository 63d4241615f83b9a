#!/usr/bin/env python
"""bench.py -- BASELINE.json's metric: STFT+iSTFT Msamples/s (input samples of the batch per
second for one forward + one inverse transform), % of HBM roofline, next to the reference's CPU
path timed on this box's host cores.

    python bench.py [--gpus N] [--steps K] [--warmup W] [--impl ours|reference] [--config c2]

Workload at N=1 is BASELINE.json configs[1] ("c2"): batch 1024 x 10 s of 16 kHz speech-shaped
noise, periodic Hann 512 / hop 128, onesided, f32.  A step = one stft + one istft over the whole
batch through the C ABI (veils_stft_dev / veils_istft_dev), inputs resident in HBM.  With N > 1
(torchrun, one rank per GPU) every rank owns its own batch of the same shape (rows are
independent -> no data-path collective, "weak" scaling); time is the max over ranks.
One JSON line is printed by rank 0.
"""
import argparse
import json
import os
import subprocess
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
    "c3": dict(batch=512, n=1440000, fs=48000.0, m=4096, hop=1024, mode="onesided", prec="f32", win="hann"),
    "c4": dict(batch=1, n=1 << 28, fs=48000.0, m=1024, hop=256, mode="twosided", prec="f32", win="gauss"),
}


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


def algorithmic_bytes(cfg, P, F, n_out):
    """forward B(n sz + F P 2 sz), inverse B(F P 2 sz + n_out sz)."""
    sz = 4 if cfg["prec"] == "f32" else 8
    B, n = cfg["batch"], cfg["n"]
    return B * (n * sz + F * P * 2 * sz), B * (F * P * 2 * sz + n_out * sz)


def cpu_reference_run(cfg, rows, threads, seed=1234, reps=1):
    """Round trip of `rows` signals with the C restatement of the reference (oracle/stft_ref.c)."""
    from oracle.ref_c import RefC, num_threads
    from oracle.stft_oracle import OracleSTFT
    x = synth_host(cfg, rows, seed)
    o = OracleSTFT(make_window(cfg), cfg["hop"], cfg["fs"], fft_mode=cfg["mode"])
    c = RefC(o)
    best = None
    for _ in range(reps):
        t0 = time.perf_counter()
        S = c.stft(x, threads=threads)
        y = c.istft(S, threads=threads)
        dt = time.perf_counter() - t0
        best = dt if best is None else min(best, dt)
    return x, S, y, best, (threads or num_threads())


def run_reference(args, cfg, rank, world):
    """--impl reference: the reference's CPU path (C restatement; the Rust crate cannot be built
    in this image) on all host threads, bounded sample per step.  Rank 0 only."""
    if rank != 0:
        return
    from oracle.ref_c import num_threads
    threads = num_threads()
    rows = max(threads, 8)
    # size the sample so one step takes ~1-2 s
    _, _, _, t1, _ = cpu_reference_run(cfg, rows, threads, reps=2)
    rows = int(max(rows, min(cfg["batch"], rows * 2.0 / max(t1, 1e-3))))
    times = []
    for i in range(args.warmup + args.steps):
        _, _, _, dt, _ = cpu_reference_run(cfg, rows, threads, seed=100 + i)
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
                         "sample": sample},
        "e2e": {"value": val, "unit": "Msamples/s", "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
        "host": {"nproc": os.cpu_count(), "cpu": cpu_model()}}))


def cpu_model():
    try:
        for line in open("/proc/cpuinfo"):
            if line.startswith("model name"):
                return line.split(":", 1)[1].strip()
    except Exception:
        pass
    return "unknown"


def workload_config(cfg, name):
    return {"workload": f"{name}: batch {cfg['batch']} x {cfg['n']} samples @ {cfg['fs']:.0f} Hz, "
                        f"{cfg['win']} {cfg['m']} / hop {cfg['hop']}, {cfg['mode']}, {cfg['prec']}, stft+istft",
            "batch": cfg["batch"], "n": cfg["n"], "m_num": cfg["m"], "hop": cfg["hop"],
            "fft_mode": cfg["mode"], "l2_policy": "working set (input + S) >> 126 MB L2, no flush needed",
            "s_row_pitch": "rows padded to 128-byte lines (ldp = ceil(P/16)*16 for f32); --dense-pitch for ldp = P"}


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=50)
    ap.add_argument("--warmup", type=int, default=5)
    ap.add_argument("--impl", default="ours", choices=["ours", "reference"])
    ap.add_argument("--config", default="c2", choices=sorted(CONFIGS))
    ap.add_argument("--batch", type=int, default=0, help="override the batch size (debug)")
    ap.add_argument("--no-cpu", action="store_true", help="skip the cpu_baseline leg")
    ap.add_argument("--no-e2e", action="store_true")
    ap.add_argument("--dense-pitch", action="store_true", help="S row pitch ldp = P (reference-dense)")
    ap.add_argument("--e2e-chunks", type=int, default=8, help="sub-batches per e2e step")
    ap.add_argument("--e2e-streams", type=int, default=3)
    args = ap.parse_args()
    cfg = dict(CONFIGS[args.config])
    if args.batch:
        cfg["batch"] = args.batch
    rank = int(os.environ.get("RANK", "0"))
    world = int(os.environ.get("WORLD_SIZE", "1"))
    local = int(os.environ.get("LOCAL_RANK", "0"))
    if args.impl == "reference":
        run_reference(args, cfg, rank, world)
        return

    import torch
    import torch.distributed as dist
    from veils_b200 import StandaloneSTFT, lib

    assert torch.cuda.is_available(), "bench.py needs a CUDA device (no CPU fallback)"
    torch.cuda.set_device(local)
    dev = torch.device("cuda", local)
    if world > 1:
        os.environ.setdefault("MASTER_ADDR", "127.0.0.1")
        dist.init_process_group("nccl", device_id=dev)

    def barrier():
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize()

    L = lib()
    st = StandaloneSTFT(make_window(cfg), cfg["hop"], cfg["fs"], fft_mode=cfg["mode"],
                        precision=cfg["prec"], device=local)
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
    barrier()
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
    barrier()
    m1 = sampler.mark()
    launches = L.veils_kernel_launch_count() - launches0
    ms_total = ev0.elapsed_time(ev1)
    t = torch.tensor([ms_total], device=dev, dtype=torch.float64)
    if world > 1:
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
    ms_step = float(t.item()) / args.steps
    value = world * B * n / (ms_step * 1e-3) / 1e6

    # perfect-reconstruction error of the device round trip (north_star: reported)
    recon = float((y[:, :n] - x).abs().max().item())

    # per-kernel durations for the roofline (same stream, CUDA events, after the timed region)
    def time_kernel(fn, reps):
        a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        fn()
        torch.cuda.synchronize()
        a.record()
        for _ in range(reps):
            fn()
        b.record()
        torch.cuda.synchronize()
        return a.elapsed_time(b) / reps
    reps = max(5, args.steps)
    ms_f_alone, ms_i_alone = time_kernel(fwd, reps), time_kernel(inv, reps)
    ms_f = sum(e[0].elapsed_time(e[1]) for e in kev) / args.steps
    ms_i = sum(e[1].elapsed_time(e[2]) for e in kev) / args.steps
    clocks = sampler.stop(m0, m1) if rank == 0 else None
    bytes_f, bytes_i = algorithmic_bytes(cfg, P, F, n_out)
    peak, peak_src = measured_peak()
    dom = ("stft", ms_f, bytes_f) if ms_f >= ms_i else ("istft", ms_i, bytes_i)
    traffic = None
    try:
        traffic = json.load(open(os.path.join(ROOT, "profiles", "traffic.json"))).get(
            f"{args.config}:{dom[0]}")
    except Exception:
        pass
    roof = {"bound": "hbm", "kernel": f"pow2 {dom[0]} ({st.kernel_name(dom[0] == 'istft')})",
            "achieved": dom[2] / (dom[1] * 1e-3) / 1e9, "peak": peak, "unit": "GB/s",
            "frac": dom[2] / (dom[1] * 1e-3) / 1e9 / peak, "traffic": traffic,
            "peak_source": peak_src, "bytes_per_launch": dom[2], "ms_per_launch": dom[1],
            "stft": {"ms": ms_f, "GBps": bytes_f / (ms_f * 1e-3) / 1e9, "frac": bytes_f / (ms_f * 1e-3) / 1e9 / peak},
            "istft": {"ms": ms_i, "GBps": bytes_i / (ms_i * 1e-3) / 1e9, "frac": bytes_i / (ms_i * 1e-3) / 1e9 / peak},
            "roundtrip_frac": (bytes_f + bytes_i) / (ms_step * 1e-3) / 1e9 / peak, "ldp": ldp,
            "timing": "per-launch CUDA events inside the timed steps (stft.ms + istft.ms = ms_per_step); "
                      "back_to_back_ms = the same kernel launched alone in a loop after the timed region",
            "back_to_back_ms": {"stft": ms_f_alone, "istft": ms_i_alone}}
    if ldp != P:                                   # the same kernels on the dense layout, for the record
        Sd = torch.empty((B, F, P), dtype=cdt, device=dev)
        ms_fd, ms_id = time_kernel(lambda: fwd(Sd, P), reps), time_kernel(lambda: inv(Sd, P), reps)
        roof["dense_pitch"] = {"ldp": P, "stft_ms": ms_fd, "istft_ms": ms_id,
                               "stft_frac": bytes_f / (ms_fd * 1e-3) / 1e9 / peak,
                               "istft_frac": bytes_i / (ms_id * 1e-3) / 1e9 / peak}
        del Sd

    # ---- e2e: host (pinned) -> device -> stft -> istft -> host (pinned), public tensor API ----
    e2e = None
    if not args.no_e2e:
        chunk = max(1, -(-B // args.e2e_chunks))
        xh = torch.empty((B, n), dtype=x.dtype, pin_memory=True)
        xh.copy_(x)
        yh = torch.empty((B, n_out), dtype=x.dtype, pin_memory=True)
        streams = [torch.cuda.Stream(dev) for _ in range(args.e2e_streams)]

        def e2e_step():
            for i, a in enumerate(range(0, B, chunk)):
                b = min(B, a + chunk)
                s = streams[i % len(streams)]
                with torch.cuda.stream(s):
                    xd = xh[a:b].to(dev, non_blocking=True)
                    Sd = st.stft(xd)
                    yd = st.istft(Sd)
                    yh[a:b].copy_(yd, non_blocking=True)
            for s in streams:
                s.synchronize()
        for _ in range(2):
            e2e_step()
        barrier()
        k = max(3, min(args.steps, 10))
        t0 = time.perf_counter()
        for _ in range(k):
            e2e_step()
        barrier()
        dt = (time.perf_counter() - t0) / k
        te = torch.tensor([dt], device=dev, dtype=torch.float64)
        if world > 1:
            dist.all_reduce(te, op=dist.ReduceOp.MAX)
        dt = float(te.item())
        e2e = {"value": world * B * n / dt / 1e6, "unit": "Msamples/s",
               "h2d_bytes_per_step": int(xh.numel() * xh.element_size()),
               "d2h_bytes_per_step": int(yh.numel() * yh.element_size()),
               "ms_per_step": dt * 1e3, "steps": k,
               "note": "x pinned host -> H2D -> stft -> istft -> D2H reconstructed signal; "
                       f"S stays in HBM; {args.e2e_chunks} sub-batches on {args.e2e_streams} streams"}
        e2e["recon_max_abs_err"] = float((yh[:, :n] - xh).abs().max().item())
        # the same round trip through the HOST-pointer flavour of the C ABI (the drop-in for the
        # reference's Vec API: veils_stft_host returns S to the host, veils_istft_host takes it back ->
        # S crosses PCIe twice), on a bounded sample of rows, for the record
        if rank == 0:
            rows_h = min(B, 256)
            cdt = torch.complex64 if cfg["prec"] == "f32" else torch.complex128
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

    # ---- cpu_baseline + parity of a sample against the reference's CPU path (rank 0, N = 1) ----
    cpu = None
    parity = None
    if rank == 0 and world == 1 and not args.no_cpu:
        from oracle.ref_c import num_threads
        threads = num_threads()
        rows = max(threads, 8)
        _, _, _, t1, _ = cpu_reference_run(cfg, rows, threads, reps=2)
        rows = int(max(rows, min(B, rows * 12.0 / max(t1, 1e-3))))
        reps = int(max(1, min(10, round(10.0 / max(t1 * rows / max(threads, 8), 1e-3)))))
        xs, Sref, yref, dt, threads = cpu_reference_run(cfg, rows, threads, seed=77, reps=reps)
        cpu = {"value": rows * n / dt / 1e6, "unit": "Msamples/s", "cores": threads, "kind": "port",
               "sample": f"{rows} of {B} rows, stft+istft in f64, C restatement of src/lib.rs "
                         f"(oracle/stft_ref.c, OpenMP over rows), best of {reps}", "seconds": dt,
               "host": {"nproc": os.cpu_count(), "cpu": cpu_model()}}
        # parity of the CUDA path on the same rows (first 8) against that CPU result
        r8 = min(8, rows)
        xd = torch.from_numpy(xs[:r8]).to(dev, x.dtype)
        Sd = st.stft(xd)
        yd = st.istft(Sd)
        torch.cuda.synchronize()
        Sn = np.stack([OracleShim.stft_cast(xd[i].cpu().numpy(), cfg) for i in range(r8)])
        err_s = float(np.max(np.abs(Sd.cpu().numpy() - Sn)) / np.max(np.abs(Sn)))
        yn = yref[:r8]
        err_y = float(np.max(np.abs(yd.cpu().numpy()[:, :n] - xs[:r8])))
        parity = {"stft_rel_err_vs_cpu": err_s, "roundtrip_max_abs_err": err_y,
                  "tolerance": 1e-5 if cfg["prec"] == "f32" else 1e-12, "rows": r8}
        del yn

    if rank == 0:
        out = {"metric": "stft_istft_throughput", "value": value, "unit": "Msamples/s",
               "n_gpus": world, "steps": args.steps, "warmup": max(args.warmup, 3),
               "ms_per_step": ms_step, "higher_is_better": True, "scaling": "weak",
               "vs_baseline": None, "dtype": cfg["prec"], "data": "synthetic",
               "config": workload_config(cfg, args.config), "clocks": clocks, "e2e": e2e,
               "gpu_launches": int(launches), "roofline": roof, "cpu_baseline": cpu,
               "parity": parity, "recon_max_abs_err": recon,
               "kernels": {"stft": st.kernel_name(False), "istft": st.kernel_name(True)}}
        print(json.dumps(out))
    if world > 1:
        dist.destroy_process_group()


class OracleShim:
    """cpu_baseline leg only: numpy oracle on the exact (precision-rounded) device input."""

    @staticmethod
    def stft_cast(xrow, cfg):
        from oracle.stft_oracle import OracleSTFT
        o = OracleSTFT(make_window(cfg), cfg["hop"], cfg["fs"], fft_mode=cfg["mode"])
        return o.stft(xrow.astype(np.float64))


if __name__ == "__main__":
    main()
