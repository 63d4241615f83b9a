"""The comparison-helper CLI (veils_b200/csrc/cli/comparison_helper.cpp): the protocol of the
reference's src/bin/python_rust_comparison_helper.rs (:14-54), replayed the way
python/final_comparison.py does (:39-41 subprocess, :250-260 configs, tolerance 1e-12 at :53,86)."""
import json
import os
import subprocess
import sys

import numpy as np
import pytest

sys.path.insert(0, os.path.dirname(os.path.abspath(__file__)))
import golden_util as G  # noqa: E402

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
CLI = os.path.join(ROOT, "veils_b200", "veils_comparison_helper")


@pytest.fixture(scope="module")
def cli():
    from veils_b200 import build
    build.build()
    assert os.path.exists(CLI)
    return CLI


def run(cli, tmp_path, obj=None, text=None):
    p = tmp_path / "in.json"
    p.write_text(text if text is not None else json.dumps(obj))
    return subprocess.run([cli, str(p)], capture_output=True, text=True, timeout=120)


def test_usage(cli):
    r = subprocess.run([cli], capture_output=True, text=True)
    assert r.returncode == 1 and "Usage:" in r.stderr and "<input_json_file>" in r.stderr


def test_missing_file_and_bad_json(cli, tmp_path):
    r = subprocess.run([cli, str(tmp_path / "nope.json")], capture_output=True, text=True)
    assert r.returncode == 1 and r.stderr.startswith("Error:")
    r = run(cli, tmp_path, text='{"signal": [1, 2, "window": [1]}')
    assert r.returncode == 1 and r.stderr.startswith("Error:")
    r = run(cli, tmp_path, obj={"signal": [1.0, 2.0], "window": [1.0, 1.0], "fs": 1.0, "fft_mode": "onesided"})
    assert r.returncode == 1 and "missing field `hop_length`" in r.stderr


def test_reference_error_texts_need_no_gpu(cli, tmp_path):
    """Constructor validation happens before any CUDA call: the reference's messages come through."""
    base = {"signal": [0.0] * 32, "window": [1.0] * 8, "hop_length": 2, "fs": 1.0, "fft_mode": "onesided"}
    r = run(cli, tmp_path, obj=dict(base, fft_mode="sideways"))
    assert r.returncode == 1 and "Unknown fft_mode: sideways" in r.stderr           # src/lib.rs:75
    r = run(cli, tmp_path, obj=dict(base, window=[]))
    assert r.returncode == 1 and "Parameter win must be non-empty!" in r.stderr      # src/lib.rs:176-178
    r = run(cli, tmp_path, obj=dict(base, hop_length=0))
    assert r.returncode == 1 and "Parameter hop=0 is not >= 1!" in r.stderr          # src/lib.rs:183-185


def test_fails_loudly_without_gpu(cli, tmp_path):
    import torch
    if torch.cuda.is_available():
        pytest.skip("a GPU is present")
    r = run(cli, tmp_path, obj={"signal": [0.0] * 32, "window": [1.0] * 8, "hop_length": 2, "fs": 1.0,
                                 "fft_mode": "onesided"})
    assert r.returncode == 1 and r.stderr.startswith("Error:") and r.stdout == ""


@pytest.mark.gpu
@pytest.mark.parametrize("fixture", ["final_comparison", "integration", "basic_usage"])
def test_cli_matches_twin_golden(cli, tmp_path, fixture):
    for rec in G.load(fixture):
        if int(rec["mfft"]) != len(rec["win"]) or int(rec["phase_shift"]) != 0 or int(rec["k_offset"]) != 0:
            continue                           # the helper only passes win / hop / fs / fft_mode
        if G.opt(rec["p0"]) is not None or G.opt(rec["p1"]) is not None or G.opt(rec["k1"]) is not None:
            continue
        r = run(cli, tmp_path, obj={"signal": rec["x"].tolist(), "window": rec["win"].tolist(),
                                     "hop_length": int(rec["hop"]), "fs": float(rec["fs"]),
                                     "fft_mode": str(rec["mode"])})
        assert r.returncode == 0, r.stderr
        out = json.loads(r.stdout)
        S = np.array([[c["real"] + 1j * c["imag"] for c in row] for row in out["stft"]])
        assert S.shape == rec["S"].shape
        assert G.rel_err(S, rec["S"]) <= 1e-12, rec["name"]
        y = np.array(out["istft"])
        yref = np.real(rec["y"])
        assert y.shape == yref.shape and G.rel_err(y, yref) <= 1e-12, rec["name"]
        pr = out["properties"]
        assert (pr["m_num"], pr["f_pts"], pr["p_min"], pr["p_max"], pr["mfft"], pr["hop"]) == (
            len(rec["win"]), int(rec["f_pts"]), int(rec["p_min"]), int(rec["p_max"]), int(rec["mfft"]), int(rec["hop"]))
        assert pr["fs"] == float(rec["fs"])
