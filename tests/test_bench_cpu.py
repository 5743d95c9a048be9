"""Host-side pieces of bench.py that need no GPU: the default workload, the algorithmic byte count
of SURVEY.md 8(d), the nnz-balanced shard plan, and the staleness rule of the measured traffic."""
import json
import os
import subprocess
import sys

import numpy as np
import torch

import bench
from conftest import ROOT


def test_default_workload_is_the_north_star_configuration():
    out = subprocess.run([sys.executable, os.path.join(ROOT, "bench.py"), "--help"],
                         capture_output=True, text=True).stdout
    assert "--workload" in out
    ap_default = [ln for ln in open(os.path.join(ROOT, "bench.py")) if '"--workload"' in ln][0]
    assert 'default="c3"' in ap_default
    wl = bench.WORKLOADS["c3"]
    assert wl["n"] == 20000 and wl["m_total"] == 1300000 and wl["rank"] == 20
    assert "m_total" in wl and "m_per_gpu" not in wl          # strong scaling


def test_algorithmic_bytes_match_the_survey_figures():
    # C2 ~ 1.30 GB, C3 ~ 17.07 GB per iteration (fp64)
    assert abs(bench.algorithmic_bytes(1.6e8, 20000, 100000, 10) / 1.30e9 - 1) < 0.01
    assert abs(bench.algorithmic_bytes(2.08e9, 20000, 1300000, 20) / 17.07e9 - 1) < 0.01


def test_shard_plan_is_chunk_aligned_balanced_and_identical_everywhere():
    from ccfindr_b200 import synth
    wl = dict(bench.WORKLOADS["smallstrong"])
    m_total = wl["m_total"]
    dev = torch.device("cpu")
    b4, how = bench.shard_plan(wl, m_total, 4, dev)
    assert b4 == bench.shard_plan(wl, m_total, 4, dev)[0]      # deterministic: same on every rank
    assert b4[0] == 0 and b4[-1] == m_total and "nnz-balanced" in how
    assert all(b % synth.TENX_CHUNK == 0 for b in b4[:-1]) and all(np.diff(b4) > 0)
    exp = synth.tenx_expected_nnz(wl["n"], m_total, wl["r_true"], wl["density"], wl["seed"], dev)
    per = [exp[b4[i] // synth.TENX_CHUNK:-(-b4[i + 1] // synth.TENX_CHUNK)].sum() for i in range(4)]
    assert max(per) / min(per) < 1.0 + 2.0 / len(exp) * 4 + 0.05   # within about one chunk
    assert bench.shard_plan(wl, m_total, 1, dev)[0] == [0, m_total]


def test_measured_traffic_is_refused_when_stale(tmp_path, monkeypatch):
    from ccfindr_b200 import build as vb_build
    prof = tmp_path / "profiles"
    prof.mkdir()
    monkeypatch.setattr(bench, "ROOT", str(tmp_path))
    assert bench.measured_traffic("c3")[0] is None                       # no capture
    (prof / "traffic.json").write_text(json.dumps({"source_hash": "0" * 16, "c3": 123}))
    val, note = bench.measured_traffic("c3")
    assert val is None and "stale" in note
    (prof / "traffic.json").write_text(json.dumps({"source_hash": vb_build.kernel_hash(), "c3": 123,
                                                   "note": "n"}))
    assert bench.measured_traffic("c3") == (123, "n")
