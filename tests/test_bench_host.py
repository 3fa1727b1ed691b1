"""Host-side pieces of bench.py that need no GPU: the workload table against BASELINE.json, the
synthetic slabs of the multi-GPU run, the roofline constants, and the reference arm's JSON contract
(with the timed CPU solve stubbed out -- the real one takes tens of seconds by design)."""

import io
import json
import os
import sys
from contextlib import redirect_stdout
from types import SimpleNamespace

import numpy as np
from conftest import ROOT

sys.path.insert(0, ROOT)
import bench  # noqa: E402


def test_workloads_follow_baseline_configs():
    with open(os.path.join(ROOT, "BASELINE.json")) as f:
        base = json.load(f)
    assert len(base["configs"]) == 5 and sorted(bench.WORKLOADS) == ["cfg1", "cfg2", "cfg3", "cfg4", "cfg5"]
    text = " ".join(base["configs"])
    for needle in ("1024x1024", "4096", "8192", "32768", "256"):
        assert needle in text
    w = bench.WORKLOADS
    assert (w["cfg1"]["solver"], w["cfg1"]["size"], w["cfg1"]["mask"]) == ("equ", 1026, "square")  # 1024^2 unknowns
    assert (w["cfg2"]["solver"], w["cfg2"]["size"], w["cfg2"]["mask"]) == ("grid", 4096, "circle")
    assert (w["cfg3"]["solver"], w["cfg3"]["size"]) == ("equ", 8192)
    assert (w["cfg4"]["solver"], w["cfg4"]["size"], w["cfg4"]["mask"]) == ("grid", 32768, "square")
    assert (w["cfg5"]["solver"], w["cfg5"]["size"], w["cfg5"]["batch"]) == ("batch", 256, 512)
    # algorithmic bytes per unknown and sweep
    assert bench.GRID_BYTES_PER_UPDATE == 36 and bench.EQU_BYTES_PER_UPDATE == 52


def test_measured_peak_prefers_the_driver_file(tmp_path, monkeypatch):
    monkeypatch.setattr(bench, "ROOT", str(tmp_path))
    peak, src = bench.measured_peak()
    assert peak == bench.FALLBACK_HBM_GBS and "fallback" in src
    (tmp_path / "MEASURED_PEAKS.json").write_text(json.dumps({"hbm_gbs": 6559.4}))
    peak, src = bench.measured_peak()
    assert peak == 6559.4 and "measured" in src


def test_slab_images_are_cuts_of_one_global_problem():
    """Every rank's slab is a row range of ONE global image pair: rows two slabs share (the halos) are
    identical, so the sharded solve is the single-GPU solve and can be checked against it."""
    from fpie_b200 import band, synth

    n, m, world, halo = 2500, 64, 3, 24  # (spans several generator chunks of bench.CHUNK_ROWS rows)
    fsrc, fmask, ftgt, unknowns = bench.slab_images(dict(mask="square"), 0, n, n, m)
    assert fsrc.shape == ftgt.shape == (n, m, 3) and fmask.shape == (n, m) and unknowns == (n - 2) * (m - 2)
    assert not np.array_equal(fsrc, ftgt)
    assert int((fmask > 127).sum()) == unknowns and not fmask[0].any() and not fmask[:, 0].any()
    total = 0
    for rank in range(world):
        plan = band.make_plan(n, world, rank, halo)
        src, mask, tgt, unk = bench.slab_images(dict(mask="square"), plan.slab_lo, plan.slab_hi, n, m)
        assert src.dtype == tgt.dtype == mask.dtype == np.uint8 and unk == unknowns
        np.testing.assert_array_equal(src, fsrc[plan.slab_lo : plan.slab_hi])
        np.testing.assert_array_equal(tgt, ftgt[plan.slab_lo : plan.slab_hi])
        np.testing.assert_array_equal(mask, fmask[plan.slab_lo : plan.slab_hi])
        total += plan.band_hi - plan.band_lo
    assert total == n
    plan = band.make_plan(n, world, 1, halo)
    _, mask, _, unknowns = bench.slab_images(dict(mask="circle"), plan.slab_lo, plan.slab_hi, n, m)
    full = synth.make_mask("circle", n, m)
    full[0] = full[-1] = 0
    full[:, 0] = full[:, -1] = 0
    np.testing.assert_array_equal(mask, full[plan.slab_lo : plan.slab_hi])
    assert unknowns == int((full > 127).sum())


def test_reference_arm_prints_the_contract_line(monkeypatch):
    calls = {}

    def fake_cpu_baseline(work, src, mask, tgt, budget_s=12.0, steps=1):
        calls["work"], calls["steps"] = work, steps
        info = {"value": 0.0, "unit": "Gupd/s", "cores": 4, "kind": "reference", "sample": "stub"}
        return info, [2.0] * steps, 100, 1000

    monkeypatch.setattr(bench, "cpu_baseline", fake_cpu_baseline)
    monkeypatch.setitem(bench.WORKLOADS, "cfg2", dict(solver="grid", size=64, mask="circle", grad="max", iters=5000))
    args = SimpleNamespace(steps=3, warmup=2, gpus=1)
    buf = io.StringIO()
    with redirect_stdout(buf):
        bench.run_reference(args, bench.WORKLOADS["cfg2"], "cfg2")
    lines = [ln for ln in buf.getvalue().splitlines() if ln.strip()]
    assert len(lines) == 1
    line = json.loads(lines[0])
    assert line["impl"] == "reference" and line["metric"] == "jacobi_gupd_per_s" and line["unit"] == "Gupd/s"
    assert (line["n_gpus"], line["steps"], line["warmup"], line["higher_is_better"]) == (1, 3, 2, True)
    assert calls["steps"] == 5  # warm-up + timed steps are all run; only the timed ones are averaged
    assert line["ms_per_step"] == 2000.0 and abs(line["value"] - 1000 * 100 / 2.0 / 1e9) < 1e-15
    assert line["e2e"] == {"value": line["value"], "unit": "Gupd/s", "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0}
    assert line["cpu_baseline"]["kind"] == "reference" and line["cpu_baseline"]["value"] == line["value"]
    assert line["vs_baseline"] is None and line["gpu_launches"] == 0 and "workload" in line["config"]


def test_reference_arm_is_silent_on_other_ranks(monkeypatch):
    monkeypatch.setenv("RANK", "1")
    buf = io.StringIO()
    with redirect_stdout(buf):
        bench.run_reference(SimpleNamespace(steps=1, warmup=0, gpus=2), bench.WORKLOADS["cfg4"], "cfg4")
    assert buf.getvalue() == ""


def test_multi_gpu_line_measures_its_own_reference_and_parity():
    """The N > 1 line carries a single-GPU figure and a sharded-vs-single parity verdict MEASURED in the run,
    not read from a committed file (round-1 verdict, item 1)."""
    import inspect

    src = inspect.getsource(bench.run_band)
    assert "single_gpu_reference(" in src and "band_parity_check(" in src
    assert not hasattr(bench, "scaling_reference")
    assert "profiles/r01_scaling" not in inspect.getsource(bench)
