"""`-m "not gpu"`: the host-side helpers of bench.py that the driver's parsing depends on (no GPU, no rendering)."""
import importlib.util
import json
import os
import subprocess
import sys

import numpy as np

from conftest import ROOT


def _bench():
    spec = importlib.util.spec_from_file_location("bench_mod", os.path.join(ROOT, "bench.py"))
    m = importlib.util.module_from_spec(spec)
    spec.loader.exec_module(m)
    return m


def test_foreign_stdout_is_kept_off_the_json_line(tmp_path):
    """The reference loader printf()s without a newline (src/n3tree.cpp:264); whatever C code writes to fd 1 inside
    StdoutToStderr must not reach stdout, and the JSON line printed afterwards must stand alone."""
    script = tmp_path / "s.py"
    script.write_text(
        "import ctypes, json, sys\n"
        f"sys.path.insert(0, {ROOT!r})\n"
        "import importlib.util\n"
        f"spec = importlib.util.spec_from_file_location('b', {os.path.join(ROOT, 'bench.py')!r})\n"
        "b = importlib.util.module_from_spec(spec); spec.loader.exec_module(b)\n"
        "libc = ctypes.CDLL(None)\n"
        "with b.StdoutToStderr():\n"
        "    libc.printf(b'INFO: Scale 0.333333 0.333333 0.333333')\n"
        "print(json.dumps({'impl': 'reference', 'value': 1.0}), flush=True)\n")
    r = subprocess.run([sys.executable, str(script)], capture_output=True, text=True, timeout=120)
    assert r.returncode == 0, r.stderr
    lines = [ln for ln in r.stdout.splitlines() if ln.strip()]
    assert len(lines) == 1 and json.loads(lines[0]) == {"impl": "reference", "value": 1.0}
    assert "INFO: Scale" in r.stderr


def test_usable_cores_and_peak_lookup():
    b = _bench()
    n = b.usable_cores()
    assert 1 <= n <= (os.cpu_count() or 1)
    peak, note = b.measured_peak()
    assert peak > 1000 and ("measured" in note or "fallback" in note)


def test_cli_leg_reads_the_fps_line():
    """main_headless.cpp:230-231 prints 'ms per frame' glued to the loader's last line; the fps line stands alone."""
    import re
    out = "INFO: Scale 0.333333 0.333333 0.3333330.1872433722 ms per frame\n5340.6340000000 fps\n"
    m = re.search(r"^\s*([0-9]+\.[0-9]+) fps\s*$", out, re.M)
    assert m and abs(1000.0 / float(m.group(1)) - 0.18724) < 1e-4


def test_clock_sampler_attributes_samples_to_the_window():
    b = _bench()
    cs = b.ClockSampler(0)
    cs.proc = object.__new__(subprocess.Popen)       # never started: stop() only terminates it
    cs.proc.terminate = lambda: None
    cs.t0, cs.t1 = 100.0, 101.0
    row = "2026/01/01 00:00:00.000, 0, {clk}, 1965, 500.0, 0x0, Not Active, Not Active, Not Active, {cap}"
    cs._ts = staticmethod(lambda s: None)            # fall back to arrival times
    cs.lines = [(99.0, row.format(clk=300, cap="Not Active")), (100.2, row.format(clk=1965, cap="Not Active")),
                (100.7, row.format(clk=1950, cap="Active")), (102.0, row.format(clk=400, cap="Not Active"))]
    r = cs.stop()
    assert r["samples"] == 2 and r["samples_total"] == 4 and r["sm_mhz"] == np.median([1965, 1950])
    assert r["reasons"] == ["sw_power_cap"] and r["sm_max_mhz"] == 1965


def test_config4_constants_match_the_band_plan():
    b = _bench()
    from volrend_b200 import dist as vd
    for world in (1, 2, 4, 8):
        rows = sum(vd.band_rows(b.C4_H, b.C4_BAND, world, r) for r in range(world))
        assert rows == b.C4_H
        covered = 0
        for r in range(world):
            for (_, _, _, _, wb, n) in vd.band_scatter_plan(b.C4_W, b.C4_H, b.C4_BAND, world, r):
                covered += wb * n
        assert covered == 4 * b.C4_W * b.C4_H
