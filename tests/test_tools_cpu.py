"""CPU checks of the measurement helpers under tools/ (no GPU): the schedule picker's rule and
the bookkeeping that ties bench.py's `roofline.traffic` to the schedule it was captured on."""
import json
import os
import subprocess
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def _pick(rows, tmp_path):
    p = tmp_path / "probe.json"
    p.write_text(json.dumps({"schedules": rows}))
    out = subprocess.run([sys.executable, os.path.join(ROOT, "tools", "pick_schedule.py"), str(p)],
                         capture_output=True, text=True, check=True).stdout.splitlines()
    return out[0], [int(i) for i in out[1].split()]


def _row(i, g, d, w, tf):
    return {"index": i, "groups": g, "div": d, "hint": 0, "window": w, "epoch_shift": 6, "tflops": tf}


def test_picker_keeps_default_when_windowed_schedules_are_slower(tmp_path):
    # the round-1 outcome: every L2-served schedule ~13 % slower than the HBM-streamed default
    rows = [_row(0, 64, 256, 0, 32.0), _row(1, 64, -1, 4, 27.8), _row(2, 16, 16, 2, 28.7), _row(3, 1, 0, 16, 28.6)]
    pick, idx = _pick(rows, tmp_path)
    assert pick == "64,256,0,0,6"
    assert idx[0] == 0 and idx[1] == 2          # default first, then the fastest windowed schedule


def test_picker_prefers_fewest_fronts_among_fast_schedules(tmp_path):
    rows = [_row(0, 64, 256, 0, 32.0), _row(1, 16, 16, 2, 32.1), _row(2, 64, -1, 4, 31.9), _row(3, 8, 8, 2, 31.0)]
    pick, idx = _pick(rows, tmp_path)
    assert pick == "64,-1,0,4,6"                # one coherent front (div < 0) beats 16 fronts; 8 fronts too slow
    assert idx[:2] == [0, 2]


def test_traffic_json_names_the_schedule_it_was_captured_on():
    t = json.load(open(os.path.join(ROOT, "profiles", "traffic.json")))
    for wl, entry in t.items():
        assert {"bytes_per_launch", "source", "algorithmic_bytes_per_launch"} <= set(entry), wl
        if wl != "dense":                       # the dense GEMM has no k-sweep schedule
            assert len(entry["schedule"].split(",")) == 5
        assert entry["bytes_per_launch"] >= entry["algorithmic_bytes_per_launch"]
