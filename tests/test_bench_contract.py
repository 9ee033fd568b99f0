"""bench.py's contract on the CPU side: the file compiles, and the reference arm (the reference's own
CPU implementation of the step through oracle/_ref, the one leg of bench.py that needs no GPU) prints
one JSON line with the keys the driver reads."""
import json
import os
import py_compile
import subprocess
import sys

import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def test_bench_compiles():
    py_compile.compile(os.path.join(ROOT, "bench.py"), doraise=True)
    py_compile.compile(os.path.join(ROOT, "__graft_entry__.py"), doraise=True)


def test_reference_arm_line():
    from oracle import bindings
    if not bindings.have_ref():
        pytest.skip("oracle/_ref not built")
    env = dict(os.environ, XVCB_BENCH_SKIP_XVCENC="1")      # the CLI context sample takes ~15 s; not part of the contract
    out = subprocess.run([sys.executable, os.path.join(ROOT, "bench.py"), "--impl", "reference", "--steps", "1", "--warmup", "0"],
                         capture_output=True, text=True, timeout=600, env=env)
    assert out.returncode == 0, out.stderr[-2000:]
    line = json.loads(out.stdout.strip().splitlines()[-1])
    assert line["impl"] == "reference" and line["unit"] == "Mpixels/s" and line["higher_is_better"] is True
    assert line["value"] > 0 and line["e2e"]["value"] == line["value"]
    assert line["e2e"]["h2d_bytes_per_step"] == 0 and line["e2e"]["d2h_bytes_per_step"] == 0
    assert line["cpu_baseline"]["kind"] in ("reference", "port") and line["cpu_baseline"]["cores"] >= 1
    assert line["config"]["workload"].startswith("1920x1080")


def test_extras_watchdog(tmp_path):
    """bench.run_extras_guarded: the JSON line is complete before the GOP / banded measurements; when they finish
    their results are merged, when they hang rank 0 still prints the line (with the reason) and the process leaves
    with exit code 0 -- on every rank, printing on rank 0 only."""
    import json
    import subprocess
    import sys
    root = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
    script = tmp_path / "guard.py"
    script.write_text(
        "import json, sys, time\n"
        "sys.path.insert(0, %r)\n"
        "import bench\n"
        "mode, rank = sys.argv[1], int(sys.argv[2])\n"
        "line = {'metric': 'm', 'value': 1.0}\n"
        "def work():\n"
        "    if mode == 'hang':\n"
        "        time.sleep(600)\n"
        "    return {'gop': {'value': 2.0}, 'banded': None}\n"
        "bench.run_extras_guarded(line, rank, 1.0, work)\n"
        "if rank == 0:\n"
        "    print(json.dumps(line))\n" % root)
    ok = subprocess.run([sys.executable, str(script), "ok", "0"], capture_output=True, text=True, timeout=120)
    assert ok.returncode == 0, ok.stderr
    assert json.loads(ok.stdout.strip().splitlines()[-1]) == {"metric": "m", "value": 1.0, "gop": {"value": 2.0}}
    hung = subprocess.run([sys.executable, str(script), "hang", "0"], capture_output=True, text=True, timeout=120)
    assert hung.returncode == 0, hung.stderr
    lines = [l for l in hung.stdout.strip().splitlines() if l.startswith("{")]
    assert len(lines) == 1
    got = json.loads(lines[0])
    assert got["value"] == 1.0 and "unavailable" in got["gop"]
    other = subprocess.run([sys.executable, str(script), "hang", "3"], capture_output=True, text=True, timeout=120)
    assert other.returncode == 0 and other.stdout.strip() == ""
