"""bench.py contract checks that need no GPU: the reference arm prints one JSON line with the required
keys, and the GPU arm refuses to run without a device (no CPU fallback)."""
import json
import os
import subprocess
import sys

import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def test_reference_arm_json_line():
    r = subprocess.run([sys.executable, "bench.py", "--impl", "reference", "--steps", "1", "--warmup", "0"],
                       cwd=ROOT, capture_output=True, text=True, timeout=600)
    assert r.returncode == 0, r.stderr
    line = json.loads(r.stdout.strip().splitlines()[-1])
    for k in ("impl", "metric", "value", "unit", "n_gpus", "steps", "warmup", "ms_per_step", "higher_is_better",
              "scaling", "vs_baseline", "dtype", "data", "config", "cpu_baseline", "e2e"):
        assert k in line, k
    assert line["impl"] == "reference" and line["unit"] == "Gpix/s" and line["value"] > 0
    assert line["cpu_baseline"]["kind"] in ("reference", "port") and line["cpu_baseline"]["cores"] >= 1
    assert line["e2e"]["h2d_bytes_per_step"] == 0 and line["e2e"]["d2h_bytes_per_step"] == 0
    assert "workload" in line["config"] and "model" not in line["config"]


def test_reference_arm_other_ranks_exit_quietly():
    env = dict(os.environ, RANK="1", LOCAL_RANK="1", WORLD_SIZE="2")
    r = subprocess.run([sys.executable, "bench.py", "--impl", "reference", "--gpus", "2", "--steps", "1"],
                       cwd=ROOT, capture_output=True, text=True, timeout=120, env=env)
    assert r.returncode == 0 and r.stdout.strip() == ""


def test_gpu_arm_fails_loudly_without_a_device():
    import torch
    if torch.cuda.is_available():
        pytest.skip("a device is present")
    r = subprocess.run([sys.executable, "bench.py", "--steps", "1"], cwd=ROOT, capture_output=True, text=True, timeout=300)
    assert r.returncode != 0 and "no CUDA device" in (r.stderr + r.stdout)


def test_live_traffic_parses_an_ncu_csv_log(tmp_path):
    """bench.py reads roofline.traffic from an `ncu --csv --log-file` of a child process: header noise, thousands separators
    and the byte units ncu picks (byte ... Gbyte) must all be handled; two launches (the two-pass step) add up."""
    sys.path.insert(0, ROOT)
    import bench
    hdr = '"ID","Process ID","Process Name","Host Name","Kernel Name","Context","Stream","Block Size","Grid Size","Device","CC","Section Name","Metric Name","Metric Unit","Metric Value"\n'
    row = '"{i}","77","python3.12","127.0.0.1","{k}","1","7","(1024, 1, 1)","(148, 1, 1)","0","10.0","Command line profiler metrics","{m}","{u}","{v}"\n'
    log = tmp_path / "t.csv"
    log.write_text("==PROF== Connected to process 77 (/usr/bin/python3.12)\n==WARNING== Note: Running with unmodified GPU clocks.\n" + hdr +
                   row.format(i=0, k="k_decode_lin", m="dram__bytes_read.sum", u="Mbyte", v="466.58") +
                   row.format(i=0, k="k_decode_lin", m="dram__bytes_write.sum", u="Gbyte", v="1.188575") +
                   row.format(i=1, k="k_rescale_pair", m="dram__bytes_read.sum", u="byte", v="937,663,488") +
                   row.format(i=1, k="k_rescale_pair", m="dram__bytes_write.sum", u="Kbyte", v="1,428,063") +
                   row.format(i=1, k="k_rescale_pair", m="gpu__time_duration.sum", u="usecond", v="666.59") +
                   "==PROF== Disconnected from process 77\n")
    total, seen = bench.parse_ncu_dram_bytes(str(log))
    assert seen == 4
    assert abs(total - (466.58e6 + 1.188575e9 + 937663488 + 1428063e3)) < 1.0
    empty = tmp_path / "e.csv"
    empty.write_text("==ERROR== nothing profiled\n")
    assert bench.parse_ncu_dram_bytes(str(empty)) == (0.0, 0)
