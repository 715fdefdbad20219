"""The reference arm of bench.py runs on the host alone: check its JSON line against the measurement contract."""
import json
import os
import subprocess
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def test_reference_arm_prints_the_contract_line(oracle_mod):
    r = subprocess.run([sys.executable, os.path.join(ROOT, "bench.py"), "--impl", "reference", "--steps", "2", "--warmup", "1",
                        "--frames", "8", "--unique", "2"], capture_output=True, text=True, timeout=600, cwd=ROOT)
    assert r.returncode == 0, r.stderr[-2000:]
    line = json.loads(r.stdout.strip().splitlines()[-1])
    assert line["impl"] == "reference" and line["unit"] == "frames/s" and line["higher_is_better"] is True
    assert line["value"] > 0 and line["steps"] == 2 and line["warmup"] == 1 and line["n_gpus"] == 1
    assert line["config"]["workload"].startswith("configs[2]")
    cb = line["cpu_baseline"]
    assert cb["kind"] == "port" and cb["cores"] >= 1 and cb["value"] == line["value"] and "frames/step" in cb["sample"]
    assert line["e2e"] == {"value": line["value"], "unit": "frames/s", "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0}


def test_reference_arm_other_ranks_do_nothing(oracle_mod):
    env = dict(os.environ, RANK="1", WORLD_SIZE="2", LOCAL_RANK="1")
    r = subprocess.run([sys.executable, os.path.join(ROOT, "bench.py"), "--impl", "reference", "--gpus", "2", "--steps", "1"],
                       capture_output=True, text=True, timeout=120, cwd=ROOT, env=env)
    assert r.returncode == 0 and r.stdout.strip() == ""


import pytest  # noqa: E402


@pytest.mark.gpu
def test_gpu_arm_prints_the_contract_line():
    r = subprocess.run([sys.executable, os.path.join(ROOT, "bench.py"), "--steps", "5", "--warmup", "3", "--frames", "16",
                        "--unique", "4", "--cpu-sample-frames", "4", "--seq-frames", "1000", "--latency-calls", "40"], capture_output=True, text=True, timeout=900, cwd=ROOT)
    assert r.returncode == 0, r.stderr[-2000:]
    line = json.loads(r.stdout.strip().splitlines()[-1])
    for key in ("metric", "value", "unit", "n_gpus", "steps", "warmup", "ms_per_step", "higher_is_better", "scaling",
                "vs_baseline", "dtype", "data", "config", "clocks", "e2e", "gpu_launches", "roofline", "cpu_baseline"):
        assert key in line, key
    assert line["n_gpus"] == 1 and line["steps"] == 5 and line["scaling"] == "weak" and line["vs_baseline"] is None
    assert line["gpu_launches"] >= 3 * 5 and line["value"] > 0
    rf = line["roofline"]
    assert rf["bound"] == "hbm" and rf["unit"] == "GB/s" and abs(rf["frac"] - rf["achieved"] / rf["peak"]) < 1e-9
    assert set(rf["stages"]) == {"cell_stats", "region_grow", "labeling"}
    e2e = line["e2e"]
    assert e2e["h2d_bytes_per_step"] == 16 * 480 * 640 * 12 and e2e["d2h_bytes_per_step"] == 16 * 480 * 640 * 4
    assert e2e["matches_device_path"] is True and line["e2e_depth16"]["matches_point_path"] is True
    assert 0 < e2e["frac_of_pcie_ceiling"] < 1.5 and e2e["pcie_ceiling"]["h2d_gbs"] > 1
    assert line["e2e_depth16"]["d2h_bytes_per_step"] == 16 * 480 * 640 * 4
    assert line["e2e_depth16"]["labels_u16"]["d2h_bytes_per_step"] == 16 * 480 * 640 * 2 and line["e2e_depth16"]["labels_u16"]["matches_int32_path"]
    assert line["cpu_baseline"]["kind"] in ("port", "reference") and line["cpu_baseline"]["cores"] == 1
    # configs[0] / configs[1]: latency with the CPU path and a roofline beside it
    for name, planes in (("tum", 34), ("icl", None)):
        lat = line["latency"][name]
        assert lat["cpu_baseline"]["cores"] == 1 and lat["cpu_baseline"]["value"] > 0
        assert set(lat["roofline"]["stages"]) == {"cell_stats", "region_grow", "labeling"}
        assert lat["host_ptr_us"]["min"] <= lat["host_ptr_us"]["mean"] <= lat["host_ptr_us"]["max"]
        if planes:
            assert lat["planes"] == planes
    # configs[3]: 1920x1080 at patch 10, 8 and 5 (82 944 cells)
    fhd = line["fhd_stress"]
    assert fhd["patch5"]["cells"] == 82944 and fhd["patch10"]["cells"] == 20736
    for rec in fhd.values():
        assert rec["e2e"]["matches_device_path"] is True and rec["cpu_baseline"]["value"] > 0
        assert rec["roofline"]["stages"]["cell_stats"]["frac"] > 0
    # configs[4]: the sharded sequence (shortened here), verified frame by frame inside the bench
    seq = line["sharded_sequence"]
    assert seq["frames"] == 1000 and seq["frames_with_wrong_labels"] == 0 and seq["checksum"] > 0
