"""bench.py's contract, as far as it can be checked without a GPU: the reference arm prints one JSON line with the
product arm's metric / unit / config and the keys the driver reads; the committed DRAM-traffic record of the headline
kernel still belongs to the kernel sources in the tree (otherwise `roofline.traffic` would be reported as null)."""
import json
import os
import subprocess
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def test_reference_arm_line():
    out = subprocess.run([sys.executable, os.path.join(ROOT, "bench.py"), "--impl", "reference", "--steps", "1", "--warmup", "0",
                          "--ref-n", "8", "--ref-solve-n", "6"], capture_output=True, text=True, check=True, cwd=ROOT).stdout
    lines = [l for l in out.splitlines() if l.startswith("{")]
    assert len(lines) == 1
    line = json.loads(lines[0])
    sys.path.insert(0, ROOT)
    import bench

    assert line["impl"] == "reference" and line["metric"] == bench.METRIC and line["unit"] == bench.UNIT
    assert line["higher_is_better"] is True and line["n_gpus"] == 1 and line["steps"] == 1 and line["warmup"] == 0
    assert line["config"]["workload"] == bench.WORKLOAD % 256 and line["config"]["cells"] == 256 ** 3
    cpu = line["cpu_baseline"]
    assert cpu["kind"] == "port" and cpu["cores"] == 1 and cpu["value"] == line["value"] and "sample" in cpu
    assert line["e2e"] == {"value": line["value"], "unit": bench.UNIT, "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0}
    assert line["gpu_launches"] == 0 and line["value"] > 0.


def test_other_ranks_of_the_reference_arm_do_nothing():
    env = dict(os.environ, RANK="1", LOCAL_RANK="1", WORLD_SIZE="2")
    r = subprocess.run([sys.executable, os.path.join(ROOT, "bench.py"), "--impl", "reference", "--gpus", "2", "--steps", "1",
                        "--warmup", "0"], capture_output=True, text=True, cwd=ROOT, env=env)
    assert r.returncode == 0 and not [l for l in r.stdout.splitlines() if l.startswith("{")]


def test_traffic_record_matches_the_kernel_sources():
    sys.path.insert(0, ROOT)
    import bench

    traffic, source = bench.measured_traffic(256, 1)
    assert traffic is not None, source
    algorithmic = bench.ASM_BYTES_PER_CELL * 256 ** 3
    assert 1.0 <= traffic / algorithmic <= 1.10   # 1.04 x at the capture
    assert bench.measured_traffic(256, 8)[0] is None and bench.measured_traffic(128, 1)[0] is None


def test_bytes_of_the_benched_pcg_iteration():
    """roofline_pcg_auxmg's numerator (bench.py:auxmg_iteration_bytes) at the 256^3 sizes of the committed bench line:
    17.25 GB per iteration -> 82.5 % of the measured copy peak at the measured 3.19 ms."""
    sys.path.insert(0, ROOT)
    import bench

    line = json.load(open(os.path.join(ROOT, "profiles", "r2_bench_default.json")))
    rows, nnz, nc = line["config"]["face_rows"], line["config"]["face_system_nnz"], line["config"]["cells"]
    nf = 3 * 256 * 256 * 257
    b = bench.auxmg_iteration_bytes(rows, nnz, nc, nf)
    assert abs(b - (8 * nnz + 109 * rows + 200 * nc + 55 * nf + 64 * nc * 8 / 7)) < 1.
    assert 17.2e9 < b < 17.3e9
    frac = b / line["pcg_s_per_iter"] / 1e9 / line["roofline"]["peak"]
    assert 0.80 < frac < 0.85
    # the Jacobi-term / FP64-level variants move less / more
    assert bench.auxmg_iteration_bytes(rows, nnz, nc, nf, blocks=False) == b - 84 * nc
    assert bench.auxmg_iteration_bytes(rows, nnz, nc, nf, fp32_cycle=False) == b + 72 * nc * 8 / 7
