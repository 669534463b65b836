"""The driver-facing contract of bench.py that can be checked without a GPU: the reference arm (the CPU oracle timed
on the host cores) prints exactly ONE JSON line with the agreed keys, and the per-shape profile summariser parses the
table `bench.py --prof-dump` writes."""
import json
import os
import subprocess
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def test_reference_arm_prints_one_json_line():
    r = subprocess.run([sys.executable, os.path.join(ROOT, "bench.py"), "--impl", "reference", "--steps", "1",
                        "--warmup", "0", "--cpu-frames", "4"], capture_output=True, text=True, timeout=600, cwd=ROOT)
    assert r.returncode == 0, r.stderr[-2000:]
    lines = [l for l in r.stdout.splitlines() if l.strip()]
    assert len(lines) == 1, r.stdout
    d = json.loads(lines[0])
    assert d["impl"] == "reference" and d["unit"] == "clips/s" and d["higher_is_better"] is True
    assert d["value"] > 0 and d["ms_per_step"] > 0
    # "reference" when the git-ignored copy baseline/_ref is present (the unmodified reference's own Trainer), else the port
    assert d["cpu_baseline"]["kind"] in ("port", "reference") and d["cpu_baseline"]["cores"] >= 1
    assert d["e2e"] == {"value": d["value"], "unit": "clips/s", "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0}
    assert d["gpu_launches"] == 0 and d["vs_baseline"] is None


def test_reference_arm_port_fallback():
    r = subprocess.run([sys.executable, os.path.join(ROOT, "bench.py"), "--impl", "reference", "--steps", "1",
                        "--warmup", "0", "--cpu-frames", "4", "--force-port"], capture_output=True, text=True,
                       timeout=600, cwd=ROOT)
    assert r.returncode == 0, r.stderr[-2000:]
    d = json.loads(r.stdout.strip().splitlines()[-1])
    assert d["cpu_baseline"]["kind"] == "port" and d["value"] > 0


def test_workload_names_follow_the_arguments():
    sys.path.insert(0, ROOT)
    import argparse
    import bench
    a = argparse.Namespace(frames=48, latent_dim=8, classes=101, ch=32, k_sample=8)
    assert bench.metric_name(a) == "clips/sec (48f x 128x128) G+Ds+Dt step"
    assert "configs[2]" in bench.workload_name(a, 32, 8) and "32/GPU x 8 GPU" in bench.workload_name(a, 32, 8)
    assert bench.step_work(a) == (32.76, 15.40)
    a.frames = 20
    assert "custom shape" in bench.workload_name(a, 4, 1) and bench.step_work(a) == (None, None)


def test_reference_arm_other_ranks_stay_silent():
    env = dict(os.environ, RANK="1", WORLD_SIZE="2")
    r = subprocess.run([sys.executable, os.path.join(ROOT, "bench.py"), "--impl", "reference", "--gpus", "2"],
                       capture_output=True, text=True, timeout=120, cwd=ROOT, env=env)
    assert r.returncode == 0 and r.stdout.strip() == ""


def test_prof_summary_parses(tmp_path):
    tsv = tmp_path / "prof.tsv"
    tsv.write_text("0\tfwd M65536 Ci256 Co512 t25 bn256 pair acc1\t47\t52.7\t2.0e13\n"
                   "1\twgrad M3145728 Ci256 Co512 t25 bn256 pair ns4\t1\t40.0\t2.0e13\n"
                   "2\tprep N3072 pix1024 C128\t26\t21.2\t8.4e10\n"
                   "3\tgru_bwd1\t576\t23.4\t0\n")
    r = subprocess.run([sys.executable, os.path.join(ROOT, "profiles", "summarize_prof.py"), str(tsv), "1800", "1"],
                       capture_output=True, text=True, timeout=60)
    assert r.returncode == 0, r.stderr
    assert "conv fwd/dgrad GEMM: 47 launches/step" in r.stdout and "379.5 TF/s" in r.stdout
    assert "operand-plane prep" in r.stdout and "memory-bound helpers" in r.stdout
