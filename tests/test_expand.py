"""The device expander's per-read / per-(read, column) logic (csrc/expand_core.h, the functions the kernels of
csrc/expand.cu wrap), run serially on the CPU by tests/expand_check.cpp and compared with the host staging layer array by
array: every stream it builds must equal staging.cpp's bit for bit (nine synthetic datasets: single-end, paired, several
read files, deep columns, read_pos / base_repeat covariates, shards, the preprocess stage, histogram-only and scoring-only
runs, and reads flagged secondary / QC-fail / duplicate / unmapped)."""
import os
import subprocess

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def test_expander_logic_equals_host_staging(tmp_path):
    exe = str(tmp_path / "expand_check")
    csrc = os.path.join(ROOT, "breseq_b200", "csrc")
    srcs = [os.path.join(ROOT, "tests", "expand_check.cpp")] + [os.path.join(csrc, f) for f in
                                                                 ("staging.cpp", "synth.cpp", "bam_io.cpp", "inflate.cpp", "expand_plan.cpp")]
    subprocess.run(["g++", "-O2", "-std=c++17", "-w", "-o", exe] + srcs + ["-lz", "-lpthread"], check=True)
    p = subprocess.run([exe], capture_output=True, text=True)
    assert p.returncode == 0, p.stdout + p.stderr
    assert p.stdout.count("equal") >= 11 and "DIFFERENT" not in p.stdout
