"""csrc/canonical.h (the error table's six-significant-digit text round trip, computed without the text on the device)
against snprintf("%.6g") + strtod, the reference's own round trip (error_count.cpp:629-690): bit for bit."""
import os
import subprocess

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def test_text_canonical_matches_printf_strtod(tmp_path):
    exe = str(tmp_path / "canonical_check")
    subprocess.run(["g++", "-O2", "-std=c++17", "-I" + os.path.join(ROOT, "breseq_b200", "csrc"),
                    os.path.join(ROOT, "tests", "canonical_check.cpp"), "-o", exe], check=True)
    out = subprocess.run([exe, "300000"], capture_output=True, text=True)
    assert out.returncode == 0 and out.stdout.startswith("ok "), out.stdout[-2000:]
