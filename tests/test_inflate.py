"""The BGZF member decoder (csrc/inflate.cpp) against zlib: tests/inflate_check.cpp compresses data of several kinds at every
level and strategy (dynamic, fixed, Huffman-only, run-length, stored blocks), decodes with fast_inflate and compares byte for
byte; truncated / corrupted streams and a short output buffer must be refused without a byte written outside the output."""
import os
import subprocess

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def test_fast_inflate_equals_zlib(tmp_path):
    exe = str(tmp_path / "inflate_check")
    subprocess.run(["g++", "-O2", "-std=c++17", "-o", exe, os.path.join(ROOT, "tests", "inflate_check.cpp"),
                    os.path.join(ROOT, "breseq_b200", "csrc", "inflate.cpp"), "-lz"], check=True)
    p = subprocess.run([exe, "2"], capture_output=True, text=True)
    assert p.returncode == 0, p.stdout + p.stderr
    assert "streams equal" in p.stdout


def test_bam_reads_the_same_with_either_inflate(tmp_path):
    """read_bam through our decoder and through zlib (BRQ_ZLIB_INFLATE=1): the same staged stream"""
    import sys
    code = ("import sys; sys.path.insert(0, %r); sys.path.insert(0, %r); import hashlib, helpers, breseq_b200 as bq\n"
            "d = helpers.generate_inputs('multi', %r)\n"
            "c = bq.Context(device=-1); c.stage_bam(d['bam'], d['fasta'], **helpers.stage_kwargs(d)); s = c.stream()\n"
            "print(hashlib.sha256(s['score_rec'].tobytes() + s['hist_rec'].tobytes() + s['side_rec'].tobytes()).hexdigest())\n"
            % (ROOT, os.path.join(ROOT, "tests"), str(tmp_path)))
    a = subprocess.run([sys.executable, "-c", code], capture_output=True, text=True, env=dict(os.environ))
    b = subprocess.run([sys.executable, "-c", code], capture_output=True, text=True, env=dict(os.environ, BRQ_ZLIB_INFLATE="1"))
    assert a.returncode == 0 and b.returncode == 0, a.stderr + b.stderr
    assert a.stdout.strip() and a.stdout == b.stdout
