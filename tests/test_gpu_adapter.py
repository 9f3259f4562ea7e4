"""The drop-in compiled for real: oracle/_ref/ref_cli_brq is the reference's own objects (Settings, Summary,
cReferenceSequences, GenomeDiff, ... unmodified) with breseq::error_count() and breseq::identify_mutations() provided by
adapters/breseq_adapter.cpp over libbrq.so (oracle/ref_build.sh).  Run with the command line of ref_cli (real Settings and
Summary objects filled the way breseq_cmdline.cpp fills them), it must write the files the reference build wrote: the golden
files under tests/golden/, byte for byte."""
import filecmp
import os
import subprocess

import pytest

import helpers

REF_CLI_BRQ = os.path.join(helpers.ROOT, "oracle", "_ref", "ref_cli_brq")
NAMES = [n for n in helpers.DATASETS if not helpers.DATASETS[n].get("no_golden")]


def run_brq(args, **kw):
    return subprocess.run([REF_CLI_BRQ] + [str(a) for a in args], capture_output=True, text=True, **kw)


def test_adapter_binary_fails_loudly_without_a_device(built, tmp_path):
    """(no GPU here: the binary must load, reach libbrq.so through the adapter and refuse, not fall back)"""
    if not os.path.exists(REF_CLI_BRQ):
        pytest.skip("oracle/_ref/ref_cli_brq is built where /root/reference exists")
    import torch
    if torch.cuda.is_available():
        pytest.skip("a device is present")
    d = helpers.generate_inputs("tiny", str(tmp_path / "tiny"))
    ec, _ = helpers.cli_args(d, str(tmp_path))
    p = run_brq(ec)
    assert p.returncode != 0 and "libbrq" in (p.stdout + p.stderr)


@pytest.mark.gpu
@pytest.mark.parametrize("name", NAMES)
def test_adapter_writes_the_reference_files(name, datasets, tmp_path):
    if not os.path.exists(REF_CLI_BRQ):
        pytest.skip("oracle/_ref/ref_cli_brq is built where /root/reference exists")
    d = datasets[name]
    out = str(tmp_path)
    ec, im = helpers.cli_args(d, out)
    im += ["--per-position", os.path.join(out, "per_position_file.tab")]
    for args in (ec, im):
        p = run_brq(args)
        assert p.returncode == 0, p.stdout[-2000:] + p.stderr[-2000:]
    gold = os.path.join(helpers.GOLDEN, name)
    for f in helpers.pass_output_names(d):
        assert filecmp.cmp(os.path.join(out, f), os.path.join(gold, f), shallow=False), f
    if name == "tiny":
        assert filecmp.cmp(os.path.join(out, "per_position_file.tab"), os.path.join(gold, "per_position_file.tab"), shallow=False)


@pytest.mark.gpu
def test_adapter_preprocess_stage_and_user_evidence(datasets, tmp_path):
    if not os.path.exists(REF_CLI_BRQ):
        pytest.skip("oracle/_ref/ref_cli_brq is built where /root/reference exists")
    d = datasets["lambda"]
    out = str(tmp_path)
    ec, im = helpers.cli_args(d, out)
    p = run_brq(ec + ["--preprocess", "--no-errors"])
    assert p.returncode == 0, p.stderr[-2000:]
    assert open(os.path.join(out, helpers.PREPROCESS_TAB)).read() == open(os.path.join(helpers.GOLDEN, "lambda", helpers.PREPROCESS_TAB)).read()
    assert run_brq(ec).returncode == 0
    p = run_brq(im + ["--user-evidence", os.path.join(helpers.GOLDEN, "lambda", "user_evidence.gd")])
    assert p.returncode == 0, p.stderr[-2000:]
    assert open(os.path.join(out, "ra_mc_evidence.gd")).read() == open(os.path.join(helpers.GOLDEN, "lambda", "ra_mc_evidence.user_evidence.gd")).read()
