"""The C++ host side above the C ABI (include/katome_gpu.hpp, examples/katome_build.cpp): it compiles
against the header and the in-tree library, fails loudly without a device, and on a GPU reproduces the
counts the reference's own tests pin (tests/build.rs:27-28, tests/pruner.rs:240-251, hm_gir.rs:40)."""
import os
import subprocess

import pytest

from katome_b200 import _lib
from tests import helpers as H
from tests.helpers import ROOT


@pytest.fixture(scope="module")
def exe(tmp_path_factory):
    out = str(tmp_path_factory.mktemp("cpp") / "katome_build")
    libdir = os.path.dirname(_lib.SO_PATH)
    cmd = ["g++", "-std=c++17", "-O1", "-Wall", "-Wextra", "-Werror", "-I", os.path.join(ROOT, "include"),
           os.path.join(ROOT, "examples", "katome_build.cpp"), "-L", libdir, "-lkatome_gpu",
           f"-Wl,-rpath,{libdir}", "-o", out]
    r = subprocess.run(cmd, capture_output=True, text=True)
    assert r.returncode == 0, r.stderr
    return out


def run(exe, *args):
    return subprocess.run([exe, *map(str, args)], capture_output=True, text=True, timeout=300)


def test_cpp_host_compiles_and_has_no_cpu_fallback(exe, tmp_path):
    assert run(exe).returncode == 2  # usage
    if _lib.lib().ktg_device_count() > 0:
        pytest.skip("a GPU is present")
    fq = H.write_fastq(tmp_path / "d.fastq", H.golden_seqs("data1"))
    r = run(exe, 40, 0, 0, 0, fq)
    assert r.returncode == 3 and "no CPU fallback" in r.stderr and r.stdout == ""


@pytest.mark.gpu
def test_cpp_host_reproduces_the_reference_fixtures(exe, golden, tmp_path):
    for name in ("data1", "data2", "data3"):
        ref = golden["reference_pinned"][name]
        fq = H.write_fastq(tmp_path / f"{name}.fastq", H.golden_seqs(name))
        r = run(exe, 40, 0, 0, 0, fq)
        assert r.returncode == 0, r.stderr
        v = [int(x) for x in r.stdout.split()]
        assert v[:3] == [ref["bytes"], ref["nodes"], ref["edges"]], (name, v)
        assert v[9:] == [ref["nodes"], ref["edges"]]  # the graph handed to Convert::create_from
    # Clean on the GIR: data1 with threshold 3 loses everything (tests/pruner.rs:240-251)
    fq = H.write_fastq(tmp_path / "data1.fastq", H.golden_seqs("data1"))
    r = run(exe, 40, 0, 3, 0, fq)
    assert r.returncode == 0 and [int(x) for x in r.stdout.split()][1:3] == [0, 0]
    # "Read is too short!" (hm_gir.rs:40) arrives as the same panic text
    fq = H.write_fastq(tmp_path / "short.fastq", golden["too_short_seqs"], qual_len=100)
    r = run(exe, 40, 0, 0, 0, fq)
    assert r.returncode == 3 and "Read is too short!" in r.stderr
    # a missing file (builder.rs:57-77)
    r = run(exe, 40, 0, 0, 0, str(tmp_path / "nope.fastq"))
    assert r.returncode == 3


@pytest.mark.gpu
def test_cpp_host_on_several_shards_writes_the_same_dump(exe, tmp_path):
    """One table sharded over 1, 2, 3, 4 and 8 devices (here: shards of one GPU, `--devices 0,0,...`; on a
    multi-GPU box scripts/r02_multi.sh runs the same with distinct devices): the sorted edge dump of SURVEY
    Appendix A.12 and the line of graph statistics are byte-identical, and Build::create stays one call."""
    import numpy as np
    rng = np.random.default_rng(12)
    genome = "".join(rng.choice(list("ACGT"), size=40_000))
    seqs = H.random_reads(rng, 6000, 70, 150, genome=genome, n_rate=0.02) + ["T" * 90, "AT" * 45, "ACGT" * 30]
    fq = H.write_fastq(tmp_path / "reads.fastq", seqs)
    for k, rc in ((31, 1), (40, 0), (63, 1)):
        ref_line = ref_dump = None
        for devices in (None, "0,0", "0,0,0", "0,0,0,0", "0,0,0,0,0,0,0,0"):
            dump = tmp_path / f"dump_{k}_{devices}.txt"
            args = (["--devices", devices] if devices else []) + ["--dump", dump, k, rc, 0, 0, fq]
            r = run(exe, *args)
            assert r.returncode == 0, (devices, r.stderr)
            text = open(dump).read()
            if ref_line is None:
                ref_line, ref_dump = r.stdout, text
                assert text.count("\n") == int(r.stdout.split()[2]) + 1
            else:
                assert r.stdout == ref_line, (k, devices)
                assert text == ref_dump, (k, devices)
    # the panics of the reference arrive through a sharded handle as well (hm_gir.rs:40)
    short = H.write_fastq(tmp_path / "short.fastq", seqs[:50] + ["ACGTACGT"] + seqs[50:100])
    r = run(exe, "--devices", "0,0", 31, 1, 0, 0, short)
    assert r.returncode == 3 and "Read is too short!" in r.stderr
