"""Parity of the CUDA path (through the C ABI / the reference-shaped Python mirror) against
  * the golden outputs of the unmodified reference,
  * the CPU oracle on seeded random inputs,
  * the reference's MH63 example (depth -> BED -> .gci).
Bit-exact for depths, intervals and text; the GCI float goes through the same Python expression
on the host, so the .gci text is compared byte for byte (tolerance 0, stricter than 1e-9)."""
import os

import numpy as np
import pytest

from helpers import GOLDEN, case_names, load_case, assert_outputs_equal, mh63_depths, run_product

pytestmark = pytest.mark.gpu


@pytest.fixture(scope="module")
def session():
    from gci_b200.pipeline import Session
    s = Session(0)
    yield s
    s.close()


@pytest.mark.parametrize("name", case_names())
def test_golden_reference_outputs(name, session, tmp_path):
    case, kw, expected = load_case(name)
    got, log = run_product(kw, str(tmp_path), session, threads=case["threads"])
    assert_outputs_equal(got, expected)


def test_stdout_matches_reference(session, tmp_path):
    case, kw, expected = load_case("dual")
    got, log = run_product(kw, str(tmp_path), session)
    ref_lines = [l for l in case["stdout"].replace("<WORK>/out", "OUT").splitlines()]
    got_lines = [l for l in log.replace(str(tmp_path) + "/out", "OUT").splitlines()]
    assert got_lines == ref_lines


def test_mh63_example(session, tmp_path):
    """reference example: MH63.depth.gz -> MH63.0.depth.bed + MH63.gci, byte for byte"""
    from gci_b200 import pipeline as P
    import contextlib, io
    names, lengths, depths = mh63_depths()
    host = {n: d for n, d in zip(names, depths)}
    with contextlib.redirect_stdout(io.StringIO()):
        dev = P._adopt(host, session)
        bed = P.merge_depth(dev, "MH63", 0, 15, str(tmp_path), True, "HiFi")
        P.compute_index(dict(zip(names, lengths)), "MH63", str(tmp_path), True, [bed], ["HiFi"], 15, 0.005,
                        {}, [dev], 0, [], session=session)
    assert open(tmp_path / "MH63.0.depth.bed").read() == open(os.path.join(GOLDEN, "mh63.0.depth.bed")).read()
    assert open(tmp_path / "MH63.gci").read() == open(os.path.join(GOLDEN, "mh63.gci")).read()
    # the depth text produced on the GPU equals the example's decompressed stream for one contig
    txt = session.ctx.depth_text(0, 11).tobytes().decode()
    assert txt == "".join(f"{v}\n" for v in depths[11].tolist())


def test_bed_from_file_scores_like_reference(session, tmp_path):
    """compute_index on a plain dict of intervals (the --bed entry of utility/GCI_score.py)"""
    from gci_b200 import pipeline as P
    import contextlib, io
    names, lengths, _ = mh63_depths()
    bed = {n: [] for n in names}
    for line in open(os.path.join(GOLDEN, "mh63.0.depth.bed")):
        c, s, e = line.split("\t")
        bed[c].append((int(s), int(e)))
    from gci_b200.pipeline import Session
    s2 = Session(0)
    try:
        with contextlib.redirect_stdout(io.StringIO()):
            P.compute_index(dict(zip(names, lengths)), "B", str(tmp_path), True, [bed], ["HiFi"], 15, 0.005, {}, [],
                            0, [], session=s2)
    finally:
        s2.close()
    assert open(tmp_path / "B.gci").read() == open(os.path.join(GOLDEN, "mh63.gci")).read()


def test_cli_on_real_files(tmp_path):
    """GCI.py command line on real BAM / PAF / FASTA files written from a golden case; outputs must equal
    the reference's."""
    import subprocess, sys
    from gci_b200 import io as gio
    case, kw, expected = load_case("hifi_bam_paf")
    names, lengths = kw["names"], kw["lengths"]
    bam, paf = kw["hifi"]
    gio.write_bam(str(tmp_path / "h.bam"), names, lengths, bam)
    gio.write_paf(str(tmp_path / "h.paf"), paf, names, lengths)
    gio.write_fasta(str(tmp_path / "ref.fa"), names, lengths, kw["n_runs"])
    root = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
    r = subprocess.run([sys.executable, os.path.join(root, "GCI.py"), "-r", str(tmp_path / "ref.fa"), "--hifi",
                        str(tmp_path / "h.bam"), str(tmp_path / "h.paf"), "-d", str(tmp_path / "out"), "-o", "T", "-t", "2",
                        "-f"], capture_output=True, text=True, timeout=600)
    assert r.returncode == 0, r.stdout + r.stderr
    assert r.stdout.startswith("Used arguments:{") and r.stdout.rstrip().endswith("GCI finished!!!\nBye!!!")
    got = {}
    for fn in sorted(os.listdir(tmp_path / "out")):
        p = str(tmp_path / "out" / fn)
        got[fn] = gio.read_depth_gz_py(p) if fn.endswith(".depth.gz") else open(p).read()
    assert_outputs_equal(got, expected)


def test_resume_from_depth_gz_tool(tmp_path):
    """tools/gci_score.py (the reference's utility/GCI_score.py entry): saved .depth.gz files + FASTA ->
    the same BED / .gci as the full run of the dual golden case"""
    import gzip, subprocess, sys
    from gci_b200 import io as gio
    case, kw, expected = load_case("dual")
    names, lengths = kw["names"], kw["lengths"]
    for sfx in ("_hifi", "_nano"):
        with gzip.open(tmp_path / f"T{sfx}.depth.gz", "wb") as f:
            for n in names:
                f.write(f">{n}\n".encode() + "".join(f"{int(v)}\n" for v in expected[f"T{sfx}.depth.gz"][n]).encode())
    gio.write_fasta(str(tmp_path / "ref.fa"), names, lengths, kw["n_runs"])
    root = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
    r = subprocess.run([sys.executable, os.path.join(root, "tools", "gci_score.py"), "--hifi", str(tmp_path / "T_hifi.depth.gz"),
                        "--nano", str(tmp_path / "T_nano.depth.gz"), "-r", str(tmp_path / "ref.fa"), "-d",
                        str(tmp_path / "out"), "-o", "T", "-f"], capture_output=True, text=True, timeout=600)
    assert r.returncode == 0, r.stdout + r.stderr
    for fn in ("T.gci", "T_hifi.0.depth.bed", "T_nano.0.depth.bed", "T_two_type.0.depth.bed"):
        assert open(tmp_path / "out" / fn).read() == expected[fn], fn
    got = gio.read_depth_gz_py(str(tmp_path / "out" / "T_two_type.depth.gz"))
    for n in names:
        assert np.array_equal(got[n], expected["T_two_type.depth.gz"][n])
