"""CLI surface (GCI.py:1037-1110): flags, dests, defaults, validation texts — CPU only."""
import os
import subprocess
import sys

import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import GCI as cli  # noqa: E402

# dest -> default, taken from the reference's argparse block / GCI() signature (GCI.py:897, :1040-1069)
REFERENCE_DEFAULTS = dict(reference=None, hifi=None, nano=None, chrs=None, regions=None, threshold=0, dist_percent=0.005,
                          threads=1, directory='.', prefix='GCI', map_qual=30, mq_cutoff=50, iden_percent=0.9,
                          ovlp_percent=0.9, clip_percent=0.1, flank_len=15, plot=False, depth_min=0.1, depth_max=4.0,
                          window_size=50000, image_type='png', force=False)


def test_defaults_and_dests_match_reference():
    args = vars(cli.build_parser("GCI.py").parse_args([]))
    assert args == REFERENCE_DEFAULTS


def test_flags_parse():
    a = vars(cli.build_parser("GCI.py").parse_args(
        "-r r.fa --hifi a.bam b.paf --nano c.bam -mq 20 --mq-cutoff 40 -ip 0.95 -op 0.8 -cp 0.05 -fl 10 -ts 3 "
        "-dp 0.01 -t 8 -d out -o P --chrs c1,c2 -R reg.bed -f".split()))
    assert a["hifi"] == ["a.bam", "b.paf"] and a["nano"] == ["c.bam"] and a["map_qual"] == 20 and a["mq_cutoff"] == 40
    assert a["iden_percent"] == 0.95 and a["ovlp_percent"] == 0.8 and a["clip_percent"] == 0.05 and a["flank_len"] == 10
    assert a["threshold"] == 3 and a["dist_percent"] == 0.01 and a["threads"] == 8 and a["directory"] == "out"
    assert a["prefix"] == "P" and a["chrs"] == "c1,c2" and a["regions"] == "reg.bed" and a["force"] is True


def test_version_and_validation_messages(tmp_path):
    r = subprocess.run([sys.executable, os.path.join(ROOT, "GCI.py"), "-v"], capture_output=True, text=True)
    assert r.stdout.strip() == "GCI version 1.0"
    r = subprocess.run([sys.executable, os.path.join(ROOT, "GCI.py"), "-r", "x.fa"], capture_output=True, text=True)
    assert r.returncode != 0 and "Please input at least one type of TGS reads alignment files" in r.stderr
    paf = tmp_path / "a.paf"
    paf.write_text("")
    r = subprocess.run([sys.executable, os.path.join(ROOT, "GCI.py"), "--hifi", str(paf)], capture_output=True, text=True)
    assert "Please input at least one PacBio HiFi reads bam file" in r.stderr
    r = subprocess.run([sys.executable, os.path.join(ROOT, "GCI.py"), "--hifi", "/nonexistent.bam"], capture_output=True,
                       text=True)
    assert '"/nonexistent.bam" is not an available file' in r.stderr
