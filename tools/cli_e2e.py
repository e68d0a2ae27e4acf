#!/usr/bin/env python
"""Whole-program timing on real files: synthetic chr19-sized BAM + FASTA on disk -> the drop-in pipeline
(native decode -> GPU filter/depth/scan/score -> GPU-gzipped .depth.gz, .bed, .gci)."""
import contextlib
import io
import json
import os
import sys
import tempfile
import time

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from gci_b200 import synth, io as gio, pipeline as P  # noqa: E402


def main():
    L = int(float(sys.argv[1])) if len(sys.argv) > 1 else 58_000_000
    threads = os.cpu_count() or 1
    d = synth.make_reads(synth.SynthSpec([L], coverage=30, seed=20240634, contig_names=["chr19"]))
    work = tempfile.mkdtemp(prefix="gci_cli_")
    bam, fa = os.path.join(work, "hifi.bam"), os.path.join(work, "ref.fa")
    t0 = time.time()
    gio.write_bam(bam, d.contigs.names, d.contigs.lengths, d.bam, level=1)
    with open(fa, "w") as f:                      # no N-runs needed for the timing; 60 columns like most assemblies
        f.write(">chr19\n")
        line = "ACGT" * 15 + "\n"
        f.write(line * (L // 60))
        f.write("A" * (L % 60) + "\n")
    prep = time.time() - t0
    out = {"genome_bases": L, "records": d.bam.n_records, "aligned_bases": d.aligned_bases,
           "bam_bytes": os.path.getsize(bam), "host_threads": threads, "prepare_files_s": prep}
    ses = P.Session(0)
    ses.ctx.set_contigs([L])                       # context + CUDA module load outside the timed run
    t_all = time.time()
    with contextlib.redirect_stdout(io.StringIO()):
        P.GCI(hifi=[bam], nano=None, directory=os.path.join(work, "out"), prefix="chr19", reference=fa, threads=threads,
              force=True, session=ses)
    out["cli_path_s"] = time.time() - t_all
    # the same again, stage by stage
    t = time.time(); gio.read_fasta_gaps(fa); out["fasta_scan_s"] = time.time() - t
    t = time.time(); nt = gio.NameTable(); names, lens, tab = gio.read_bam(bam, nt, threads); out["bam_decode_s"] = time.time() - t
    ctx = ses.ctx
    t = time.time()
    ctx.reads_begin(len(nt)); ctx.upload_bam(tab)
    res = ctx.pipeline(0, 1)
    out["upload_and_gpu_path_s"] = time.time() - t
    t = time.time()
    gz = ctx.depth_gzip(0, 0, header=b">chr19\n")
    out["depth_gz_on_gpu_s"] = time.time() - t
    out["depth_gz_bytes"] = int(gz.nbytes)
    out["depth_text_bytes"] = int(sum(len(str(v)) + 1 for v in [0])) and None
    out["survivors"], out["issue_intervals"] = int(res[0]), int(res[1])
    out["files"] = {fn: os.path.getsize(os.path.join(work, "out", fn)) for fn in sorted(os.listdir(os.path.join(work, "out")))}
    out["aligned_gbases_per_s_whole_program"] = d.aligned_bases / out["cli_path_s"] / 1e9
    print(json.dumps(out))


if __name__ == "__main__":
    main()
