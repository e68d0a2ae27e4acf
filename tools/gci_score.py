#!/usr/bin/env python
"""Resume from saved outputs, like the reference's utility/GCI_score.py: recompute the issue BED and the .gci
table from `.depth.gz` files (one per read type) or straight from BED files, on the GPU.

    python tools/gci_score.py --hifi sample_hifi.depth.gz [--nano sample_nano.depth.gz] [-ts 0 -fl 15 -dp 0.005]
                              [-r ref.fa] [-R regions.bed] [-d out] [-o prefix] [-f]
    python tools/gci_score.py --bed issues.bed --lengths contig_lengths.tsv
"""
import argparse
import os
import sys

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from gci_b200 import io as gio, pipeline as P  # noqa: E402
from gci_b200._lib import TRACK_HIFI, TRACK_NANO  # noqa: E402


def main():
    ap = argparse.ArgumentParser(description=__doc__, formatter_class=argparse.RawTextHelpFormatter)
    ap.add_argument('--hifi'); ap.add_argument('--nano'); ap.add_argument('--bed'); ap.add_argument('--lengths')
    ap.add_argument('-r', '--reference'); ap.add_argument('-R', '--regions')
    ap.add_argument('-ts', '--threshold', type=int, default=0); ap.add_argument('-fl', '--flank-len', type=int, default=15)
    ap.add_argument('-dp', '--dist-percent', type=float, default=0.005)
    ap.add_argument('-d', dest='directory', default='.'); ap.add_argument('-o', dest='prefix', default='GCI')
    ap.add_argument('-f', '--force', action='store_true')
    a = ap.parse_args()
    os.makedirs(a.directory, exist_ok=True)
    ses = P.Session(0)
    if a.bed:
        lengths = {}
        for line in open(a.lengths):
            name, n = line.split()[:2]
            lengths[name] = int(n)
        bed = {t: [] for t in lengths}
        for line in open(a.bed):
            t, s, e = line.strip().split('\t')[:3]
            bed[t].append((int(s), int(e)))
        P.compute_index(lengths, a.prefix, a.directory, a.force, [bed], ['HiFi'], a.flank_len, a.dist_percent, {}, [],
                        a.threshold, [], session=ses)
        return
    Ns_bed = gio.read_fasta_gaps(a.reference)[1] if a.reference else None
    regions = gio.read_regions_bed(a.regions) if a.regions else {}
    tracks, beds, labels, targets_length = [], [], [], None
    for path, track, label, sfx in ((a.hifi, TRACK_HIFI, 'HiFi', '_hifi'), (a.nano, TRACK_NANO, 'Nano', '_nano')):
        if not path:
            continue
        host = gio.read_depth_gz(path)
        targets_length = {k: len(v) for k, v in host.items()}
        dev = P.merge_gaps_depths(P._adopt(host, ses, track), Ns_bed or None)
        tracks.append(dev); labels.append(label)
        both = a.hifi and a.nano
        beds.append(P.merge_depth(dev, a.prefix + (sfx if both else ''), a.threshold, a.flank_len, a.directory, a.force, label))
    if a.hifi and a.nano:
        ses.scan_hint = None
        merged = P.merge_two_type_depth(tracks[0], tracks[1], a.prefix + '_two_type', a.directory, a.force, 1)
        merged = P.merge_gaps_depths(merged, Ns_bed or None)
        tracks.append(merged); labels.append('HiFi + Nano')
        beds.append(P.merge_depth(merged, a.prefix + '_two_type', a.threshold, a.flank_len, a.directory, a.force, 'two_types'))
    P.compute_index(targets_length, a.prefix, a.directory, a.force, beds, labels, a.flank_len, a.dist_percent, regions,
                    tracks, a.threshold, [], session=ses)


if __name__ == '__main__':
    main()
