#!/usr/bin/env python
"""Drop-in command line for the B200 hot path: same flags, dests, defaults, option groups, version string and
exit texts as the reference's `GCI.py` (yeeus/GCI, GCI.py:1031-1113), so existing invocations keep working.
The work is done by `gci_b200.pipeline.GCI` through libgci_cuda.so.  `-p` (plotting) is outside the hot path.
"""
import argparse
import os
import sys

sys.path.insert(0, os.path.dirname(os.path.abspath(__file__)))

VERSION = 'GCI version 1.0'
HELP_HINT = 'Please read the help message use "-h" or "--help"'

# The command-line surface as data: (option strings, argparse keywords) per help group.  Texts and defaults
# are the reference's (they are what `-h` prints and what `GCI(**vars(args))` receives).
_INT, _FLOAT = dict(metavar='INT', type=int), dict(metavar='FLOAT', type=float)
_SWITCH = dict(action='store_const', const=True, default=False)
CLI = {
    "Input/Output": [
        (('-r', '--reference'), dict(metavar='FILE', help='The reference file')),
        (('--hifi',), dict(nargs='+', metavar='', help='PacBio HiFi reads alignment files (at least one bam file)')),
        (('--nano',), dict(nargs='+', metavar='', help='Oxford Nanopore long reads alignment files (at least one bam file)')),
        (('--chrs',), dict(metavar='', help='A list of chromosomes separated by comma')),
        (('-R', '--regions'), dict(metavar='FILE', help='Bed file containing regions\nBe cautious! If both specify `--chrs` and `--regions`, chromosomes in regions bed file should be included in the chromosomes list')),
        (('-ts', '--threshold'), dict(_INT, default=0, help='The threshold of depth to be reported as issues [0]')),
        (('-dp', '--dist-percent'), dict(_FLOAT, default=0.005, help='The distance between the candidate gap intervals for combining in chromosome units [0.005]')),
        (('-t', '--threads'), dict(_INT, default=1, help='Number of threads [1]')),
        (('-d',), dict(dest='directory', metavar='PATH', default='.', help='The directory of output files [.]')),
        (('-o', '--output'), dict(dest='prefix', metavar='STR', default='GCI', help='Prefix of output files [GCI]')),
    ],
    "Filter Options": [
        (('-mq', '--map-qual'), dict(_INT, default=30, help='Minium mapping quality for alignments [30]')),
        (('--mq-cutoff',), dict(_INT, default=50, help='The cutoff of mapping quality for keeping the alignment [50]\n(only used when inputting more than one alignment files)')),
        (('-ip', '--iden-percent'), dict(_FLOAT, default=0.9, help='Minimum identity (num_match_res/len_aln) of alignments [0.9]')),
        (('-op', '--ovlp-percent'), dict(_FLOAT, default=0.9, help='Minimum overlapping percentage of the same read alignment if inputting more than one alignment files [0.9]')),
        (('-cp', '--clip-percent'), dict(_FLOAT, default=0.1, help='Maximum clipped percentage of the alignment [0.1]')),
        (('-fl', '--flank-len'), dict(_INT, default=15, help='The flanking length of the clipped bases [15]')),
    ],
    "Plot Options": [
        (('-p', '--plot'), dict(_SWITCH, help='Visualize the finally filtered whole genome (and regions if providing the option `-R`) depth [False]\n(not provided by the GPU build: a warning is printed, every other output is written)')),
        (('-dmin', '--depth-min'), dict(_FLOAT, default=0.1, help='Minimum depth in folds of mean coverage for plotting [0.1]')),
        (('-dmax', '--depth-max'), dict(_FLOAT, default=4.0, help='Maximum depth in folds of mean coverage for plotting [4.0]')),
        (('-ws', '--window-size'), dict(_INT, default=50000, help='The window size when plotting [50000]')),
        (('-it', '--image-type'), dict(metavar='STR', default='png', help='The format of the output images: png or pdf [png]')),
    ],
    "Other Options": [
        (('-f', '--force'), dict(_SWITCH, help='Force rewriting of existing files [False]')),
        (('-h', '--help'), dict(action='help', help='Show this help message and exit')),
        (('-v', '--version'), dict(action='version', version=VERSION, help="Show program's version number and exit")),
    ],
}


def build_parser(prog=None):
    parser = argparse.ArgumentParser(prog=prog or sys.argv[0], add_help=False,
                                     formatter_class=argparse.RawTextHelpFormatter,
                                     description='A program for assessing the T2T genome',
                                     epilog='Examples:\npython GCI.py -r ref.fa --hifi hifi.bam hifi.paf ... --nano nano.bam nano.paf ...')
    for title, options in CLI.items():
        group = parser.add_argument_group(title)
        for flags, kwargs in options:
            group.add_argument(*flags, **kwargs)
    return parser


def _readable(path):
    return os.path.exists(path) and os.access(path, os.R_OK)


def check_inputs(args):
    """The reference's input validation (GCI.py:1076-1110) with its messages."""
    read_types = (('hifi', 'PacBio HiFi reads'), ('nano', 'Oxford Nanopore long reads'))
    if all(args[key] is None for key, _ in read_types):
        sys.exit('ERROR!!! Please input at least one type of TGS reads alignment files (PacBio HiFi and/or Oxford Nanopore long reads)\n' + HELP_HINT)
    for key, label in read_types:
        files = args[key]
        if files is None:
            continue
        # the first unavailable file in argument order is reported before the bam count is looked at
        for f in files:
            if not _readable(f):
                sys.exit(f'ERROR!!! "{f}" is not an available file')
        if not any(f.endswith('.bam') for f in files):
            sys.exit(f'ERROR!!! Please input at least one {label} bam file\n' + HELP_HINT)
    if args['reference'] is None:
        sys.exit('ERROR!!! Please input the reference file\n' + HELP_HINT)
    if not _readable(args['reference']):
        sys.exit(f'ERROR!!! "{args["reference"]}" is not an available file')
    if args['map_qual'] > args['mq_cutoff']:
        print(f'WARNING!!! The minium mapping quality ({args["map_qual"]}) is higher than the cutoff ({args["mq_cutoff"]}), which means that wouldn\'t filter any reads\n' + HELP_HINT, file=sys.stderr)


def main(argv=None):
    parser = build_parser()
    argv = sys.argv[1:] if argv is None else argv
    args = vars(parser.parse_args(argv))
    if not argv:
        parser.print_help()
        sys.exit()
    check_inputs(args)
    print(f'Used arguments:{args}')
    from gci_b200.pipeline import GCI
    GCI(**args)


if __name__ == '__main__':
    main()
