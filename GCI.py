#!/usr/bin/env python
"""Drop-in command line of the reference `GCI.py` (yeeus/GCI, GCI.py:1031-1113) on the B200 hot path.

Same flags, dests, defaults, groups, version string and exit texts; the work is done by
`gci_b200.pipeline.GCI` through libgci_cuda.so.  `-p` (plotting) is outside the hot path.
"""
import argparse
import os
import sys

sys.path.insert(0, os.path.dirname(os.path.abspath(__file__)))


def build_parser(prog=None):
    version = 'GCI version 1.0'
    parser = argparse.ArgumentParser(prog=prog or sys.argv[0], add_help=False,
                                     formatter_class=argparse.RawTextHelpFormatter,
                                     description='A program for assessing the T2T genome',
                                     epilog='Examples:\npython GCI.py -r ref.fa --hifi hifi.bam hifi.paf ... --nano nano.bam nano.paf ...')
    io = parser.add_argument_group("Input/Output")
    io.add_argument('-r', '--reference', metavar='FILE', help='The reference file')
    io.add_argument('--hifi', nargs='+', metavar='', help='PacBio HiFi reads alignment files (at least one bam file)')
    io.add_argument('--nano', nargs='+', metavar='', help='Oxford Nanopore long reads alignment files (at least one bam file)')
    io.add_argument('--chrs', metavar='', help='A list of chromosomes separated by comma')
    io.add_argument('-R', '--regions', metavar='FILE', help='Bed file containing regions\nBe cautious! If both specify `--chrs` and `--regions`, chromosomes in regions bed file should be included in the chromosomes list')
    io.add_argument('-ts', '--threshold', metavar='INT', type=int, help='The threshold of depth to be reported as issues [0]', default=0)
    io.add_argument('-dp', '--dist-percent', metavar='FLOAT', type=float, help='The distance between the candidate gap intervals for combining in chromosome units [0.005]', default=0.005)
    io.add_argument('-t', '--threads', metavar='INT', type=int, help='Number of threads [1]', default=1)
    io.add_argument('-d', dest='directory', metavar='PATH', help='The directory of output files [.]', default='.')
    io.add_argument('-o', '--output', dest='prefix', metavar='STR', help='Prefix of output files [GCI]', default='GCI')

    fo = parser.add_argument_group("Filter Options")
    fo.add_argument('-mq', '--map-qual', metavar='INT', type=int, help='Minium mapping quality for alignments [30]', default=30)
    fo.add_argument('--mq-cutoff', metavar='INT', type=int, help='The cutoff of mapping quality for keeping the alignment [50]\n(only used when inputting more than one alignment files)', default=50)
    fo.add_argument('-ip', '--iden-percent', metavar='FLOAT', type=float, help='Minimum identity (num_match_res/len_aln) of alignments [0.9]', default=0.9)
    fo.add_argument('-op', '--ovlp-percent', metavar='FLOAT', type=float, help='Minimum overlapping percentage of the same read alignment if inputting more than one alignment files [0.9]', default=0.9)
    fo.add_argument('-cp', '--clip-percent', metavar='FLOAT', type=float, help='Maximum clipped percentage of the alignment [0.1]', default=0.1)
    fo.add_argument('-fl', '--flank-len', metavar='INT', type=int, help='The flanking length of the clipped bases [15]', default=15)

    po = parser.add_argument_group("Plot Options")
    po.add_argument('-p', '--plot', action='store_const', help='Visualize the finally filtered whole genome (and regions if providing the option `-R`) depth [False]', const=True, default=False)
    po.add_argument('-dmin', '--depth-min', metavar='FLOAT', type=float, help='Minimum depth in folds of mean coverage for plotting [0.1]', default=0.1)
    po.add_argument('-dmax', '--depth-max', metavar='FLOAT', type=float, help='Maximum depth in folds of mean coverage for plotting [4.0]', default=4.0)
    po.add_argument('-ws', '--window-size', metavar='INT', type=int, help='The window size when plotting [50000]', default=50000)
    po.add_argument('-it', '--image-type', metavar='STR', help='The format of the output images: png or pdf [png]', default='png')

    op = parser.add_argument_group("Other Options")
    op.add_argument('-f', '--force', action='store_const', help='Force rewriting of existing files [False]', const=True, default=False)
    op.add_argument('-h', '--help', action="help", help="Show this help message and exit")
    op.add_argument('-v', '--version', action="version", version=version, help="Show program's version number and exit")
    return parser


def check_inputs(args):
    """Input validation of GCI.py:1076-1110 (same messages)."""
    if (args['hifi'] == None) and (args['nano'] == None):
        sys.exit('ERROR!!! Please input at least one type of TGS reads alignment files (PacBio HiFi and/or Oxford Nanopore long reads)\nPlease read the help message use "-h" or "--help"')
    for key, label in (('hifi', 'PacBio HiFi reads'), ('nano', 'Oxford Nanopore long reads')):
        if args[key] != None:
            bam_num = 0
            for file in args[key]:
                if os.path.exists(file) and os.access(file, os.R_OK):
                    if file.endswith('.bam'):
                        bam_num += 1
                else:
                    sys.exit(f'ERROR!!! "{file}" is not an available file')
            if bam_num == 0:
                sys.exit(f'ERROR!!! Please input at least one {label} bam file\nPlease read the help message use "-h" or "--help"')
    if args['reference'] == None:
        sys.exit('ERROR!!! Please input the reference file\nPlease read the help message use "-h" or "--help"')
    elif not (os.path.exists(args['reference']) and os.access(args['reference'], os.R_OK)):
        sys.exit(f'ERROR!!! \"{args["reference"]}\" is not an available file')
    if args['map_qual'] > args['mq_cutoff']:
        print(f'WARNING!!! The minium mapping quality ({args["map_qual"]}) is higher than the cutoff ({args["mq_cutoff"]}), which means that wouldn\'t filter any reads\nPlease read the help message use "-h" or "--help"', file=sys.stderr)


def main(argv=None):
    parser = build_parser()
    argv = sys.argv[1:] if argv is None else argv
    args = vars(parser.parse_args(argv))
    if len(argv) == 0:
        parser.print_help()
        sys.exit()
    check_inputs(args)
    print(f'Used arguments:{args}')
    from gci_b200.pipeline import GCI
    GCI(**args)


if __name__ == '__main__':
    main()
