"""
snpmatch_b200 — B200-native genotype-matching hot path behind SNPmatch's `inbred` / `cross` interface.

The command line mirrors the reference's `snpmatch/__init__.py` for the two sub-commands on the
matching path (`inbred`, `cross`: flags of __init__.py:44-63), `parser` (:80-84) and the callers next to the
path that work on the resident panel (`pairsnp` :86-92, `simulate` :101-111, `genotype_cross`
:65-78 without its HMM mode); the other
sub-commands of the reference are outside this package's scope.
"""
import argparse
import logging
import os
import os.path
import sys

__version__ = '0.1.0'
__reference_version__ = '5.0.1'


def setLog(logDebug):
    log = logging.getLogger()
    level = logging.DEBUG if logDebug else logging.ERROR
    handler = logging.StreamHandler()
    handler.setLevel(level)
    handler.setFormatter(logging.Formatter("%(asctime)s - %(name)s - %(levelname)s - %(message)s"))
    log.setLevel(level)
    log.addHandler(handler)


def die(msg):
    sys.stderr.write('Error: ' + msg + '\n')
    sys.exit(1)


def check_file(inFile):
    if not inFile:
        die("file: %s not specified" % inFile)
    if not os.path.isfile(inFile):
        die("input file does not exist: " + inFile)


def snpmatch_inbred(args):
    check_file(args['inFile'])
    from .core import snpmatch
    snpmatch.potatoGenotyper(args)


def snpmatch_cross(args):
    check_file(args['inFile'])
    from .core import csmatch
    csmatch.potatoCrossIdentifier(args)


def snpmatch_parser(args):
    check_file(args['inFile'])
    if not args['outFile'] and os.path.isfile(args['inFile'] + ".snpmatch.npz"):
        os.remove(args['inFile'] + ".snpmatch.npz")
    from .core import parsers
    parsers.potatoParser(inFile=args['inFile'], logDebug=args['logDebug'], outFile=args['outFile'])


def snpmatch_paircomparions(args):
    check_file(args['inFile_1'])
    check_file(args['inFile_2'])
    from .core import snpmatch
    snpmatch.pairwiseScore(args['inFile_1'], args['inFile_2'], args['logDebug'], args['outFile'], args['hdf5File'])


def simulate_snps(args):
    from .core import simulate
    simulate.potatoSimulate(args)


def genotype_cross(args):
    if not args['parents']:
        die("parents not specified")
    from .core import genotype_cross as gtm
    gtm.potatoCrossGenotyper(args)


def get_options(description, version_message):
    p = argparse.ArgumentParser(description=description)
    p.add_argument('-V', '--version', action='version', version=version_message)
    sub = p.add_subparsers(title='subcommands', description='Choose a command to run', help='Following commands are supported')
    db_help = "Path to the SNP database: the reference's row-chunked hdf5 file (needs h5py) or a packed .npz written by Genotype.save_packed"
    inbred = sub.add_parser('inbred', help="SNPmatch on the inbred samples")
    inbred.add_argument("-i", "--input_file", dest="inFile", help="VCF/BED file for the variants in the sample")
    inbred.add_argument("-d", "--hdf5_file", default=None, dest="hdf5File", help=db_help)
    inbred.add_argument("-e", "--hdf5_acc_file", default=None, dest="hdf5accFile", help="Column-chunked hdf5 file of the reference (accepted for compatibility; one resident copy serves both)")
    inbred.add_argument("--refine", action="store_true", dest="refine", default=False, help="Refine scores for indistinguishable lines")
    inbred.add_argument("--skip_db_hets", action="store_true", dest="skip_db_hets", default=False, help="Replace heterozygous calls in DB with nan during the analysis.")
    inbred.add_argument("-v", "--verbose", action="store_true", dest="logDebug", default=False, help="Show verbose debugging output")
    inbred.add_argument("-o", "--output", dest="outFile", default="identify_inbred", help="Output file with the probability scores")
    inbred.set_defaults(func=snpmatch_inbred)

    cross = sub.add_parser('cross', help="SNPmatch on the crosses (F2s and F3s) of A. thaliana")
    cross.add_argument("-i", "--input_file", dest="inFile", help="VCF/BED file for the variants in the sample")
    cross.add_argument("-d", "--hdf5_file", default=None, dest="hdf5File", help=db_help)
    cross.add_argument("-e", "--hdf5_acc_file", default=None, dest="hdf5accFile", help="Column-chunked hdf5 file of the reference (accepted for compatibility)")
    cross.add_argument("-b", "--binLength", dest="binLen", help="Length of bins to calculate the likelihoods", default=300000, type=int)
    cross.add_argument("--genome", dest="genome", default="athaliana_tair10", help="Path to Reference JSON file, if you are working with non-thaliana tair10 assembly")
    cross.add_argument("--skip_db_hets", action="store_true", dest="skip_db_hets", default=False, help="Replace heterozygous calls in DB with nan during the analysis.")
    cross.add_argument("-v", "--verbose", action="store_true", dest="logDebug", default=False, help="Show verbose debugging output")
    cross.add_argument("-o", "--output", dest="outFile", default="identify_cross", help="Output files with the probability scores and scores along windows")
    cross.set_defaults(func=snpmatch_cross)

    parser = sub.add_parser('parser', help="parse the input file")
    parser.add_argument("-i", "--input_file", dest="inFile", help="VCF/BED file for the variants in the sample")
    parser.add_argument("-v", "--verbose", action="store_true", dest="logDebug", default=False, help="Show verbose debugging output")
    parser.add_argument("-o", "--output", dest="outFile", help="output + .npz file is generater required for SNPmatch")
    parser.set_defaults(func=snpmatch_parser)

    gc = sub.add_parser('genotype_cross', help="Genotype the crosses by windows given parents")
    gc.add_argument("-i", "--input_file", dest="inFile", help="VCF file for the variants in the sample")
    gc.add_argument("-d", "--hdf5_file", default=None, dest="hdf5File", help=db_help)
    gc.add_argument("-e", "--hdf5_acc_file", default=None, dest="hdf5accFile", help="Column-chunked hdf5 file of the reference (accepted for compatibility)")
    gc.add_argument("-p", "--parents", dest="parents", help="Parents for the cross, parent1 x parent2")
    gc.add_argument("-q", "--father", dest="father", help="VCF/BED file of parent 2 when the parents are given as files (then -p is the file of parent 1)")
    gc.add_argument("-b", "--binLength", dest="binLen", help="bin length", type=int, default=200000)
    gc.add_argument("--good_samples", dest="good_samples", help="accepted for compatibility (unused by the reference's window genotyper)", default=None)
    gc.add_argument("--lr_thres", dest="lr_thres", default=1.5, type=float, help="Likelihood ratio threshold for genotype calling.")
    gc.add_argument("--hmm", dest="hmm", action="store_true", help="HMM Viterbi genotyper of the reference: not part of this package")
    gc.add_argument("--genome", dest="genome", default="athaliana_tair10", help="Path to Reference JSON file, if you are working with non-thaliana tair10 assembly")
    gc.add_argument("-o", "--output", dest="outFile", default="genotype_cross", help="output file")
    gc.add_argument("-v", "--verbose", action="store_true", dest="logDebug", default=False, help="Show verbose debugging output")
    gc.set_defaults(func=genotype_cross)

    pair = sub.add_parser('pairsnp', help="pairwise comparison of two snp files")
    pair.add_argument("-i", "--input_file_1", dest="inFile_1", help="VCF/BED file for the variants in the sample one")
    pair.add_argument("-j", "--input_file_2", dest="inFile_2", help="VCF/BED file for the variants in the sample two")
    pair.add_argument("-d", "--hdf5_file", dest="hdf5File", default=None, help=db_help)
    pair.add_argument("-v", "--verbose", action="store_true", dest="logDebug", default=False, help="Show verbose debugging output")
    pair.add_argument("-o", "--output", dest="outFile", default="pairsnp", help="output json file")
    pair.set_defaults(func=snpmatch_paircomparions)

    sim = sub.add_parser('simulate', help="Given SNP database, check the genotyping efficiency randomly selecting 'n' number of SNPs")
    sim.add_argument("-d", "--hdf5_file", default=None, dest="hdf5File", help=db_help)
    sim.add_argument("-e", "--hdf5_acc_file", default=None, dest="hdf5accFile", help="Column-chunked hdf5 file of the reference (accepted for compatibility)")
    sim.add_argument("-a", "--ecotype_id", dest="AccID", help="Ecotype ID you want draw the SNPs")
    sim.add_argument("-n", "--number_of_snps", dest="numSNPs", help="number of SNPs to draw in random to genotype the sample", type=int)
    sim.add_argument("-p", "--error_rate", dest="err_rate", help="error rate while matching the SNPs, error rate of 0 gives perfect match to the accession", default=0.001, type=float)
    sim.add_argument("--f1", action="store_true", dest="simF1", default=False, help="Simulate SNPs for an F1, give parents as 1061x1062 in argument '-a'")
    sim.add_argument("--het_frac", default=1, type=float, dest="rm_het", help="For simulated F1s: fraction of segregating sites kept heterozygous; the rest become homozygous ref or alt with equal probability")
    sim.add_argument("-o", "--output", dest="outFile", help="Output file with scores")
    sim.add_argument("-v", "--verbose", action="store_true", dest="logDebug", default=False, help="Show verbose debugging output")
    sim.set_defaults(func=simulate_snps)
    return p


def main(argv=None):
    """Exit codes as in the reference (__init__.py:155-183): 0 ok, 2 on an exception."""
    version_message = '%%(prog)s v%s (B200 matching path of SNPmatch %s)' % (__version__, __reference_version__)
    parser = get_options("SNPmatch genotype matching on NVIDIA B200", version_message)
    args = vars(parser.parse_args(argv))
    setLog(args.get('logDebug', False))
    if 'func' not in args:
        parser.print_help()
        return 0
    try:
        args['func'](args)
        return 0
    except KeyboardInterrupt:
        return 0
    except Exception as e:
        logging.exception(e)
        return 2


if __name__ == '__main__':
    sys.exit(main())
