"""Command line and driver with the reference's surface (AmpliPy.py:113-171 parse_args, 774-963
run_amplipy, 966-1025 dispatch): subcommands trim / variants / consensus / aio with the same flags,
defaults, validation messages and log lines.  The per-read and per-position loops run on the GPU
through the C ABI (engine.Engine); there is no CPU path."""
import argparse
import gzip
import os
import sys
from datetime import datetime

from . import alnio, calling, vcf
from .primers import InputError, find_overlapping_primers, load_primers, max_primer_len

VERSION = "0.0.2"
DESCRIPTION = "\nAmpliPy: Python toolkit for viral amplicon sequencing\n"

# default arguments (AmpliPy.py:22-30)
DEFAULT_MIN_DEPTH_CONSENSUS = 10
DEFAULT_MIN_DEPTH_VARIANTS = 1
DEFAULT_MIN_FREQ_CONSENSUS = 0
DEFAULT_MIN_FREQ_VARIANTS = 0.03
DEFAULT_MIN_LENGTH = 30
DEFAULT_MIN_QUALITY = 20
DEFAULT_PRIMER_POS_OFFSET = 0
DEFAULT_SLIDING_WINDOW_WIDTH = 4
DEFAULT_UNKNOWN_SYMBOL = 'N'
PROGRESS_NUM_READS = 50000          # AmpliPy.py:19

# messages (AmpliPy.py:47-78)
ERROR_TEXT_FILE_NOT_FOUND = "File not found"
ERROR_TEXT_INVALID_FASTA = "Invalid FASTA file"
ERROR_TEXT_INVALID_MIN_DEPTH = "Minimum depth must be positive"
ERROR_TEXT_INVALID_MIN_FREQ = "Minimum frequency must be between 0 and 1"
ERROR_TEXT_INVALID_MIN_LENGTH = "Minimum length must be >= 1"
ERROR_TEXT_INVALID_SLIDING_WINDOW_WIDTH = "Sliding window width must be >= 1"
ERROR_TEXT_INVALID_UNKNOWN_SYMBOL_LENGTH = "Unknown symbol must be exactly 1 character"
ERROR_TEXT_MULTIPLE_REF_SEQS = "Multiple sequences in FASTA file"
ERROR_TEXT_NEGATIVE_MIN_QUALITY = "Minimum quality must be non-negative"
ERROR_TEXT_NEGATIVE_PRIMER_POS_OFFSET = "Primer position offset must be non-negative"
HELP_TEXT_CONSENSUS = "Consensus Sequence (FASTA)"
HELP_TEXT_MIN_DEPTH_CONSENSUS = "Minimum depth to call consensus"
HELP_TEXT_MIN_DEPTH_VARIANTS = "Minimum depth to call variant"
HELP_TEXT_MIN_FREQ_CONSENSUS = "Minimum frequency threshold (0-1) to call consensus"
HELP_TEXT_MIN_FREQ_VARIANTS = "Minimum frequency threshold (0-1) to call variant"
HELP_TEXT_MIN_QUAL = "Minimum quality threshold"
HELP_TEXT_PRIMER = "Primer File (BED)"
HELP_TEXT_READS_UNTRIMMED = "Untrimmed Reads (SAM/BAM)"
HELP_TEXT_READS_TRIMMED = "Trimmed Reads (SAM/BAM)"
HELP_TEXT_REFERENCE = "Reference Genome (FASTA)"
HELP_TEXT_TRIM_INCLUDE_READS_NO_PRIMER = "Include reads with no primers"
HELP_TEXT_TRIM_MIN_LENGTH = "Minimum length of read to retain after trimming"
HELP_TEXT_TRIM_PRIMER_POS_OFFSET = ("Primer position offset. Reads that occur at the specified offset positions relative to "
                                    "primer positions will also be trimmed")
HELP_TEXT_TRIM_SLIDING_WINDOW_WIDTH = "Width of sliding window (average quality of this window must be >= minimum quality threshold)"
HELP_TEXT_UNKNOWN_SYMBOL = "Character to print in regions with less than minimum coverage"
HELP_TEXT_VARIANTS = "Variant Calls (VCF)"


def print_log(s='', end='\n'):
    print("[%s] %s" % (datetime.now().strftime("%Y-%m-%d %H:%M:%S"), s), end=end, file=sys.stderr)
    sys.stderr.flush()


def error(s=None):
    print_log("ERROR" if s is None else "ERROR: %s" % s)
    sys.exit(1)


def parse_args(argv=None):
    argv = list(sys.argv[1:] if argv is None else argv)
    if len(argv) == 0:
        argv.append('-h')
    F = argparse.ArgumentDefaultsHelpFormatter
    parser = argparse.ArgumentParser(description=DESCRIPTION, formatter_class=F)
    sub = parser.add_subparsers(dest='command')

    t = sub.add_parser("trim", description=DESCRIPTION, formatter_class=F)
    t.add_argument('-i', '--input', required=False, type=str, default='stdin', help=HELP_TEXT_READS_UNTRIMMED)
    t.add_argument('-p', '--primer', required=True, type=str, help=HELP_TEXT_PRIMER)
    t.add_argument('-r', '--reference', required=True, type=str, help=HELP_TEXT_REFERENCE)
    t.add_argument('-o', '--output', required=False, type=str, default='stdout', help=HELP_TEXT_READS_TRIMMED)
    t.add_argument('-x', '--primer_pos_offset', required=False, type=int, default=DEFAULT_PRIMER_POS_OFFSET, help=HELP_TEXT_TRIM_PRIMER_POS_OFFSET)
    t.add_argument('-ml', '--min_length', required=False, type=int, default=DEFAULT_MIN_LENGTH, help=HELP_TEXT_TRIM_MIN_LENGTH)
    t.add_argument('-mq', '--min_quality', required=False, type=int, default=DEFAULT_MIN_QUALITY, help=HELP_TEXT_MIN_QUAL)
    t.add_argument('-s', '--sliding_window_width', required=False, type=int, default=DEFAULT_SLIDING_WINDOW_WIDTH, help=HELP_TEXT_TRIM_SLIDING_WINDOW_WIDTH)
    t.add_argument('-e', '--include_no_primer', action='store_true', help=HELP_TEXT_TRIM_INCLUDE_READS_NO_PRIMER)

    v = sub.add_parser("variants", description=DESCRIPTION, formatter_class=F)
    v.add_argument('-i', '--input', required=False, type=str, default='stdin', help=HELP_TEXT_READS_TRIMMED)
    v.add_argument('-r', '--reference', required=True, type=str, help=HELP_TEXT_REFERENCE)
    v.add_argument('-o', '--output', required=False, type=str, default='stdout', help=HELP_TEXT_VARIANTS)
    v.add_argument('-mq', '--min_quality', required=False, type=int, default=DEFAULT_MIN_QUALITY, help=HELP_TEXT_MIN_QUAL)
    v.add_argument('-mf', '--min_freq', required=False, type=float, default=DEFAULT_MIN_FREQ_VARIANTS, help=HELP_TEXT_MIN_FREQ_VARIANTS)
    v.add_argument('-md', '--min_depth', required=False, type=int, default=DEFAULT_MIN_DEPTH_VARIANTS, help=HELP_TEXT_MIN_DEPTH_VARIANTS)

    c = sub.add_parser("consensus", description=DESCRIPTION, formatter_class=F)
    c.add_argument('-i', '--input', required=False, type=str, default='stdin', help=HELP_TEXT_READS_TRIMMED)
    c.add_argument('-r', '--reference', required=True, type=str, help=HELP_TEXT_REFERENCE)
    c.add_argument('-o', '--output', required=False, type=str, default='stdout', help=HELP_TEXT_CONSENSUS)
    c.add_argument('-mq', '--min_quality', required=False, type=int, default=DEFAULT_MIN_QUALITY, help=HELP_TEXT_MIN_QUAL)
    c.add_argument('-mf', '--min_freq', required=False, type=float, default=DEFAULT_MIN_FREQ_CONSENSUS, help=HELP_TEXT_MIN_FREQ_CONSENSUS)
    c.add_argument('-md', '--min_depth', required=False, type=int, default=DEFAULT_MIN_DEPTH_CONSENSUS, help=HELP_TEXT_MIN_DEPTH_CONSENSUS)
    c.add_argument('-n', '--unknown_symbol', required=False, type=str, default=DEFAULT_UNKNOWN_SYMBOL, help=HELP_TEXT_UNKNOWN_SYMBOL)

    a = sub.add_parser("aio", description=DESCRIPTION, formatter_class=F)
    a.add_argument('-i', '--input', required=False, type=str, default='stdin', help=HELP_TEXT_READS_UNTRIMMED)
    a.add_argument('-p', '--primer', required=True, type=str, help=HELP_TEXT_PRIMER)
    a.add_argument('-r', '--reference', required=True, type=str, help=HELP_TEXT_REFERENCE)
    a.add_argument('-ot', '--output_trimmed_reads', required=True, type=str, help=HELP_TEXT_READS_TRIMMED)
    a.add_argument('-ov', '--output_variants', required=True, type=str, help=HELP_TEXT_VARIANTS)
    a.add_argument('-oc', '--output_consensus', required=True, type=str, help=HELP_TEXT_CONSENSUS)
    a.add_argument('-x', '--primer_pos_offset', required=False, type=int, default=DEFAULT_PRIMER_POS_OFFSET, help=HELP_TEXT_TRIM_PRIMER_POS_OFFSET)
    a.add_argument('-ml', '--min_length', required=False, type=int, default=DEFAULT_MIN_LENGTH, help=HELP_TEXT_TRIM_MIN_LENGTH)
    a.add_argument('-mq', '--min_quality', required=False, type=int, default=DEFAULT_MIN_QUALITY, help=HELP_TEXT_MIN_QUAL)
    a.add_argument('-s', '--sliding_window_width', required=False, type=int, default=DEFAULT_SLIDING_WINDOW_WIDTH, help=HELP_TEXT_TRIM_SLIDING_WINDOW_WIDTH)
    a.add_argument('-mfc', '--min_freq_consensus', required=False, type=float, default=DEFAULT_MIN_FREQ_CONSENSUS, help=HELP_TEXT_MIN_FREQ_CONSENSUS)
    a.add_argument('-mfv', '--min_freq_variants', required=False, type=float, default=DEFAULT_MIN_FREQ_VARIANTS, help=HELP_TEXT_MIN_FREQ_VARIANTS)
    a.add_argument('-mdc', '--min_depth_consensus', required=False, type=int, default=DEFAULT_MIN_DEPTH_CONSENSUS, help=HELP_TEXT_MIN_DEPTH_CONSENSUS)
    a.add_argument('-mdv', '--min_depth_variants', required=False, type=int, default=DEFAULT_MIN_DEPTH_VARIANTS, help=HELP_TEXT_MIN_DEPTH_VARIANTS)
    a.add_argument('-n', '--unknown_symbol', required=False, type=str, default=DEFAULT_UNKNOWN_SYMBOL, help=HELP_TEXT_UNKNOWN_SYMBOL)
    a.add_argument('-e', '--include_no_primer', action='store_true', help=HELP_TEXT_TRIM_INCLUDE_READS_NO_PRIMER)
    return parser.parse_args(argv)


def load_ref_genome(reference_fn):
    """Single-record FASTA -> (ID, sequence) (AmpliPy.py:212-232)."""
    if not os.path.isfile(reference_fn):
        raise InputError("%s: %s" % (ERROR_TEXT_FILE_NOT_FOUND, reference_fn))
    with open(reference_fn, 'r') as f:
        lines = f.read().strip().splitlines()
    if len(lines) < 2 or not lines[0].startswith('>'):
        raise InputError("%s: %s" % (ERROR_TEXT_INVALID_FASTA, reference_fn))
    ref_id = lines[0][1:].split()[0].strip()
    seq = ''.join(lines[1:])
    if '>' in seq:
        raise InputError("%s: %s" % (ERROR_TEXT_MULTIPLE_REF_SEQS, reference_fn))
    return ref_id, seq


def run_amplipy(untrimmed_reads_fn=None, primer_fn=None, reference_fn=None, trimmed_reads_fn=None, variants_fn=None,
                consensus_fn=None, primer_pos_offset=None, min_length=None, min_quality=None, sliding_window_width=None,
                min_freq_consensus=None, min_freq_variants=None, min_depth_consensus=None, min_depth_variants=None,
                unknown_symbol=None, include_no_primer=None, run_trim=False, run_variants=False, run_consensus=False,
                argv=None, device=0):
    """Same keyword arguments and behaviour as the reference's run_amplipy (AmpliPy.py:774-963)."""
    argv = list(sys.argv if argv is None else argv)
    try:
        return _run(untrimmed_reads_fn, primer_fn, reference_fn, trimmed_reads_fn, variants_fn, consensus_fn, primer_pos_offset,
                    min_length, min_quality, sliding_window_width, min_freq_consensus, min_freq_variants, min_depth_consensus,
                    min_depth_variants, unknown_symbol, include_no_primer, run_trim, run_variants, run_consensus, argv, device)
    except InputError as e:
        error(str(e))


def _run(untrimmed_reads_fn, primer_fn, reference_fn, trimmed_reads_fn, variants_fn, consensus_fn, primer_pos_offset, min_length,
         min_quality, sliding_window_width, min_freq_consensus, min_freq_variants, min_depth_consensus, min_depth_variants,
         unknown_symbol, include_no_primer, run_trim, run_variants, run_consensus, argv, device):
    # validity checks (AmpliPy.py:837-854)
    if primer_pos_offset is not None and primer_pos_offset < 0:
        error("%s: %s" % (ERROR_TEXT_NEGATIVE_PRIMER_POS_OFFSET, primer_pos_offset))
    if min_length is not None and min_length < 1:
        error("%s: %s" % (ERROR_TEXT_INVALID_MIN_LENGTH, min_length))
    if min_quality is not None and min_quality < 0:
        error("%s: %s" % (ERROR_TEXT_NEGATIVE_MIN_QUALITY, min_quality))
    if sliding_window_width is not None and sliding_window_width < 1:
        error("%s: %s" % (ERROR_TEXT_INVALID_SLIDING_WINDOW_WIDTH, sliding_window_width))
    if min_freq_consensus is not None and (min_freq_consensus < 0 or min_freq_consensus > 1):
        error("%s: %s" % (ERROR_TEXT_INVALID_MIN_FREQ, min_freq_consensus))
    if min_freq_variants is not None and (min_freq_variants < 0 or min_freq_variants > 1):
        error("%s: %s" % (ERROR_TEXT_INVALID_MIN_FREQ, min_freq_variants))
    if min_depth_consensus is not None and min_depth_consensus < 0:
        error("%s: %s" % (ERROR_TEXT_INVALID_MIN_DEPTH, min_depth_consensus))
    if min_depth_variants is not None and min_depth_variants < 0:
        error("%s: %s" % (ERROR_TEXT_INVALID_MIN_DEPTH, min_depth_variants))
    if unknown_symbol is not None and len(unknown_symbol) != 1:
        error("%s: %s" % (ERROR_TEXT_INVALID_UNKNOWN_SYMBOL_LENGTH, unknown_symbol))

    # mode banner (AmpliPy.py:857-866)
    if not (run_trim or run_variants or run_consensus):
        error("Not running any of the AmpliPy operations")
    if run_trim and not (run_variants or run_consensus):
        print_log("Executing AmpliPy Trim (v%s)" % VERSION)
    elif run_variants and not (run_trim or run_consensus):
        print_log("Executing AmpliPy Variants (v%s)" % VERSION)
    elif run_consensus and not (run_trim or run_variants):
        print_log("Executing AmpliPy Consensus (v%s)" % VERSION)
    else:
        print_log("Executing AmpliPy All-In-One (v%s)" % VERSION)

    print_log("Loading reference genome: %s" % reference_fn)
    ref_id, ref_seq = load_ref_genome(reference_fn)
    L = len(ref_seq)
    tables, mpl = None, 0
    if primer_fn is not None:
        print_log("Loading primers: %s" % primer_fn)
        primers = load_primers(primer_fn)
        mpl = max_primer_len(primers)
        print_log("Precalculating overlapping primers...")
        tables = find_overlapping_primers(L, primers, primer_pos_offset)
    in_fn = untrimmed_reads_fn if run_trim else trimmed_reads_fn
    if run_trim:
        print_log("Input untrimmed SAM/BAM: %s" % untrimmed_reads_fn)
        print_log("Output trimmed SAM/BAM: %s" % trimmed_reads_fn)
        alnio.check_output_path(trimmed_reads_fn)
    else:
        print_log("Input trimmed SAM/BAM: %s" % trimmed_reads_fn)
    if variants_fn is not None:
        print_log("Output variants VCF: %s" % variants_fn)
        vcf.check_output_path(variants_fn)
    import threading
    import time as _time
    tm = {}
    from .engine import AmpError, Engine, TrimResult

    def make_engine(indel_rich):
        return Engine(ref_len=L, primer_tables=tables, max_primer_len=mpl,
                      min_quality=DEFAULT_MIN_QUALITY if min_quality is None else min_quality,
                      sliding_window_width=DEFAULT_SLIDING_WINDOW_WIDTH if sliding_window_width is None else sliding_window_width,
                      min_length=DEFAULT_MIN_LENGTH if min_length is None else min_length,
                      include_no_primer=bool(include_no_primer), device=device,
                      ins_slots=(1 << 24) if indel_rich else 0, ins_arena_bytes=(1 << 30) if indel_rich else 0)
    pile = run_variants or run_consensus
    _t = _time.perf_counter()
    aln = eng = trim = None
    n_reads = 0
    if (in_fn.lower().endswith(".bam") and os.path.isfile(in_fn) and not os.environ.get("AMPLIPY_HOST_DECODE") and
            hasattr(Engine, "decode_bam")):
        # BAM: the compressed file goes to the GPU as it is and is decoded there (amp_bam_decode_host); the host inflates it too,
        # at the same time, only when the records are needed for the trimmed output
        with open(in_fn, "rb") as f:
            raw = f.read()
        layout = alnio.bam_layout(raw)
        box = {}
        host = None
        # BAM in -> BAM out: the trimmed records are rebuilt and compressed on the device as well (amp_decoded_write_bam), unless a
        # zlib level is asked for (AMPLIPY_BAM_LEVEL; 6 = what htslib writes)
        dev_bam_out = (run_trim and trimmed_reads_fn.lower().endswith(".bam") and "AMPLIPY_BAM_LEVEL" not in os.environ and
                       hasattr(Engine, "decoded_write_bam"))
        if run_trim and not dev_bam_out:
            def _host_decode():
                try:
                    box["aln"] = alnio._read_bam(raw)
                except BaseException as e:
                    box["err"] = e
            host = threading.Thread(target=_host_decode)
            host.start()
        out_header = alnio.header_with_pg(layout["header_text"], argv) if run_trim else None
        eng = make_engine(layout["ops_per_read"] > 8)
        print_log("Processing reads...")
        try:
            info = eng.decode_bam(raw, layout)
            outs = eng.process_decoded(trim=run_trim, pileup=pile, download=not dev_bam_out)
            n_reads = info["n"]
        except AmpError:
            eng.close()
            eng = None                       # e.g. records straddling BGZF blocks (not written by htslib): host decoder below
            dev_bam_out = False
        tm["gpu_decode_process"] = _time.perf_counter() - _t
        if host is not None:
            host.join()
            if "err" in box:
                raise box["err"]
            aln = box["aln"]
            if eng is not None:
                trim = TrimResult(aln.batch, *outs)
        tm["decode"] = _time.perf_counter() - _t
        if eng is not None:
            eng.raise_on_device_errors()
    else:
        dev_bam_out = False
    if eng is None:
        aln = aln or alnio.read_alignments(in_fn)
        tm["decode"] = _time.perf_counter() - _t
        out_header = alnio.header_with_pg(aln.header_text, argv) if run_trim else None
        indel_rich = aln.n and (aln.batch.cigar.size / aln.n) > 8
        eng = make_engine(indel_rich)
        print_log("Processing reads...")
        _t = _time.perf_counter()
        trim = eng.process(aln.batch, trim=run_trim, pileup=pile)
        eng.raise_on_device_errors()
        tm["gpu_process"] = _time.perf_counter() - _t
        n_reads = aln.n
    # the reference reports its progress every PROGRESS_NUM_READS reads of its loop (AmpliPy.py:19, 898-899); the reads
    # are processed as one batch here, so the same lines are written once the batch is through
    for k in range(PROGRESS_NUM_READS, n_reads, PROGRESS_NUM_READS):
        print_log("Processed %d reads..." % k)
    writer = None
    if run_trim:
        # the trimmed reads are encoded (record rewrite + BGZF deflate, both outside the GIL) while calling and the text outputs run
        def _write():
            t0 = _time.perf_counter()
            try:
                if dev_bam_out:
                    data, _ = eng.decoded_write_bam(alnio._bam_header_bytes(out_header, layout["refs"]))
                    with open(trimmed_reads_fn, "wb") as f:
                        f.write(data)
                else:
                    # BGZF deflate on the GPU unless a zlib level is asked for
                    on_gpu = "AMPLIPY_BAM_LEVEL" not in os.environ and hasattr(eng, "bgzf_deflate")
                    alnio.write_alignments(trimmed_reads_fn, aln, out_header, trim, deflater=eng.bgzf_deflate if on_gpu else None)
            except BaseException as e:      # re-raised on the main thread
                tm["_write_error"] = e
            tm["encode_bam"] = _time.perf_counter() - t0
        writer = threading.Thread(target=_write)
        writer.start()
    _t = _time.perf_counter()
    if pile:
        counts = eng.counts()
        ins = eng.insertions()
        res = eng.call(ref_seq,
                       DEFAULT_MIN_DEPTH_CONSENSUS if min_depth_consensus is None else min_depth_consensus,
                       DEFAULT_MIN_FREQ_CONSENSUS if min_freq_consensus is None else min_freq_consensus,
                       DEFAULT_MIN_DEPTH_VARIANTS if min_depth_variants is None else min_depth_variants,
                       DEFAULT_MIN_FREQ_VARIANTS if min_freq_variants is None else min_freq_variants)
        if run_variants:
            vcf.write_vcf(variants_fn, ref_id, VERSION, argv, calling.variant_records(res, ins, ref_seq, counts))
        if run_consensus:
            text = '>sample\n%s\n' % calling.consensus_string(res, ins, 0, unknown_symbol)
            if consensus_fn.lower() == 'stdout':
                sys.stdout.write(text)
            elif consensus_fn.lower().endswith('.gz'):
                with gzip.open(consensus_fn, 'wt') as f:
                    f.write(text)
            else:
                with open(consensus_fn, 'w') as f:
                    f.write(text)
    tm["call_and_text"] = _time.perf_counter() - _t
    if writer is not None:
        writer.join()
        if "_write_error" in tm:
            raise tm.pop("_write_error")
    if os.environ.get("AMP_CLI_TIMINGS"):      # bench.py's file-to-file leg: where the wall time went
        import json
        with open(os.environ["AMP_CLI_TIMINGS"], "w") as f:
            json.dump({k: round(v, 4) for k, v in tm.items()}, f)
    print_log("Finished Processing %d reads" % max(n_reads - 1, 0))
    return eng


def main(argv=None):
    args = parse_args(argv)
    full_argv = list(sys.argv) if argv is None else ["AmpliPy.py"] + list(argv)
    if args.command == 'trim':
        run_amplipy(untrimmed_reads_fn=args.input, primer_fn=args.primer, reference_fn=args.reference,
                    trimmed_reads_fn=args.output, primer_pos_offset=args.primer_pos_offset, min_length=args.min_length,
                    min_quality=args.min_quality, sliding_window_width=args.sliding_window_width,
                    include_no_primer=args.include_no_primer, run_trim=True, argv=full_argv)
    elif args.command == 'variants':
        run_amplipy(trimmed_reads_fn=args.input, reference_fn=args.reference, variants_fn=args.output,
                    min_quality=args.min_quality, min_freq_variants=args.min_freq, min_depth_variants=args.min_depth,
                    run_variants=True, argv=full_argv)
    elif args.command == 'consensus':
        run_amplipy(trimmed_reads_fn=args.input, reference_fn=args.reference, consensus_fn=args.output,
                    min_quality=args.min_quality, min_freq_consensus=args.min_freq, min_depth_consensus=args.min_depth,
                    unknown_symbol=args.unknown_symbol, run_consensus=True, argv=full_argv)
    elif args.command == 'aio':
        run_amplipy(untrimmed_reads_fn=args.input, primer_fn=args.primer, reference_fn=args.reference,
                    trimmed_reads_fn=args.output_trimmed_reads, variants_fn=args.output_variants,
                    consensus_fn=args.output_consensus, primer_pos_offset=args.primer_pos_offset, min_length=args.min_length,
                    min_quality=args.min_quality, sliding_window_width=args.sliding_window_width,
                    min_freq_consensus=args.min_freq_consensus, min_freq_variants=args.min_freq_variants,
                    min_depth_consensus=args.min_depth_consensus, min_depth_variants=args.min_depth_variants,
                    unknown_symbol=args.unknown_symbol, include_no_primer=args.include_no_primer, run_trim=True,
                    run_variants=True, run_consensus=True, argv=full_argv)


if __name__ == "__main__":
    main()
