"""Flag strings of the reference workflow -> parameter structs (same meaning as the reference's
Parameters members; mm/commons/Parameters.cpp:422-439,871-892; src/commons/LocalParameters.h:96-102)."""
from common import parse_flags, multi
import oracle_binding as ob

U64MAX = (1 << 64) - 1


def km_fields(args, nucl):
    f = parse_flags(args)
    return dict(
        kmer_size=int(f["-k"]),
        alph_size=int(multi(f["--alph-size"], nucl)),
        kmers_per_seq=int(f["--kmer-per-seq"]),
        kmers_per_seq_scale=float(multi(f["--kmer-per-seq-scale"], nucl)),
        hash_shift=int(f["--hash-shift"]),
        include_only_extendable=int(f["--include-only-extendable"]),
        ignore_multi_kmer=int(f["--ignore-multi-kmer"]),
        cov_mode=int(f["--cov-mode"]),
        cov_thr=float(f["-c"]),
    )


def rs_fields(args):
    f = parse_flags(args)
    return dict(rescore_mode=int(f["--rescore-mode"]), seq_id_thr=float(f["--min-seq-id"]), eval_thr=float(f["-e"]),
                cov_mode=int(f["--cov-mode"]), cov_thr=float(f["-c"]), aln_len_thr=int(f["--min-aln-len"]),
                seq_id_mode=int(f["--seq-id-mode"]))


def ex_fields(args):
    f = parse_flags(args)
    return dict(seq_id_thr=float(f["--min-seq-id"]), max_seq_len=int(f["--max-seq-len"]),
                keep_target=int(f["--keep-target"]), rescore_mode=int(f["--rescore-mode"]))


def oracle_km(args, nucl):
    p = ob.KmParams(**km_fields(args, nucl))
    p.hash_start, p.hash_end = 0, U64MAX
    return p


def oracle_rs(args):
    return ob.RsParams(**rs_fields(args))


def oracle_ex(args):
    return ob.ExParams(**ex_fields(args))
