"""BAM files in, VCF file out — the whole call path without a line of the reference in the loop.

    decode_bams (host threads: BGZF inflate + BAM record split, host/bam_decode.c)
      -> Engine.pileup_run per window (device: allele counting, candidate screen, site packing, then the statistics engine)
      -> format_records (host threads: the reference's record text, host/vcf_writer.c)

It covers what `CRISP.binary --bams list --ref ref.fa --VCF out.vcf --poolsize n` does for substitution variants on
gap-free alignments (readmultiplebams.c:main -> callvariants -> jointvariantcaller -> newCRISPcaller -> print_pooledvariant);
alignments with gaps and overlapping mate pairs are refused by the decoder / the device pileup, never mis-called.  The
records are the reference's byte for byte (tests/test_pileup.py); the header states the same run parameters in this
program's own words.
"""
from __future__ import annotations

import os
import time

import numpy as np

from .abi import EngineConfig
from .engine import Engine
from .pileup import decode_bams, make_window, read_fasta
from .vcf import bgzf_compress, expand_genotypes, format_records

_INFO = [("NP", "1", "Integer", "pools with data"), ("DP", ".", "Integer", "filtered reads on the forward strand, reverse strand and from both mates, all pools"),
         ("CT", ".", "Float", "log10 contingency-table p-value per ALT allele"), ("AC", ".", "Integer", "estimated allele count over all pools per ALT allele"),
         ("AF", ".", "Float", "allele frequency estimated jointly over all pools (EM) per ALT allele"), ("FLANKSEQ", "1", "String", "reference sequence around the variant"),
         ("VT", "1", "String", "variant type"), ("EMstats", ".", "String", "likelihood-ratio statistics of the EM caller"),
         ("HWEstats", ".", "String", "Hardy-Weinberg test across pools"), ("VF", ".", "String", "EMpass* or EMfail per ALT allele"),
         ("VP", ".", "Integer", "pools carrying the ALT allele"), ("HP", ".", "Integer", "homopolymer / repeat length at an indel"),
         ("MQS", ".", "Integer", "reads with mapping quality 0-9, 10-19, 20-39, 40+")]
_FILTER = [("StrandBias", "alleles distributed unevenly over the strands"), ("LowDepth", "fewer than two filtered reads per haplotype on average"),
           ("LowMQ20", "more than 20 percent of the reads map with quality below 20"), ("LowMQ10", "more than 10 percent of the reads map with quality below 20")]
_FORMAT = [("GT", "1", "String", "genotype"), ("DP", "1", "Integer", "reads of the pool at the site"), ("GQ", "1", "Integer", "genotype quality"),
           ("AC", ".", "Integer", "maximum-likelihood allele count of the pool per ALT allele"), ("AF", ".", "Float", "allele frequency in the pool per ALT allele"),
           ("ADf", ".", "Integer", "forward-strand reads for REF and each ALT allele"), ("ADr", ".", "Integer", "reverse-strand reads for REF and each ALT allele"),
           ("ADb", ".", "Integer", "reads seen from both mates for REF and each ALT allele")]


def sample_id(path: str) -> str:
    """Pool name as the reference derives it from the BAM path (crisp_header_options.c:7-16): base name up to its first '.'."""
    base = path[path.rfind("/") + 1:]
    dot = base.find(".")
    return base[:dot] if dot > 0 else (base if dot < 0 else path)


def vcf_header(bams, fasta: str, poolsize: int, options: dict) -> bytes:
    lines = ["##fileformat=VCFv4.2", "##fileDate=" + time.strftime("%Y-%m-%d"), "##source=crisp_b200 (CRISP statistics on B200)", f"##reference={fasta}",
             f"##options: poolsize={poolsize},bamfiles={len(bams)},qvoffset={options.get('qv_offset', 33)},min_base_quality={options.get('min_q', 13)}"]
    lines += [f'##INFO=<ID={i},Number={n},Type={t},Description="{d}">' for i, n, t, d in _INFO]
    lines += [f'##FILTER=<ID={i},Description="{d}">' for i, d in _FILTER]
    lines += [f'##FORMAT=<ID={i},Number={n},Type={t},Description="{d}">' for i, n, t, d in _FORMAT]
    lines.append("#CHROM\tPOS\tID\tREF\tALT\tQUAL\tFILTER\tINFO\tFORMAT\t" + "\t".join(sample_id(b) for b in bams))
    return ("\n".join(lines) + "\n").encode()


_ARENA = {}   # the last call's pinned alignment arrays, taken over by the next call (pinning costs more than decoding)


def call_bams(bams, fasta: str, poolsize: int, out_vcf: str | None = None, options: dict | None = None, window: int = 1 << 20, threads: int | None = None,
              device: int = 0, min_mapq: int = 20, tid: int = 0, expand_gt: bool = False, region: tuple[int, int] | None = None) -> dict:
    """Call the first reference sequence of `fasta` from one BAM per pool.  Returns timings, counts and (when out_vcf is None)
    the VCF body.  `out_vcf` ending in .gz is written as BGZF (what bgzip writes); `expand_gt` writes pooled genotypes as
    pool-size allele lists ("0/0/0/1") the way scripts/convert_pooled_vcf.py converts them (host/vcf_post.c).  `region` =
    (start, end), 0-based half-open: only sites in it are called — the reference's --regions; the decoder reaches it through
    the BAI index when the files have one, so shards of a chromosome can run as separate processes (one per GPU)."""
    options = dict(options or {})
    threads = threads or os.cpu_count() or 1
    t0 = time.perf_counter()
    aln = decode_bams(bams, tid=tid, threads=threads, reuse=_ARENA.pop("aln", None), **({} if region is None else dict(start=int(region[0]), end=int(region[1]))))
    t1 = time.perf_counter()
    ref = read_fasta(fasta)
    L = ref.shape[0]
    P = len(bams)
    cfg = EngineConfig(ploidy=np.full(P, poolsize), options=dict(options, minimal_results=1))
    t2 = time.perf_counter()
    eng = Engine(cfg, device=device)
    t3 = time.perf_counter()
    body, n_sites, n_records, t_dev, t_fmt = [], 0, 0, 0.0, 0.0
    chrom = aln.stats["first_ref_name"]
    out = {}
    try:
        # the region's alignments and reference go to the device once; the windows are called on them where they lie
        window = max(1, min(window, max(65536, (2 << 30) // (32 * P))))   # the window's count table: 32 B per position and pool
        ta = time.perf_counter()
        r0, r1 = (0, L) if region is None else (max(0, int(region[0])), min(L, int(region[1])))
        eng.pileup_upload(aln, make_window(tid, r0, min(r1, r0 + window), ref, min_mapq=min_mapq))
        t_dev += time.perf_counter() - ta
        for start in range(r0, r1, window):
            end = min(r1, start + window)
            ta = time.perf_counter()
            if start != r0:
                eng.pileup_set_window(make_window(tid, start, end, None, min_mapq=min_mapq))
            sites, res = eng.pileup_run_resident(fetch=True)
            tb = time.perf_counter()
            text = format_records(sites, res, cfg.ploidy, chrom, ref, chrom_len=L, em_flag=cfg.options["em_flag"], varpoolsize=cfg.options["varpoolsize"],
                                  threads=min(threads, 16))
            tc = time.perf_counter()
            body.append(text)
            n_sites += sites.n_sites
            n_records += text.count(b"\n")
            t_dev += tb - ta
            t_fmt += tc - tb
        out = dict(decode_s=t1 - t0, fasta_s=t2 - t1, engine_create_s=t3 - t2, device_s=t_dev, format_s=t_fmt, candidate_sites=n_sites, records=n_records,
                   alignments=aln.n_reads, positions=L, decode_stats=aln.stats)
        if out_vcf:
            text = vcf_header(bams, fasta, poolsize, cfg.options) + b"".join(body)
            if expand_gt and poolsize != 2:   # genotypes of pool-size alleles instead of allele counts (scripts/convert_pooled_vcf.py);
                text, out["expand_stats"] = expand_genotypes(text, poolsize=poolsize)   # diploid output carries GT already
            with open(out_vcf, "wb") as f:
                f.write(bgzf_compress(text, threads=threads) if out_vcf.endswith(".gz") else text)   # .gz: BGZF, as bgzip writes it
        else:
            out["vcf_body"] = b"".join(body)
        out["wall_s"] = time.perf_counter() - t0   # BAM files read to VCF file written
    finally:
        t4 = time.perf_counter()
        eng.close()
        _ARENA["aln"] = aln
        out["teardown_s"] = time.perf_counter() - t4   # engine destruction (device buffers, pinned arenas): once per run
    return out
