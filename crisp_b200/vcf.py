"""Batch VCF writer: host mirror of host/vcf_writer.c.

`format_records` turns the candidate sites of a batch (what the front end or `Engine.pileup_run` knows: positions,
allele counts, read depths, mapping-quality classes, the reference around them) and the engine's `Results` into the VCF
records `print_pooledvariant` (crisp/newcrispprint.c:347-476) would print — the same bytes — formatted on several host
threads into one buffer.  `vcf_header` is the reference's header for the same run (crisp_header_options.c), minus its
time stamp and command line, so that a whole file can be written without the reference.
"""
from __future__ import annotations

import ctypes as C
import os

import numpy as np

from .abi import CResults, Results, _RESULT_FIELDS

VCFLIB_PATH = os.path.join(os.path.dirname(os.path.abspath(__file__)), "csrc", "libcrispvcf.so")


class CVcfSites(C.Structure):
    _fields_ = [("n_sites", C.c_int32), ("n_pools", C.c_int32), ("chrom", C.c_char_p), ("pos", C.c_void_p), ("refbase", C.c_void_p),
                ("counts", C.c_void_p), ("indcounts", C.c_void_p), ("readdepths", C.c_void_p), ("mqcounts", C.c_void_p), ("refseq", C.c_void_p),
                ("ref_start", C.c_int32), ("ref_len", C.c_int32), ("chrom_len", C.c_int64)]


class CVcfOptions(C.Structure):
    _fields_ = [("ploidy", C.c_void_p), ("em_flag", C.c_int32), ("varpoolsize", C.c_int32), ("threads", C.c_int32), ("reserved", C.c_int32)]


class CExpandStats(C.Structure):
    _fields_ = [("n_header_lines", C.c_int64), ("n_records_in", C.c_int64), ("n_records_out", C.c_int64), ("n_multiallelic_indel", C.c_int64),
                ("n_over_poolsize", C.c_int64)]


_lib = None


def load_vcf_library() -> C.CDLL:
    global _lib
    if _lib is None:
        if not os.path.exists(VCFLIB_PATH):
            raise RuntimeError(f"{VCFLIB_PATH} is missing: build it with `python -m crisp_b200.build`")
        lib = C.CDLL(VCFLIB_PATH)
        lib.crisp_vcf_format.argtypes = [C.POINTER(CVcfSites), C.POINTER(CResults), C.POINTER(CVcfOptions), C.POINTER(C.c_void_p),
                                         C.POINTER(C.c_size_t), C.POINTER(C.c_int32)]
        lib.crisp_vcf_format.restype = C.c_int
        lib.crisp_vcf_free.argtypes = [C.c_void_p]
        lib.crisp_vcf_free.restype = None
        lib.crisp_vcf_expand_genotypes.argtypes = [C.c_char_p, C.c_size_t, C.c_int32, C.c_void_p, C.c_int32, C.POINTER(C.c_void_p), C.POINTER(C.c_size_t),
                                                   C.POINTER(CExpandStats)]
        lib.crisp_vcf_expand_genotypes.restype = C.c_int
        lib.crisp_bgzf_compress.argtypes = [C.c_char_p, C.c_size_t, C.c_int, C.c_int, C.POINTER(C.c_void_p), C.POINTER(C.c_size_t)]
        lib.crisp_bgzf_compress.restype = C.c_int
        lib.crisp_vcf_post_free.argtypes = [C.c_void_p]
        lib.crisp_vcf_post_free.restype = None
        _lib = lib
    return _lib


def results_as_c(res: Results):
    """(CResults viewing the numpy arrays of `res`, keep-alive list)."""
    c = CResults()
    c.n_sites, c.n_calls = res.n_sites, res.n_calls
    keep = []
    for name, dt, _ in _RESULT_FIELDS:
        a = np.ascontiguousarray(getattr(res, name), dtype=dt)
        keep.append(a)
        setattr(c, name, a.ctypes.data)
    ac, qv = np.ascontiguousarray(res.ac, np.int32), np.ascontiguousarray(res.qv, np.float64)
    keep += [ac, qv]
    c.ac, c.qv = ac.ctypes.data, qv.ctypes.data
    return c, keep


def format_records(sites, res: Results, ploidy, chrom: str, refseq: np.ndarray, ref_start: int = 0, chrom_len: int | None = None, em_flag: int = 1,
                   varpoolsize: int = 0, threads: int | None = None) -> bytes:
    """VCF records (bytes, one line per called site, coordinate order) for `sites` (a crisp_b200.pileup.PileupSites or any
    object with pos, refbase, counts, indcounts, readdepths, mqcounts) and their `res`."""
    lib = load_vcf_library()
    ploidy = np.ascontiguousarray(ploidy, np.int32)
    P = ploidy.shape[0]
    arr = {k: np.ascontiguousarray(getattr(sites, k), dt) for k, dt in (("pos", np.int32), ("refbase", np.uint8), ("counts", np.int32),
                                                                           ("indcounts", np.int32), ("readdepths", np.int32), ("mqcounts", np.int32))}
    ref = np.ascontiguousarray(refseq, np.uint8)
    cs = CVcfSites(res.n_sites, P, chrom.encode(), *[arr[k].ctypes.data for k in ("pos", "refbase", "counts", "indcounts", "readdepths", "mqcounts")],
                   ref.ctypes.data, ref_start, ref.shape[0], ref_start + ref.shape[0] if chrom_len is None else chrom_len)
    co = CVcfOptions(ploidy.ctypes.data, em_flag, varpoolsize, threads or min(16, os.cpu_count() or 1), 0)
    cres, keep = results_as_c(res)
    out, n, nrec = C.c_void_p(), C.c_size_t(0), C.c_int32(0)
    rc = lib.crisp_vcf_format(C.byref(cs), C.byref(cres), C.byref(co), C.byref(out), C.byref(n), C.byref(nrec))
    if rc != 0:
        raise RuntimeError(f"crisp_vcf_format failed: status {rc}" + (" (a called insertion/deletion allele is outside the batch writer's scope)" if rc == -5 else ""))
    try:
        return C.string_at(out, n.value)
    finally:
        lib.crisp_vcf_free(out)
        del keep


def expand_genotypes(vcf: bytes, poolsize: int | None = None, pool_sizes=None) -> tuple[bytes, dict]:
    """A pooled CRISP VCF with the per-pool allele counts of its first FORMAT field written out as genotypes of pool-size
    alleles ("0/0/0/1", key GT) — what scripts/convert_pooled_vcf.py does so that standard tools read the file; host/vcf_post.c.
    One pool size for all (`poolsize`: the counts are those of the ALT alleles) or one per pool (`pool_sizes`: the counts
    include the reference allele first).  Returns (text, counters of what was read, written and left out)."""
    lib = load_vcf_library()
    ps = None if pool_sizes is None else np.ascontiguousarray(pool_sizes, np.int32)
    out, n, st = C.c_void_p(), C.c_size_t(0), CExpandStats()
    rc = lib.crisp_vcf_expand_genotypes(vcf, len(vcf), int(poolsize) if poolsize else -1, None if ps is None else ps.ctypes.data, 0 if ps is None else ps.shape[0],
                                        C.byref(out), C.byref(n), C.byref(st))
    if rc != 0:
        raise ValueError(f"crisp_vcf_expand_genotypes failed: status {rc}" + (" (no pool size given, a malformed record, or a count that is not a number)" if rc == -1 else ""))
    try:
        return C.string_at(out, n.value), {k: getattr(st, k) for k, _ in CExpandStats._fields_}
    finally:
        lib.crisp_vcf_post_free(out)


def bgzf_compress(data: bytes, level: int = -1, threads: int | None = None) -> bytes:
    """`data` as BGZF (what bgzip writes: tabix/bcftools read it, and so does gzip), blocks deflated on host threads."""
    lib = load_vcf_library()
    out, n = C.c_void_p(), C.c_size_t(0)
    rc = lib.crisp_bgzf_compress(data, len(data), level, threads or min(16, os.cpu_count() or 1), C.byref(out), C.byref(n))
    if rc != 0:
        raise RuntimeError(f"crisp_bgzf_compress failed: status {rc}")
    try:
        return C.string_at(out, n.value)
    finally:
        lib.crisp_vcf_post_free(out)
