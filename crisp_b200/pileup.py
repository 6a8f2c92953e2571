"""Device pileup: host mirror of `crispgpu_pileup_run` and of the BAM decoder (host/bam_decode.c).

`decode_bams` inflates and splits the pools' BAM files on host threads into the alignment arrays the pileup kernels read
(SURVEY.md §8f item 3); `Engine.pileup_run` (crisp_b200.engine) hands them to the device, which piles them up, screens the
candidate positions, packs the candidate sites and runs the statistics engine on them (§8f items 1 and 2) — the work of
calculate_allelecounts / advance_read (parsebam/variant.c:225-342, 56-131), sort_allelecounts and the potential-variant
test (variantcalls.c:62-92).  There is no CPU path: both libraries must be built (python -m crisp_b200.build).
"""
from __future__ import annotations

import ctypes as C
import os
from dataclasses import dataclass

import numpy as np

from .abi import ALTREAD_DTYPE, Batch

BAMLIB_PATH = os.path.join(os.path.dirname(os.path.abspath(__file__)), "csrc", "libcrispbam.so")

ALNREC_DTYPE = np.dtype([("pos", "<i4"), ("base_off", "<u4"), ("clip", "<u2"), ("mlen", "<u2"), ("mapq", "u1"), ("strand", "u1"),
                         ("reserved", "<u2")])
assert ALNREC_DTYPE.itemsize == 16

BAM_FILTER_MASK = 0x4 | 0x100 | 0x200 | 0x400  # unmapped | secondary | qcfail | duplicate (readmultiplebams.c:68)


class CAlignments(C.Structure):
    _fields_ = [("n_pools", C.c_int32), ("reserved", C.c_int32), ("pool_off", C.c_void_p), ("rec", C.c_void_p), ("mate_pos", C.c_void_p),
                ("isize", C.c_void_p), ("seq4", C.c_void_p), ("qual", C.c_void_p), ("n_bases", C.c_uint64)]


class CWindow(C.Structure):
    _fields_ = [("tid", C.c_int32), ("start", C.c_int32), ("end", C.c_int32), ("ref_start", C.c_int32), ("ref_len", C.c_int32),
                ("want_streams", C.c_int32), ("min_mapq", C.c_int32), ("reserved", C.c_int32), ("refseq", C.c_void_p)]


class CPileupSites(C.Structure):
    _fields_ = [("n_sites", C.c_int32), ("n_positions", C.c_int32)] + [(k, C.c_void_p) for k in (
        "pos", "refbase", "maxvec_al", "maxvec_ct", "counts", "indcounts", "readdepths", "mqcounts", "filtered", "read_off", "reads",
        "alt_off", "alt_reads")]


class CBamOptions(C.Structure):
    _fields_ = [("tid", C.c_int32), ("start", C.c_int32), ("end", C.c_int32), ("filter_mask", C.c_uint32), ("threads", C.c_int32),
                ("paired", C.c_int32), ("reuse", C.c_int32), ("use_index", C.c_int32), ("reserved", C.c_int32), ("alloc", C.c_void_p), ("release", C.c_void_p)]


class CBamReads(C.Structure):
    _fields_ = [("aln", CAlignments), ("n_records", C.c_int64), ("n_flag_filtered", C.c_int64), ("n_outside", C.c_int64),
                ("n_complex", C.c_int64), ("compressed_bytes", C.c_int64), ("inflated_bytes", C.c_int64), ("read_s", C.c_double),
                ("inflate_s", C.c_double), ("parse_s", C.c_double), ("n_refs", C.c_int32), ("first_ref_len", C.c_int32),
                ("first_ref_name", C.c_char * 64), ("cap_reads", C.c_uint64), ("cap_bases", C.c_uint64), ("cap_pools", C.c_int32), ("cap_paired", C.c_int32)]


_bamlib = None


def load_bam_library() -> C.CDLL:
    global _bamlib
    if _bamlib is None:
        if not os.path.exists(BAMLIB_PATH):
            raise RuntimeError(f"{BAMLIB_PATH} is missing: build it with `python -m crisp_b200.build`")
        lib = C.CDLL(BAMLIB_PATH)
        lib.crisp_bam_decode.argtypes = [C.POINTER(C.c_char_p), C.c_int32, C.POINTER(CBamOptions), C.POINTER(CBamReads)]
        lib.crisp_bam_decode.restype = C.c_int
        lib.crisp_bam_release.argtypes = [C.POINTER(CBamReads), C.c_void_p]
        lib.crisp_bam_release.restype = None
        _bamlib = lib
    return _bamlib


@dataclass
class Alignments:
    """Struct-of-arrays alignment set (`crispgpu_alignments`): numpy views, pool-major, BAM order inside a pool."""
    n_pools: int
    pool_off: np.ndarray          # uint32 [P+1]
    rec: np.ndarray               # ALNREC_DTYPE [R]
    seq4: np.ndarray              # uint8 [n_bases/2]
    qual: np.ndarray              # uint8 [n_bases]
    mate_pos: np.ndarray | None = None
    isize: np.ndarray | None = None
    stats: dict | None = None
    _owner: object = None

    @property
    def n_reads(self) -> int:
        return int(self.pool_off[-1])

    @property
    def n_bases(self) -> int:
        return int(self.qual.shape[0])

    def nbytes(self) -> int:
        return sum(a.nbytes for a in (self.pool_off, self.rec, self.seq4, self.qual, self.mate_pos, self.isize) if a is not None)

    def as_c(self) -> CAlignments:
        ptr = lambda a: a.ctypes.data if a is not None and a.size else None  # noqa: E731
        return CAlignments(self.n_pools, 0, self.pool_off.ctypes.data, ptr(self.rec), ptr(self.mate_pos), ptr(self.isize), ptr(self.seq4),
                           ptr(self.qual), self.n_bases)


class _DecodedOwner:
    def __init__(self, lib, reads, release):
        self.lib, self.reads, self.release = lib, reads, release

    def __del__(self):
        try:
            self.lib.crisp_bam_release(C.byref(self.reads), self.release)
        except Exception:
            pass


def decode_bams(paths, tid: int = 0, start: int = 0, end: int = 2**31 - 1, threads: int | None = None, paired: bool = False,
                pinned: bool = True, filter_mask: int = BAM_FILTER_MASK, allow_complex: bool = False, reuse: Alignments | None = None,
                use_index: bool = True) -> Alignments:
    """Decode one coordinate-sorted BAM per pool into an `Alignments` set (alignments overlapping [start, end) of reference
    `tid`).  Raises when the files hold alignments outside the device pileup's scope (CIGARs with gaps) unless
    `allow_complex` (they are then dropped and counted in stats['n_complex']).  `reuse`: an earlier result of this function
    (same `pinned`) whose arrays are taken over when large enough — pinning memory costs more than decoding into it; that
    earlier object must not be used afterwards.  `use_index`: a region is reached through the BAI index next to each file
    (x.bam.bai or x.bai) when there is one — only the blocks it names are inflated; the result is the same without."""
    lib = load_bam_library()
    alloc = release = None
    if pinned:
        from .engine import load_library
        g = load_library()
        alloc = C.cast(g.crispgpu_host_alloc, C.c_void_p).value
        release = C.cast(g.crispgpu_host_free, C.c_void_p).value
    out = CBamReads()
    taking = reuse is not None and isinstance(reuse._owner, _DecodedOwner) and reuse._owner.release == release
    if taking:
        C.memmove(C.byref(out), C.byref(reuse._owner.reads), C.sizeof(CBamReads))
        reuse._owner.reads = CBamReads()   # ownership moves to this call
        reuse.rec = reuse.seq4 = reuse.qual = reuse.pool_off = None
    opt = CBamOptions(tid, start, end, filter_mask, threads or (os.cpu_count() or 1), 1 if paired else 0, 1 if taking else 0, 1 if use_index else 0, 0, alloc, release)
    arr = (C.c_char_p * len(paths))(*[os.fsencode(p) for p in paths])
    rc = lib.crisp_bam_decode(arr, len(paths), C.byref(opt), C.byref(out))
    if rc != 0:
        raise RuntimeError(f"crisp_bam_decode failed: status {rc}")
    owner = _DecodedOwner(lib, out, release)
    P = len(paths)
    view = lambda ptr, dt, n: np.ctypeslib.as_array(C.cast(ptr, C.POINTER(C.c_uint8)), shape=(n * np.dtype(dt).itemsize,)).view(dt) if n else np.zeros(0, dt)  # noqa: E731
    pool_off = view(out.aln.pool_off, np.uint32, P + 1)
    R, B = int(pool_off[-1]), int(out.aln.n_bases)
    stats = {k: getattr(out, k) for k in ("n_records", "n_flag_filtered", "n_outside", "n_complex", "compressed_bytes", "inflated_bytes",
                                          "read_s", "inflate_s", "parse_s", "n_refs", "first_ref_len")}
    stats["first_ref_name"] = out.first_ref_name.decode()
    if out.n_complex and not allow_complex:
        raise RuntimeError(f"{out.n_complex} alignments have CIGARs outside the device pileup's scope (gaps); route this data through the host pileup")
    return Alignments(P, pool_off, view(out.aln.rec, ALNREC_DTYPE, R), view(out.aln.seq4, np.uint8, B // 2), view(out.aln.qual, np.uint8, B),
                      view(out.aln.mate_pos, np.int32, R) if paired else None, view(out.aln.isize, np.int32, R) if paired else None, stats, owner)


def pack_alignments(pools, pinned: bool = False) -> Alignments:
    """Build an `Alignments` set from Python data (tests): pools = list over pools of lists of dicts
    {pos, seq (str or codes), qual (ints), clip=0, mapq=60, strand=0}, each pool sorted by pos."""
    nt16 = {c: i for i, c in enumerate("=ACMGRSVTWYHKDBN")}
    recs, seqs, quals, off = [], [], [], [0]
    base = 0
    for reads in pools:
        for r in reads:
            s = r["seq"]
            codes = np.array([nt16[c] for c in s], np.uint8) if isinstance(s, str) else np.asarray(s, np.uint8)
            L = codes.shape[0]
            pad = (L + 7) & ~7
            clip = int(r.get("clip", 0))
            mlen = int(r.get("mlen", L - clip))
            recs.append((int(r["pos"]), base, clip, mlen, int(r.get("mapq", 60)), int(r.get("strand", 0)), 0))
            c8 = np.zeros(pad, np.uint8)
            c8[:L] = codes
            seqs.append((c8[0::2] << 4) | c8[1::2])
            q8 = np.zeros(pad, np.uint8)
            q8[:L] = np.asarray(r["qual"], np.uint8)
            quals.append(q8)
            base += pad
        off.append(len(recs))
    rec = np.array(recs, dtype=ALNREC_DTYPE) if recs else np.zeros(0, ALNREC_DTYPE)
    seq4 = np.concatenate(seqs) if seqs else np.zeros(0, np.uint8)
    qual = np.concatenate(quals) if quals else np.zeros(0, np.uint8)
    return Alignments(len(pools), np.array(off, np.uint32), rec, seq4, qual)


@dataclass
class PileupSites:
    """Owning numpy copy of `crispgpu_pileup_sites` (+ the packed streams when they were requested)."""
    n_sites: int
    n_positions: int
    pos: np.ndarray
    refbase: np.ndarray
    maxvec_al: np.ndarray
    maxvec_ct: np.ndarray
    counts: np.ndarray
    indcounts: np.ndarray
    readdepths: np.ndarray
    mqcounts: np.ndarray
    filtered: np.ndarray
    read_off: np.ndarray | None = None
    reads: np.ndarray | None = None
    alt_off: np.ndarray | None = None
    alt_reads: np.ndarray | None = None

    def as_batch(self, n_pools: int, tid: int = 0) -> Batch:
        """The candidate sites as a host `Batch` (needs the streams): what the reference's front end would have packed."""
        return Batch(n_pools, tid=np.full(self.n_sites, tid, np.int32), pos=self.pos, refbase=self.refbase, maxvec_al=self.maxvec_al,
                     maxvec_ct=self.maxvec_ct, counts=self.counts, indcounts=self.indcounts, read_off=self.read_off, reads=self.reads,
                     alt_off=self.alt_off, alt_reads=self.alt_reads)


def sites_from_c(cs: CPileupSites, P: int) -> PileupSites:
    n = cs.n_sites

    def take(ptr, dt, shape):
        size = int(np.prod(shape))
        if not ptr or size == 0:
            return np.zeros(shape, dt)
        return np.ctypeslib.as_array(C.cast(ptr, C.POINTER(C.c_uint8)), shape=(size * np.dtype(dt).itemsize,)).view(dt).reshape(shape).copy()

    out = PileupSites(n, cs.n_positions, take(cs.pos, np.int32, (n,)), take(cs.refbase, np.uint8, (n,)), take(cs.maxvec_al, np.uint8, (n, 8)),
                      take(cs.maxvec_ct, np.int32, (n, 8)), take(cs.counts, np.int32, (n, 24)), take(cs.indcounts, np.int32, (n, P, 24)),
                      take(cs.readdepths, np.int32, (n, P)), take(cs.mqcounts, np.int32, (n, 4)), take(cs.filtered, np.int32, (n, 2)))
    if cs.read_off:
        out.read_off = take(cs.read_off, np.uint32, (n * P + 1,))
        out.reads = take(cs.reads, np.uint16, (int(out.read_off[-1]),))
        out.alt_off = take(cs.alt_off, np.uint32, (n * 5 + 1,))
        out.alt_reads = take(cs.alt_reads, ALTREAD_DTYPE, (int(out.alt_off[-1]),))
    elif n == 0:
        out.read_off, out.reads = np.zeros(1, np.uint32), np.zeros(0, np.uint16)
        out.alt_off, out.alt_reads = np.zeros(1, np.uint32), np.zeros(0, ALTREAD_DTYPE)
    return out


def make_window(tid: int, start: int, end: int, refseq, ref_start: int = 0, want_streams: bool = False, min_mapq: int = 20):
    """(CWindow, keep-alive) for reference bytes `refseq` (uint8 array of characters) starting at position ref_start; refseq None:
    a window over the reference stretch that is already on the device (Engine.pileup_set_window)."""
    if refseq is None:
        return CWindow(tid, start, end, ref_start, 0, 1 if want_streams else 0, min_mapq, 0, None), None
    ref = np.ascontiguousarray(refseq, dtype=np.uint8)
    return CWindow(tid, start, end, ref_start, ref.shape[0], 1 if want_streams else 0, min_mapq, 0, ref.ctypes.data), ref


def read_fasta(path: str) -> np.ndarray:
    """The first sequence of a FASTA file as a uint8 array of characters."""
    with open(path, "rb") as f:
        lines = f.read().split(b"\n")
    body = []
    for ln in lines[1:]:
        if ln.startswith(b">"):
            break
        body.append(ln.strip())
    return np.frombuffer(b"".join(body), dtype=np.uint8).copy()
