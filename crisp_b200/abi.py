"""ctypes mirror of include/crispgpu.h — the C ABI of the B200 statistics engine.

Every structure here is field-for-field the one declared in include/crispgpu.h (which cites the
reference interface each field replaces: `struct VARIANT`, parsebam/variant.h:50-95, and the option
globals of optionparser.c:17-136).  `Batch` is the host-side struct-of-arrays container a front end
fills where the reference calls newCRISPcaller (variantcalls.c:115).
"""
from __future__ import annotations

import ctypes as C
from dataclasses import dataclass, field
from typing import Optional

import numpy as np

ABI_VERSION = 3
MAXALLELES = 8
NCOUNT = 24
MAXSLOTS = 5

SLOT_UNTESTED, SLOT_REJECT_AF, SLOT_REJECT_CT, SLOT_NOCALL, SLOT_CALLED = range(5)

OK = 0
ERR_ARG, ERR_NO_DEVICE, ERR_CUDA, ERR_NOMEM, ERR_UNSUPPORTED, ERR_TICKET = -1, -2, -3, -4, -5, -6

ALTREAD_DTYPE = np.dtype(
    [("pool", "<u2"), ("offset", "<u2"), ("aligned", "<u2"), ("strand", "u1"), ("pad", "u1")]
)


class Config(C.Structure):
    _fields_ = [
        ("abi_version", C.c_int32),
        ("n_pools", C.c_int32),
        ("ploidy", C.POINTER(C.c_int32)),
        ("varpoolsize", C.c_int32),
        ("em_flag", C.c_int32),
        ("max_iter", C.c_int32),
        ("chisq_permutation", C.c_int32),
        ("fast_filter", C.c_int32),
        ("use_base_qvs", C.c_int32),
        ("min_reads", C.c_int32),
        ("qv_offset", C.c_int32),
        ("min_q", C.c_int32),
        ("pivot_sample", C.c_int32),
        ("association", C.c_int32),
        ("phenotypes", C.POINTER(C.c_int32)),
        ("thresh1", C.c_double),
        ("min_qpv", C.c_double),
        ("thresh3", C.c_double),
        ("ser", C.c_double),
        ("relax", C.c_double),
        ("qmin", C.c_double),
        ("agilent_ref_bias", C.c_double),
        ("theta", C.c_double),
        ("rng_seed", C.c_uint64),
        ("exact_mode", C.c_int32),
        ("use_qv", C.c_int32),
        ("blocking_wait", C.c_int32),
        ("minimal_results", C.c_int32),
    ]


class CBatch(C.Structure):
    _fields_ = [
        ("n_sites", C.c_int32),
        ("reserved", C.c_int32),
        ("tid", C.c_void_p),
        ("pos", C.c_void_p),
        ("refbase", C.c_void_p),
        ("maxvec_al", C.c_void_p),
        ("maxvec_ct", C.c_void_p),
        ("counts", C.c_void_p),
        ("indcounts", C.c_void_p),
        ("indel_len", C.c_void_p),
        ("hplen", C.c_void_p),
        ("read_off", C.c_void_p),
        ("reads", C.c_void_p),
        ("alt_off", C.c_void_p),
        ("alt_reads", C.c_void_p),
        ("amb_idx", C.c_void_p),
        ("amb_dec", C.c_void_p),
        ("amb_ep", C.c_void_p),
        ("n_amb", C.c_int32),
        ("reserved2", C.c_int32),
        ("qhigh", C.c_void_p),
        ("stats", C.c_void_p),
        ("tstats", C.c_void_p),
        ("bin_off", C.c_void_p),
        ("bins", C.c_void_p),
        ("ic_off", C.c_void_p),
        ("ic_sparse", C.c_void_p),
        ("arena_base", C.c_void_p),
        ("arena_bytes", C.c_uint64),
        ("bins16", C.c_void_p),
        ("bin_len", C.c_void_p),
        ("bin_site", C.c_void_p),
        ("ic16", C.c_void_p),
        ("ic_len", C.c_void_p),
        ("ic_site", C.c_void_p),
        ("alt32", C.c_void_p),
    ]


_RESULT_FIELDS = [
    # name, numpy dtype, per-(site) width (5 = per slot, 1 = per site)
    ("varalleles", np.int32, 1),
    ("strandpass", np.int32, 1),
    ("status", np.uint8, 5),
    ("allele", np.uint8, 5),
    ("call_rank", np.int8, 5),
    ("call_row", np.int32, 5),
    ("paf", np.float64, 5),
    ("ctpval", np.float64, 5),
    ("chi2pval", np.float64, 5),
    ("af_ml", np.float64, 5),
    ("delta", np.float64, 5),
    ("deltaf", np.float64, 5),
    ("hwe", np.float64, 5),
    ("qvpf", np.float64, 5),
    ("qvpr", np.float64, 5),
    ("varpools", np.int32, 5),
    ("allelecount", np.int32, 5),
    ("variantpools", np.int32, 5),
    ("em_iters", np.int32, 5),
    ("perm_k", np.int32, 5),
    ("perm_hits", np.int32, 5),
    ("assoc_p", np.float64, 5),
    ("assoc_llr", np.float64, 5),
    ("assoc_or", np.float64, 5),
    ("assoc_p1", np.float64, 5),
    ("assoc_llr1", np.float64, 5),
    ("assoc_or1", np.float64, 5),
]


class CResults(C.Structure):
    _fields_ = (
        [("n_sites", C.c_int32), ("n_calls", C.c_int32)]
        + [(name, C.c_void_p) for name, _, _ in _RESULT_FIELDS]
        + [("ac", C.c_void_p), ("qv", C.c_void_p)]
    )


class Timing(C.Structure):
    _fields_ = [
        ("h2d_ms", C.c_float),
        ("evaluate_ms", C.c_float),
        ("permute_ms", C.c_float),
        ("filter_ms", C.c_float),
        ("em_ms", C.c_float),
        ("legacy_ms", C.c_float),
        ("d2h_ms", C.c_float),
        ("total_ms", C.c_float),
        ("n_tasks_em", C.c_int32),
        ("n_tasks_perm", C.c_int32),
        ("launches", C.c_int64),
        ("h2d_bytes", C.c_int64),
        ("d2h_bytes", C.c_int64),
        ("fetched_bytes", C.c_int64),
        ("pileup_ms", C.c_float),
        ("pile_tile_ms", C.c_float),
        ("n_positions", C.c_int32),
        ("n_alignments", C.c_int32),
        ("pile_bytes", C.c_int64),
    ]


def default_config_dict() -> dict:
    """The reference's defaults (readmultiplebams.c:21-47, newcrispcaller.c:20-32, crispEM.c:14,26)."""
    return dict(
        varpoolsize=0, em_flag=1, max_iter=20000, chisq_permutation=0, fast_filter=1, use_base_qvs=1,
        min_reads=4, qv_offset=33, min_q=13, pivot_sample=0, association=0, thresh1=-3.5, min_qpv=-5.0,
        thresh3=-2.0, ser=-1.0, relax=0.75, qmin=23.0, agilent_ref_bias=0.5, theta=0.001, rng_seed=1,
        exact_mode=0, use_qv=0, blocking_wait=0, minimal_results=0,
    )


@dataclass
class EngineConfig:
    """Python-side configuration; `to_c()` snapshots it into the ABI struct."""

    ploidy: np.ndarray
    options: dict = field(default_factory=dict)
    phenotypes: Optional[np.ndarray] = None

    def __post_init__(self):
        self.ploidy = np.ascontiguousarray(self.ploidy, dtype=np.int32)
        opts = default_config_dict()
        unknown = set(self.options) - set(opts)
        if unknown:
            raise KeyError(f"unknown engine options: {sorted(unknown)}")
        opts.update(self.options)
        if "varpoolsize" not in self.options:
            # init_poolsizes (init_variant.c:143-151): 1 when adjacent pool sizes differ
            opts["varpoolsize"] = int(np.any(self.ploidy[1:] != self.ploidy[:-1]))
        self.options = opts
        if self.phenotypes is not None:
            self.phenotypes = np.ascontiguousarray(self.phenotypes, dtype=np.int32)

    @property
    def n_pools(self) -> int:
        return int(self.ploidy.shape[0])

    def to_c(self) -> Config:
        cfg = Config()
        cfg.abi_version = ABI_VERSION
        cfg.n_pools = self.n_pools
        cfg.ploidy = self.ploidy.ctypes.data_as(C.POINTER(C.c_int32))
        cfg.phenotypes = (
            self.phenotypes.ctypes.data_as(C.POINTER(C.c_int32)) if self.phenotypes is not None else None
        )
        for k, v in self.options.items():
            setattr(cfg, k, v)
        cfg._keep = (self.ploidy, self.phenotypes)
        return cfg


_BATCH_ARRAYS = [
    # name, dtype, required
    ("tid", np.int32, True),
    ("pos", np.int32, True),
    ("refbase", np.uint8, True),
    ("maxvec_al", np.uint8, True),   # required by the constructor; without_maxvec() drops them (sorted on the device)
    ("maxvec_ct", np.int32, True),
    ("counts", np.int32, True),
    ("indcounts", np.int32, True),
    ("indel_len", np.int16, False),
    ("hplen", np.int16, False),
    ("read_off", np.uint32, True),
    ("reads", np.uint16, True),
    ("alt_off", np.uint32, False),
    ("alt_reads", ALTREAD_DTYPE, False),
    ("amb_idx", np.int32, False),
    ("amb_dec", np.int32, False),
    ("amb_ep", np.float64, False),
    ("qhigh", np.float64, False),
    ("stats", np.float64, False),
    ("tstats", np.float64, False),
    ("bin_off", np.uint32, False),
    ("bins", np.uint32, False),
    ("ic_off", np.uint32, False),
    ("ic_sparse", np.uint32, False),
]


def sparse_counts_from_dense(indcounts: np.ndarray):
    """(ic_off, ic_sparse): the non-zero entries of every pool's 24-entry indcounts row as `crispgpu_count` words."""
    rows = indcounts.reshape(-1, 24)
    cell, idx = np.nonzero(rows)
    val = rows[cell, idx].astype(np.int64)
    if val.size and (val.min() < 0 or val.max() >= (1 << 27)):
        raise ValueError("allele counts outside the 27-bit range of the sparse format")
    ic = ((idx.astype(np.uint64) << 27) | val.astype(np.uint64)).astype(np.uint32)
    off = np.concatenate([[0], np.cumsum(np.bincount(cell, minlength=rows.shape[0]))]).astype(np.uint32)
    return off, ic


def narrow_arrays(batch: "Batch") -> dict:
    """numpy statement of the narrow transport form of a compact batch (crispgpu_batch.bins16 ... ic_site): what
    crispgpu_batch_pack_narrow writes.  Test aid; the engine's path goes through the C packer."""
    n, P = batch.n_sites, batch.n_pools

    def split(words, cap, cnt_of, make):
        cnt = cnt_of(words).astype(np.int64)
        pieces = (cnt + cap - 1) // cap
        idx = np.repeat(np.arange(words.shape[0]), pieces)
        k = np.arange(idx.shape[0]) - np.repeat(np.cumsum(pieces) - pieces, pieces)
        part = np.minimum(cnt[idx] - k * cap, cap)
        return make(words[idx], part).astype(np.uint16), pieces

    b16, bp = split(batch.bins, 32, lambda w: w >> 16,
                    lambda w, c: (w & 0x7F) | (((w >> 8) & 0xF) << 7) | ((c - 1).astype(np.uint32) << 11))
    i16, ip = split(batch.ic_sparse, 2047, lambda w: w & 0x7FFFFFF, lambda w, c: ((w >> 27) << 11) | c.astype(np.uint32))

    def lens(off, pieces):
        cum = np.concatenate([[0], np.cumsum(pieces)])
        per = cum[off[1:].astype(np.int64)] - cum[off[:-1].astype(np.int64)]
        site = np.concatenate([[0], np.cumsum(per.reshape(n, P).sum(axis=1))]).astype(np.uint32)
        return per.astype(np.uint8), site

    bl, bs = lens(batch.bin_off, bp)
    il, isite = lens(batch.ic_off, ip)
    return dict(bins16=b16, bin_len=bl, bin_site=bs, ic16=i16, ic_len=il, ic_site=isite)


def bins_from_reads(read_off: np.ndarray, reads: np.ndarray):
    """(bin_off, bins) of the compact read format (`crispgpu_bin`): per (site, pool) cell the distinct read values with
    their multiplicities, ascending by value; counts above 65535 are split over several bins."""
    ro = read_off.astype(np.int64)
    ncell = ro.size - 1
    cell = np.repeat(np.arange(ncell, dtype=np.int64), ro[1:] - ro[:-1])
    key, cnt = np.unique((cell << 16) | reads.astype(np.int64), return_counts=True)
    rep = (cnt + 65534) // 65535
    if (rep > 1).any():
        key = np.repeat(key, rep)
        first = np.cumsum(rep) - rep
        part = np.arange(key.size) - np.repeat(first, rep)
        cnt = np.minimum(np.repeat(cnt, rep) - part * 65535, 65535)
    bins = ((cnt.astype(np.uint64) << 16) | (key & 0xFFFF).astype(np.uint64)).astype(np.uint32)
    per = np.bincount(key >> 16, minlength=ncell)
    bin_off = np.concatenate([[0], np.cumsum(per)]).astype(np.uint32)
    return bin_off, bins


def reads_from_bins(bin_off: np.ndarray, bins: np.ndarray) -> np.ndarray:
    """The read multiset a binned batch stands for (cell by cell, bins in order)."""
    return np.repeat((bins & 0xFFFF).astype(np.uint16), (bins >> 16).astype(np.int64))


class Batch:
    """Struct-of-arrays batch of candidate sites (`crispgpu_batch`)."""

    def __init__(self, n_pools: int, **arrays):
        self.n_pools = int(n_pools)
        # the compact forms stand in for the dense arrays they replace
        optional = set()
        if arrays.get("bins") is not None or arrays.get("bin_off") is not None:
            optional |= {"reads", "read_off"}
        if arrays.get("ic_off") is not None:
            optional |= {"indcounts"}
        if arrays.get("maxvec_al") is None and arrays.get("maxvec_ct") is None:
            optional |= {"maxvec_al", "maxvec_ct"}   # sorted on the device (k_candidates)
        for name, dt, req in _BATCH_ARRAYS:
            a = arrays.pop(name, None)
            if a is None:
                if req and name not in optional:
                    raise ValueError(f"batch array {name!r} is required")
                setattr(self, name, None)
                continue
            setattr(self, name, np.ascontiguousarray(a, dtype=dt))
        if arrays:
            raise KeyError(f"unknown batch arrays: {sorted(arrays)}")
        self.n_sites = int(self.tid.shape[0])
        self.validate()

    def validate(self) -> None:
        n, P = self.n_sites, self.n_pools
        def want(name, size):
            a = getattr(self, name)
            if a is not None and a.size != size:
                raise ValueError(f"batch array {name!r}: {a.size} elements, expected {size}")
        want("pos", n); want("refbase", n); want("maxvec_al", n * 8); want("maxvec_ct", n * 8)
        want("counts", n * 24); want("indcounts", n * P * 24); want("indel_len", n * 8); want("hplen", n * 8)
        want("read_off", n * P + 1); want("alt_off", n * 5 + 1); want("amb_idx", n * 5)
        want("qhigh", n * P * 16); want("stats", n * P * 16); want("tstats", n * 16)
        if n and self.reads is not None and int(self.read_off[-1]) != self.reads.size:
            raise ValueError("read_off[-1] != len(reads)")
        if self.ic_off is not None:
            want("ic_off", n * P + 1)
            if n and int(self.ic_off[-1]) != (0 if self.ic_sparse is None else self.ic_sparse.size):
                raise ValueError("ic_off[-1] != len(ic_sparse)")
        elif self.indcounts is None:
            raise ValueError("batch needs indcounts or ic_off/ic_sparse")
        if self.bins is not None:
            want("bin_off", n * P + 1)
            if n and (int(self.bin_off[-1]) != self.bins.size or (self.read_off is not None and int((self.bins >> 16).sum(dtype=np.int64)) != int(self.read_off[-1]))):
                raise ValueError("bins do not match bin_off / read_off")
        if self.alt_off is not None and (self.alt_reads is None or int(self.alt_off[-1]) != self.alt_reads.size):
            raise ValueError("alt_off[-1] != len(alt_reads)")
        self.n_amb = 0
        if self.amb_idx is not None:
            self.n_amb = int(self.amb_idx.max(initial=-1)) + 1
            if self.n_amb:
                want("amb_dec", self.n_amb * P * 4); want("amb_ep", self.n_amb * P * 2)

    def nbytes(self) -> int:
        return sum(getattr(self, name).nbytes for name, _, _ in _BATCH_ARRAYS if getattr(self, name) is not None)

    def binned(self) -> "Batch":
        """The same batch in the compact read format: `bins`/`bin_off` instead of `reads` (what a front end ships when
        PCIe is the bound; the engine rebuilds the reads on the device)."""
        if self.bins is not None:
            return self
        out = Batch.__new__(Batch)
        out.__dict__.update(self.__dict__)
        out.__dict__.pop("_arena", None)   # the new arrays live outside a pinned arena the original may declare
        out.bin_off, out.bins = bins_from_reads(self.read_off, self.reads)
        out.reads = None
        out.read_off = None   # indexes the read stream only
        return out

    def compact(self) -> "Batch":
        """Transfer-optimised form: reads as (value, count) bins and the per-pool allele counts as their non-zero entries
        (both rebuilt in device memory by the engine); about a fifth of the bytes on binned-quality data."""
        out = self.binned()
        if out is self:
            out = Batch.__new__(Batch)
            out.__dict__.update(self.__dict__)
        if out.ic_off is None:
            out.__dict__.pop("_arena", None)
            out.ic_off, out.ic_sparse = sparse_counts_from_dense(self.indcounts)
            out.indcounts = None
        return out

    def without_maxvec(self) -> "Batch":
        """The batch without maxvec_al/maxvec_ct: the engine sorts the alleles itself (k_candidates)."""
        out = Batch.__new__(Batch)
        out.__dict__.update(self.__dict__)
        out.maxvec_al = out.maxvec_ct = None
        return out

    def as_c(self) -> CBatch:
        b = CBatch()
        b.n_sites = self.n_sites
        b.n_amb = self.n_amb
        for name, _, _ in _BATCH_ARRAYS:
            a = getattr(self, name)
            setattr(b, name, a.ctypes.data if a is not None else None)
        arena = getattr(self, "_arena", None)   # (address, bytes) of the one pinned allocation holding every array (pin_batch)
        if arena is not None:
            b.arena_base, b.arena_bytes = arena
        b._keep = self
        return b

    def select(self, idx) -> "Batch":
        """Sub-batch of the given site indices (used by the region sharder and the tests)."""
        idx = np.asarray(idx, dtype=np.int64)
        P = self.n_pools
        out = {}
        for name in ("tid", "pos", "refbase"):
            out[name] = getattr(self, name)[idx]
        for name, w in (("maxvec_al", 8), ("maxvec_ct", 8), ("counts", 24), ("indcounts", P * 24),
                        ("indel_len", 8), ("hplen", 8), ("qhigh", P * 16), ("stats", P * 16), ("tstats", 16)):
            a = getattr(self, name)
            if a is not None:
                out[name] = a.reshape(self.n_sites, w)[idx].reshape(-1)
        # ragged arrays indexed per (site, pool) cell: reads, or their compact bins; sparse per-pool counts
        def ragged(off, data, empty_dtype):
            o = off.astype(np.int64)
            seg = [data[a:b] for a, b in zip(o[idx * P], o[idx * P + P])] if data is not None else []
            per = (o[1:] - o[:-1]).reshape(self.n_sites, P)[idx].reshape(-1)
            return (np.concatenate(seg) if seg else np.zeros(0, empty_dtype)), np.concatenate([[0], np.cumsum(per)]).astype(np.uint32)

        if self.reads is not None:
            out["reads"], out["read_off"] = ragged(self.read_off, self.reads, np.uint16)
        if self.bin_off is not None:
            out["bins"], out["bin_off"] = ragged(self.bin_off, self.bins, np.uint32)
        if self.ic_off is not None:
            out["ic_sparse"], out["ic_off"] = ragged(self.ic_off, self.ic_sparse, np.uint32)
        if self.alt_off is not None:
            ao = self.alt_off.astype(np.int64)
            seg = [self.alt_reads[ao[s * 5]:ao[s * 5 + 5]] for s in idx]
            out["alt_reads"] = np.concatenate(seg) if seg else np.zeros(0, ALTREAD_DTYPE)
            per = (ao[1:] - ao[:-1]).reshape(self.n_sites, 5)[idx].reshape(-1)
            out["alt_off"] = np.concatenate([[0], np.cumsum(per)]).astype(np.uint32)
        if self.amb_idx is not None:
            ai = self.amb_idx.reshape(self.n_sites, 5)[idx].reshape(-1).copy()
            used = np.unique(ai[ai >= 0])
            remap = -np.ones(max(self.n_amb, 1), dtype=np.int32)
            remap[used] = np.arange(used.size, dtype=np.int32)
            ai[ai >= 0] = remap[ai[ai >= 0]]
            out["amb_idx"] = ai
            if used.size:
                out["amb_dec"] = self.amb_dec.reshape(self.n_amb, P * 4)[used].reshape(-1)
                if self.amb_ep is not None:
                    out["amb_ep"] = self.amb_ep.reshape(self.n_amb, P * 2)[used].reshape(-1)
        return Batch(P, **out)


def concat_batches(parts: list) -> Batch:
    """Concatenate batches of the same pool layout (site order preserved)."""
    if not parts:
        raise ValueError("no batches")
    P = parts[0].n_pools
    out = {}
    for name in ("tid", "pos", "refbase", "maxvec_al", "maxvec_ct", "counts", "indcounts", "reads"):
        out[name] = np.concatenate([getattr(b, name) for b in parts])
    for name, w, fill in (("indel_len", 8, 0), ("hplen", 8, 0), ("qhigh", P * 16, None), ("stats", P * 16, None), ("tstats", 16, None)):
        have = [getattr(b, name) is not None for b in parts]
        if not any(have):
            continue
        if not all(have) and fill is None:
            raise ValueError(f"cannot concatenate: {name} present in only some batches")
        dt = next(getattr(b, name) for b in parts if getattr(b, name) is not None).dtype
        out[name] = np.concatenate([getattr(b, name) if getattr(b, name) is not None else np.full(b.n_sites * w, fill, dt) for b in parts])
    def cat_off(name):
        offs, base = [np.zeros(1, np.int64)], 0
        for b in parts:
            o = getattr(b, name).astype(np.int64)
            offs.append(o[1:] + base)
            base += int(o[-1])
        return np.concatenate(offs).astype(np.uint32)
    out["read_off"] = cat_off("read_off")
    if any(b.alt_off is not None for b in parts):
        if not all(b.alt_off is not None for b in parts):
            raise ValueError("cannot concatenate: alt_off present in only some batches")
        out["alt_off"] = cat_off("alt_off")
        out["alt_reads"] = np.concatenate([b.alt_reads for b in parts])
    if any(b.amb_idx is not None for b in parts):
        idx, dec, ep, base = [], [], [], 0
        for b in parts:
            ai = b.amb_idx.copy() if b.amb_idx is not None else -np.ones(b.n_sites * 5, np.int32)
            ai[ai >= 0] += base
            idx.append(ai)
            if b.n_amb:
                dec.append(b.amb_dec)
                ep.append(b.amb_ep if b.amb_ep is not None else np.zeros(b.n_amb * P * 2))
                base += b.n_amb
        out["amb_idx"] = np.concatenate(idx)
        if dec:
            out["amb_dec"] = np.concatenate(dec)
            out["amb_ep"] = np.concatenate(ep)
    return Batch(P, **out)


class Results:
    """Owning numpy copy of a `crispgpu_results` block."""

    def __init__(self, n_sites: int, n_pools: int, max_calls: Optional[int] = None):
        self.n_sites, self.n_pools = n_sites, n_pools
        self.n_calls = 0
        self.max_calls = n_sites * 5 if max_calls is None else max_calls
        for name, dt, w in _RESULT_FIELDS:
            setattr(self, name, np.zeros((n_sites, w) if w > 1 else (n_sites,), dtype=dt))
        self.ac = np.zeros((self.max_calls, n_pools), np.int32)
        self.qv = np.zeros((self.max_calls, n_pools), np.float64)

    @classmethod
    def from_c(cls, cres: CResults, n_pools: int) -> "Results":
        n = cres.n_sites
        r = cls(n, n_pools, max_calls=max(cres.n_calls, 0))
        r.n_calls = cres.n_calls
        for name, dt, w in _RESULT_FIELDS:
            ptr = getattr(cres, name)
            dst = getattr(r, name)
            if ptr and dst.size:
                C.memmove(dst.ctypes.data, ptr, dst.nbytes)
        if r.n_calls:
            C.memmove(r.ac.ctypes.data, cres.ac, r.ac.nbytes)
            C.memmove(r.qv.ctypes.data, cres.qv, r.qv.nbytes)
        return r

    def called_sites(self) -> np.ndarray:
        return np.nonzero(self.varalleles >= 1)[0]
