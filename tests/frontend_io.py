"""Reader for the batch dumps written by host/crisp_frontend.c (CRISPGPU_DUMP): what the reference's front end packed at
variantcalls.c:115 for every candidate site of a BAM run, as `crisp_b200.abi.Batch` objects plus the engine options the
front end derived from the reference's option globals."""
from __future__ import annotations

import struct

import numpy as np

from crisp_b200.abi import ALTREAD_DTYPE, Batch, EngineConfig

_HEAD = ["n_sites", "n_pools", "n_amb", "em_flag", "max_iter", "chisq_permutation", "fast_filter", "use_base_qvs", "min_reads",
         "qv_offset", "min_q", "pivot_sample", "varpoolsize", "use_qv", "association", "_pad"]
_DHEAD = ["thresh1", "min_qpv", "thresh3", "ser", "relax", "qmin", "agilent_ref_bias", "theta"]
_ARRAYS = [("ploidy", np.int32), ("tid", np.int32), ("pos", np.int32), ("refbase", np.uint8), ("maxvec_al", np.uint8),
           ("maxvec_ct", np.int32), ("counts", np.int32), ("indcounts", np.int32), ("indel_len", np.int16), ("hplen", np.int16),
           ("read_off", np.uint32), ("reads", np.uint16), ("alt_off", np.uint32), ("alt_reads", ALTREAD_DTYPE),
           ("amb_idx", np.int32), ("amb_dec", np.int32), ("amb_ep", np.float64), ("qhigh", np.float64), ("stats", np.float64),
           ("tstats", np.float64), ("bin_off", np.uint32), ("bins", np.uint32), ("ic_off", np.uint32), ("ic_sparse", np.uint32),
           ("readdepths", np.int32), ("mqcounts", np.int32), ("refb", np.uint8), ("filtered", np.int32)]


def read_dump(path: str):
    """[(EngineConfig, Batch, extras)] — one entry per flushed batch.  The dense `indcounts` table is always present
    (the front end keeps it for the printer), so a compact dump can be checked against its dense equivalent."""
    out = []
    with open(path, "rb") as f:
        data = f.read()
    o = 0
    while o < len(data):
        if data[o:o + 8] != b"CRSPBTC1":
            raise ValueError(f"bad magic at {o}")
        o += 8
        head = dict(zip(_HEAD, struct.unpack_from("<16i", data, o)))
        o += 64
        dhead = dict(zip(_DHEAD, struct.unpack_from("<8d", data, o)))
        o += 64
        arr = {}
        for name, dt in _ARRAYS:
            (n,) = struct.unpack_from("<Q", data, o)
            o += 8
            arr[name] = np.frombuffer(data, dtype=dt, count=n // np.dtype(dt).itemsize, offset=o).copy() if n else None
            o += (n + 7) & ~7
        P = head["n_pools"]
        opts = {k: head[k] for k in ("em_flag", "max_iter", "chisq_permutation", "fast_filter", "use_base_qvs", "min_reads",
                                     "qv_offset", "min_q", "pivot_sample", "varpoolsize", "use_qv", "association")}
        opts.update(dhead)
        cfg = EngineConfig(ploidy=arr.pop("ploidy"), options=opts)
        extras = {k: arr.pop(k) for k in ("readdepths", "mqcounts", "refb", "filtered")}
        n = head["n_sites"]
        if arr["alt_reads"] is None:
            arr["alt_reads"] = np.zeros(0, ALTREAD_DTYPE)
        if arr["reads"] is None and arr["bins"] is None:
            arr["reads"] = np.zeros(0, np.uint16)
        if head["n_amb"] == 0:
            arr["amb_dec"] = arr["amb_ep"] = None
        compact = arr["bins"] is not None
        dense = dict(arr)
        if compact:
            # the dump carries both the sparse and the dense counts: hand out the dense table with the bins
            dense["ic_off"] = dense["ic_sparse"] = None
        batch = Batch(P, **{k: v for k, v in dense.items() if v is not None})
        extras["compact"] = compact
        extras["ic_off"], extras["ic_sparse"] = arr["ic_off"], arr["ic_sparse"]
        assert batch.n_sites == n
        out.append((cfg, batch, extras))
    return out


def vcf_body(path: str):
    """Records of a VCF file as lists of columns (header lines dropped)."""
    with open(path) as f:
        return [line.rstrip("\n").split("\t") for line in f if line.strip() and not line.startswith("#")]
