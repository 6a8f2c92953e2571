"""Comparison helpers shared by the parity tests."""
from __future__ import annotations

import numpy as np

from crisp_b200.abi import SLOT_CALLED, SLOT_NOCALL, SLOT_REJECT_AF, SLOT_REJECT_CT, SLOT_UNTESTED

REL_TOL = 1e-9  # north_star: likelihood-ratio statistics, EM frequencies and p-values within 1e-9 relative

INT_FIELDS = ["varalleles", "strandpass", "status", "allele", "call_rank", "varpools", "allelecount", "variantpools"]
FP_FIELDS = ["paf", "ctpval", "chi2pval", "af_ml", "delta", "deltaf", "hwe", "qvpf", "qvpr",
             "assoc_p", "assoc_llr", "assoc_or", "assoc_p1", "assoc_llr1", "assoc_or1"]


def rel_err(x, y):
    x = np.asarray(x, dtype=np.float64)
    y = np.asarray(y, dtype=np.float64)
    with np.errstate(invalid="ignore"):
        e = np.abs(x - y) / np.maximum(1.0, np.maximum(np.abs(x), np.abs(y)))
    # NaN/inf are results too (e.g. the reference takes the log of a negative likelihood ratio): equal when both are
    same_special = (np.isnan(x) & np.isnan(y)) | (np.isinf(x) & np.isinf(y) & (np.sign(x) == np.sign(y)))
    e = np.where(same_special, 0.0, e)
    return np.where(np.isnan(e), np.inf, e)


def rows_of(res):
    """AC/QV of every called (site, slot), keyed by flat slot index, independent of row order."""
    out = {}
    st = res.status.reshape(-1)
    cr = res.call_row.reshape(-1)
    for o in np.nonzero(st == SLOT_CALLED)[0]:
        if cr[o] >= 0:
            out[int(o)] = (res.ac[cr[o]], res.qv[cr[o]])
    return out


def compare_results(got, want, fp_skip=(), int_skip=(), tol=REL_TOL, allow_flips=0, label=""):
    """Assert `got` (engine) equals `want` (checker): integers bit-exact, doubles within tol.  Slots whose decision
    flipped are reported; up to `allow_flips` are tolerated (threshold-adjacent sites)."""
    msgs = []
    flips = np.nonzero(got.status.reshape(-1) != want.status.reshape(-1))[0]
    if flips.size > allow_flips:
        o = flips[:8]
        msgs.append(f"{label}: {flips.size} slot decisions differ, e.g. slots {o.tolist()} got {got.status.reshape(-1)[o].tolist()} want {want.status.reshape(-1)[o].tolist()}")
    same = np.ones(got.status.size, bool)
    same[flips] = False
    same_site = np.ones(got.n_sites, bool)
    same_site[np.unique(flips // 5)] = False
    for name in INT_FIELDS:
        if name in int_skip:
            continue
        g, w = getattr(got, name), getattr(want, name)
        m = same_site if g.ndim == 1 else same.reshape(g.shape)
        bad = np.nonzero((g != w) & m)
        if bad[0].size:
            msgs.append(f"{label}: {name} differs at {bad[0].size} places, first {[b[0] for b in bad]}: {g[tuple(b[0] for b in bad)]} vs {w[tuple(b[0] for b in bad)]}")
    for name in FP_FIELDS:
        if name in fp_skip:
            continue
        g, w = getattr(got, name), getattr(want, name)
        e = rel_err(g, w) * same.reshape(g.shape)
        if e.max(initial=0) > tol:
            i = np.unravel_index(np.argmax(e), e.shape)
            msgs.append(f"{label}: {name} max rel err {e.max():.3e} at {i}: {g[i]!r} vs {w[i]!r}")
    gr, wr = rows_of(got), rows_of(want)
    for o in gr:
        if o not in wr or not same[o]:
            continue
        if not np.array_equal(gr[o][0], wr[o][0]):
            msgs.append(f"{label}: AC differs at slot {o}: {gr[o][0].tolist()} vs {wr[o][0].tolist()}")
            break
        e = rel_err(gr[o][1], wr[o][1]).max()
        if e > tol:
            msgs.append(f"{label}: QV max rel err {e:.3e} at slot {o}")
            break
    assert not msgs, "\n".join(msgs)
    return int(flips.size)


def batch_from_counts(ic, refbase, quality=30):
    """A consistent batch for hand-made per-pool allele counts ic[n][P][24] (forward / reverse entries only): one read of the
    given quality per counted base, alleles sorted as the reference does."""
    from crisp_b200.abi import Batch
    from crisp_b200.synth import sort_alleles
    ic = np.asarray(ic, dtype=np.int32)
    n, P, _ = ic.shape
    assert ic[:, :, 16:].sum() == 0
    refbase = np.asarray(refbase, dtype=np.uint8)
    counts = ic.sum(axis=1)
    al, ct = sort_alleles(counts.reshape(n, 3, 8).sum(axis=1), refbase)
    reads, per = [], np.zeros(n * P, np.int64)
    for s in range(n):
        for i in range(P):
            for idx in np.nonzero(ic[s, i, :16])[0]:
                base, rev = int(idx) % 8, int(idx) // 8
                w = ((quality + 33) & 0xFF) | ((base & 3) << 8) | (rev << 10)
                reads.extend([w] * int(ic[s, i, idx]))
                per[s * P + i] += int(ic[s, i, idx])
    return Batch(P, tid=np.zeros(n, np.int32), pos=1000 + 100 * np.arange(n, dtype=np.int32), refbase=refbase, maxvec_al=al.reshape(-1),
                 maxvec_ct=ct.reshape(-1), counts=counts.reshape(-1), indcounts=ic.reshape(-1), reads=np.asarray(reads, np.uint16),
                 read_off=np.concatenate([[0], np.cumsum(per)]).astype(np.uint32))
