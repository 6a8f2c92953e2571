"""Synthetic candidate-site batches (SURVEY.md §8d generative model).

Produces the struct-of-arrays batch the host front end would pack at variantcalls.c:115 for pooled data of a named
shape (P pools x poolsize n at a given depth): per-pool/per-strand allele counts, the packed per-read
(base, quality, strand, bidir) stream, the alt-read side list and — optionally — indel alleles with homopolymer
ambiguity rows.  Sites that do not satisfy the reference's candidate rule (some pool has >= 3 reads of a
non-reference allele, variantcalls.c:85,92) are dropped, so every emitted site is one the reference would hand to
newCRISPcaller.

Model: read length 100, 50/50 strand, base quality iid from {20,25,30,35,40} with probabilities
{.05,.10,.25,.35,.25}, errors at 10^(-Q/10) to a uniformly random other base; a fraction `variant_frac` of the
sites carries a planted variant with population AF ~ 1/(P n) + Exp(0.03) clipped to 1 and per-pool allele counts
~ Binomial(n, AF); the other sites are sequencing-error candidates (one pool forced to >= 3 error reads of one
allele, which is what conditioning the error process on the candidate rule amounts to).
"""
from __future__ import annotations

import numpy as np

from .abi import ALTREAD_DTYPE, Batch

QUALS = np.array([20, 25, 30, 35, 40], dtype=np.int64)
QUAL_P = np.array([0.05, 0.10, 0.25, 0.35, 0.25])


def sort_alleles(counts8: np.ndarray, refbase: np.ndarray):
    """maxvec of sort_allelecounts (parsebam/allelecounts.c:155-187): reference first, then the exchange sort
    `if ct[i] < ct[j] swap` over i=1..6, j=i+1..7, vectorised over sites."""
    n = counts8.shape[0]
    ct = counts8.astype(np.int64).copy()
    al = np.tile(np.arange(8, dtype=np.int64), (n, 1))
    rows = np.arange(n)
    r = refbase.astype(np.int64)
    # swap entry 0 with entry r when r >= 1
    c0, cr = ct[rows, 0].copy(), ct[rows, r].copy()
    ct[rows, 0], ct[rows, r] = cr, c0
    a0, ar = al[rows, 0].copy(), al[rows, r].copy()
    al[rows, 0], al[rows, r] = ar, a0
    for i in range(1, 7):
        for j in range(i + 1, 8):
            sw = ct[:, i] < ct[:, j]
            ci, cj = ct[:, i].copy(), ct[:, j].copy()
            ct[:, i] = np.where(sw, cj, ci)
            ct[:, j] = np.where(sw, ci, cj)
            ai, aj = al[:, i].copy(), al[:, j].copy()
            al[:, i] = np.where(sw, aj, ai)
            al[:, j] = np.where(sw, ai, aj)
    return al.astype(np.uint8), ct.astype(np.int32)


def synth_batch(
    n_sites: int,
    n_pools: int,
    ploidy,
    depth: float,
    seed: int,
    variant_frac: float = 0.15,
    bidir_frac: float = 0.0,
    indel_frac: float = 0.0,
    rare: bool = False,
    second_alt_frac: float = 0.0,
    tid: int = 0,
    pos_start: int = 1000,
    pos_step: int = 75,
    fixed_depth: bool = False,
    with_qhigh: bool = False,
    with_stats: bool = False,
    quals=None,
) -> Batch:
    rng = np.random.default_rng(seed)
    P = int(n_pools)
    ploidy = np.broadcast_to(np.asarray(ploidy, dtype=np.int64), (P,)).copy()
    n = int(n_sites)
    refbase = rng.integers(0, 4, n)
    alt = (refbase + 1 + rng.integers(0, 3, n)) % 4
    alt2 = (refbase + 1 + ((alt - refbase - 1) % 4 + 1 + rng.integers(0, 2, n)) % 3) % 4  # a third base
    is_var = rng.random(n) < variant_frac
    has_alt2 = is_var & (rng.random(n) < second_alt_frac)
    K = int(ploidy.sum())
    if rare:
        # rare-variant enriched (BASELINE config 4): AC in {1,2,3} in 1-2 pools
        ac = np.zeros((n, P), dtype=np.int64)
        npools = rng.integers(1, 3, n)
        for t in range(2):
            pool = rng.integers(0, P, n)
            use = is_var & (npools > t)
            np.add.at(ac, (np.nonzero(use)[0], pool[use]), rng.integers(1, 4, int(use.sum())))
        ac = np.minimum(ac, ploidy[None, :])
    else:
        af = np.clip(1.0 / K + rng.exponential(0.03, n), 0, 1)
        ac = rng.binomial(ploidy[None, :], np.where(is_var, af, 0.0)[:, None])
    af2 = np.clip(1.0 / K + rng.exponential(0.02, n), 0, 0.5)
    ac2 = rng.binomial(np.maximum(ploidy[None, :] - ac, 0), np.where(has_alt2, af2, 0.0)[:, None])

    d = np.full((n, P), int(depth), dtype=np.int64) if fixed_depth else rng.poisson(depth, (n, P))
    cell = d.reshape(-1)
    R = int(cell.sum())
    cell_of_read = np.repeat(np.arange(n * P, dtype=np.int64), cell)
    site_of_read = cell_of_read // P
    if quals is None:
        q = QUALS[rng.choice(5, size=R, p=QUAL_P)]
    else:  # e.g. range(14, 42): the many-valued quality spectrum of older instruments
        q = np.asarray(list(quals), dtype=np.int64)[rng.integers(0, len(quals), R)]
    strand = (rng.random(R) < 0.5).astype(np.int64)
    bidir = (rng.random(R) < bidir_frac).astype(np.int64) if bidir_frac > 0 else np.zeros(R, np.int64)
    u = rng.random(R)
    fa = (ac / ploidy[None, :]).reshape(-1)[cell_of_read]
    fb = (ac2 / ploidy[None, :]).reshape(-1)[cell_of_read]
    base = np.where(u < fa, alt[site_of_read], np.where(u < fa + fb, alt2[site_of_read], refbase[site_of_read]))
    err = rng.random(R) < 10.0 ** (-q / 10.0)
    base = np.where(err, (base + 1 + rng.integers(0, 3, R)) % 4, base)
    # error candidates: one pool forced to k >= 3 error reads of one allele
    read_off = np.concatenate([[0], np.cumsum(cell)])
    ns = np.nonzero(~is_var)[0]
    if ns.size:
        pool = rng.integers(0, P, ns.size)
        k = 3 + rng.poisson(0.35, ns.size)
        c = ns * P + pool
        start, ln = read_off[c], cell[c]
        k = np.minimum(k, ln)
        for j in range(int(k.max(initial=0))):
            m = (j < k) & (ln > 0)
            kk = np.maximum(k, 1)
            seg = np.maximum(ln // kk, 1)
            idx = start + (j * ln) // kk + (rng.random(ns.size) * seg).astype(np.int64)
            idx = np.minimum(idx, start + np.maximum(ln - 1, 0))
            base[idx[m]] = alt[ns[m]]

    cls = np.where(bidir == 1, 2, strand)
    key = cell_of_read * 24 + base + 8 * cls
    indcounts = np.bincount(key, minlength=n * P * 24).astype(np.int32).reshape(n, P, 24)

    # optional indel allele (slot allele 4): counts only, not part of the packed stream (type != 0 reads)
    indel_len = np.zeros((n, 8), dtype=np.int16)
    hplen = np.zeros((n, 8), dtype=np.int16)
    is_indel = rng.random(n) < indel_frac if indel_frac > 0 else np.zeros(n, bool)
    ind_reads = None
    if is_indel.any():
        ii = np.nonzero(is_indel)[0]
        sign = np.where(rng.random(ii.size) < 0.5, -1, 1)
        indel_len[ii, 4] = sign * rng.integers(1, 4, ii.size)
        hplen[ii, 4] = np.where(rng.random(ii.size) < 0.6, rng.integers(1, 9, ii.size), 0)
        iaf = np.clip(1.0 / K + rng.exponential(0.05, ii.size), 0, 1)
        iac = rng.binomial(ploidy[None, :], iaf[:, None])
        lam = d[ii] * (iac / ploidy[None, :]) + 0.02
        nf = rng.poisson(lam * 0.5)
        nr = rng.poisson(lam * 0.5)
        indcounts[ii, :, 4] = nf
        indcounts[ii, :, 12] = nr
        ind_reads = (ii, nf, nr)

    counts = indcounts.sum(axis=1, dtype=np.int64).astype(np.int32)
    tot8 = counts.reshape(n, 3, 8).sum(axis=1)
    # candidate rule, variantcalls.c:80-92
    per = indcounts.reshape(n, P, 3, 8).sum(axis=2)  # [n][P][8]
    nonref = np.ones((n, 8), bool)
    nonref[np.arange(n), refbase] = False
    cand = ((per >= 3) & nonref[:, None, :]).any(axis=(1, 2))

    maxvec_al, maxvec_ct = sort_alleles(tot8, refbase)

    # alt-read side list: for slot k (allele maxvec_al[k+1], ct >= 3): reads whose base is that allele
    offs = rng.integers(0, 100, R)
    slot_of_allele = -np.ones((n, 8), dtype=np.int64)
    for ksl in range(5):
        ok = maxvec_ct[:, ksl + 1] >= 3
        slot_of_allele[np.nonzero(ok)[0], maxvec_al[ok, ksl + 1]] = ksl
    rslot = slot_of_allele[site_of_read, base]
    sel = np.nonzero((rslot >= 0) & cand[site_of_read])[0]
    a_site, a_slot = site_of_read[sel], rslot[sel]
    a_pool = cell_of_read[sel] % P
    a_off, a_strand = offs[sel], strand[sel]
    a_order = np.arange(sel.size)
    if ind_reads is not None:
        ii, nf, nr = ind_reads
        for st, cnt in ((0, nf), (1, nr)):
            tot = cnt.reshape(-1)
            srep = np.repeat(np.repeat(ii, P), tot)
            prep = np.repeat(np.tile(np.arange(P), ii.size), tot)
            keep = (slot_of_allele[srep, 4] >= 0) & cand[srep]
            srep, prep = srep[keep], prep[keep]
            a_site = np.concatenate([a_site, srep])
            a_slot = np.concatenate([a_slot, slot_of_allele[srep, 4]])
            a_pool = np.concatenate([a_pool, prep])
            a_off = np.concatenate([a_off, rng.integers(0, 100, srep.size)])
            a_strand = np.concatenate([a_strand, np.full(srep.size, st)])
            a_order = np.concatenate([a_order, np.arange(srep.size) + a_order.size + st * 10**9])

    # ambiguity rows for indel slots with HPlength > 0 (what remove_ambiguous_bases would subtract)
    amb_idx = -np.ones((n, 5), dtype=np.int32)
    amb_rows, amb_eps = [], []
    if is_indel.any():
        for s in np.nonzero(is_indel & cand)[0]:
            ksl = slot_of_allele[s, 4]
            hp = hplen[s, 4]
            if ksl < 0 or hp <= 0:
                continue
            rb = refbase[s]
            avail = np.stack([indcounts[s, :, rb], indcounts[s, :, rb + 8]], axis=1).astype(np.int64)
            dec = rng.binomial(avail, min(0.9, 0.01 * hp))
            row = np.zeros((P, 4), np.int32)
            row[:, :2] = dec
            # sum of error probabilities of the removed reads = the first dec[i][st] reference reads of that strand
            eps = np.zeros((P, 2))
            for i in range(P):
                lo, hi = read_off[s * P + i], read_off[s * P + i + 1]
                for st in range(2):
                    if dec[i, st] == 0:
                        continue
                    m = np.nonzero((base[lo:hi] == rb) & (strand[lo:hi] == st) & (bidir[lo:hi] == 0))[0][: dec[i, st]]
                    eps[i, st] = (10.0 ** (-q[lo:hi][m] / 10.0)).sum()
            amb_idx[s, ksl] = len(amb_rows)
            amb_rows.append(row)
            amb_eps.append(eps)

    # drop non-candidate sites
    keep_sites = np.nonzero(cand)[0]
    m = keep_sites.size
    remap = -np.ones(n, dtype=np.int64)
    remap[keep_sites] = np.arange(m)
    kr = cand[site_of_read]
    packed = ((q + 33) | (base << 8) | (strand << 10) | (bidir << 11)).astype(np.uint16)[kr]
    cell_keep = cand.repeat(P)
    new_off = np.concatenate([[0], np.cumsum(cell[cell_keep])]).astype(np.uint32)

    order = np.lexsort((a_order, a_pool, a_slot, remap[a_site]))
    alt_reads = np.zeros(order.size, dtype=ALTREAD_DTYPE)
    alt_reads["pool"] = a_pool[order]
    alt_reads["offset"] = a_off[order]
    alt_reads["aligned"] = 100
    alt_reads["strand"] = a_strand[order]
    akey = remap[a_site[order]] * 5 + a_slot[order]
    alt_off = np.concatenate([[0], np.cumsum(np.bincount(akey, minlength=m * 5))]).astype(np.uint32)

    arrays = dict(
        tid=np.full(m, tid, np.int32),
        pos=(pos_start + keep_sites * pos_step).astype(np.int32),
        refbase=refbase[keep_sites].astype(np.uint8),
        maxvec_al=maxvec_al[keep_sites].reshape(-1),
        maxvec_ct=maxvec_ct[keep_sites].reshape(-1),
        counts=counts[keep_sites].reshape(-1),
        indcounts=indcounts[keep_sites].reshape(-1),
        read_off=new_off,
        reads=packed,
        alt_off=alt_off,
        alt_reads=alt_reads,
    )
    if is_indel.any():
        arrays["indel_len"] = indel_len[keep_sites].reshape(-1)
        arrays["hplen"] = hplen[keep_sites].reshape(-1)
        arrays["amb_idx"] = amb_idx[keep_sites].reshape(-1)
        if amb_rows:
            arrays["amb_dec"] = np.stack(amb_rows).reshape(-1)
            arrays["amb_ep"] = np.stack(amb_eps).astype(np.float64).reshape(-1)
    if with_qhigh or with_stats:
        # Qhighcounts / stats as the pileup accumulates them by strand (parsebam/variant.c:318-320), USE_QV = 0
        skey = cell_of_read * 16 + base + 8 * strand
        if with_qhigh:
            qh = np.bincount(skey, minlength=n * P * 16).astype(np.float64).reshape(n, P, 16)
            if is_indel.any():
                qh[:, :, 4] = indcounts[:, :, 4]
                qh[:, :, 12] = indcounts[:, :, 12]
            arrays["qhigh"] = qh[keep_sites].reshape(-1)
        if with_stats:
            ep = 10.0 ** (-q / 10.0)
            st = np.bincount(skey, weights=ep, minlength=n * P * 16).reshape(n, P, 16)
            arrays["stats"] = st[keep_sites].reshape(-1)
            arrays["tstats"] = st[keep_sites].sum(axis=1).reshape(-1)
    return Batch(P, **arrays)


def survey_demo_batch() -> Batch:
    """The 6-pool known-answer site of SURVEY.md §8c level 3 (pool-major xorshift64 qualities, alt G in the
    first {7,0,0,0,0,1} reads of pools 0..5, ref A, strand = read index & 1)."""
    P, reads_per_pool = 6, 100
    state = 88172645463325252
    mask = (1 << 64) - 1
    nalt = [7, 0, 0, 0, 0, 1]
    packed = []
    indcounts = np.zeros((1, P, 24), np.int32)
    alt_entries = []
    for i in range(P):
        for j in range(reads_per_pool):
            state ^= (state << 13) & mask
            state ^= state >> 7
            state ^= (state << 17) & mask
            qv = 20 + 5 * (((state >> 11) & 0xFFFFFFFF) % 5)
            base = 2 if j < nalt[i] else 0
            st = j & 1
            packed.append((qv + 33) | (base << 8) | (st << 10))
            indcounts[0, i, base + 8 * st] += 1
            if base == 2:
                alt_entries.append((i, 50, 100, st, 0))
    counts = indcounts.sum(axis=1)
    tot8 = counts.reshape(1, 3, 8).sum(axis=1)
    al, ct = sort_alleles(tot8, np.array([0]))
    return Batch(
        P,
        tid=[0], pos=[999], refbase=[0], maxvec_al=al.reshape(-1), maxvec_ct=ct.reshape(-1),
        counts=counts.reshape(-1), indcounts=indcounts.reshape(-1),
        read_off=np.arange(P + 1, dtype=np.uint32) * reads_per_pool, reads=np.array(packed, np.uint16),
        alt_off=np.array([0, len(alt_entries), len(alt_entries), len(alt_entries), len(alt_entries), len(alt_entries)], np.uint32),
        alt_reads=np.array(alt_entries, dtype=ALTREAD_DTYPE),
    )
