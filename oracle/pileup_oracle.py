"""TEST INFRASTRUCTURE ONLY — numpy restatement of the reference's pileup for substitution-only alignments.

What the reference's front end computes for every reference position before it calls into the statistics engine,
restated on the struct-of-arrays alignment set the device pileup consumes (crisp_b200.pileup.Alignments):

  parse_cigar            parsebam/bamread.c:10-76     mismatches of the match block against the reference (either case; 'N'
                                                      in the read never counts), alignedbases, lastpos
  advance_read           parsebam/variant.c:56-131    a read covers p when position <= p < lastpos (one 'M' block here)
  calculate_allelecounts parsebam/variant.c:225-342   readdepths, MQcounts, filteredreads[0..1], the MAX_MM / MINQ / MIN_M
                                                      filters, counts and indcounts by BTI[base] + 8 * strand
  sort_allelecounts      parsebam/allelecounts.c:134-183   maxvec (reference allele first, exchange sort, IUPAC off)
  jointvariantcaller     variantcalls.c:44-92         potential-variant score and the candidate rule (>= 2, reference in ACGT)
  host/crisp_shim.c                                   the packed read stream and alt-read lists of a candidate site

Pinned against the reference itself: tests/test_pileup.py compares it with the batches the reference's own front end
(oracle/_ref/CRISP.gpu in dump mode: the unmodified parsebam/ objects) packs from the same BAM files, and with the
committed fixture tests/golden/pileup/frontend.npz made from such a dump by scripts/make_pileup_golden.py.
Only tests/ may import this module.
"""
from __future__ import annotations

import numpy as np

NT16 = np.frombuffer(b"=ACMGRSVTWYHKDBN", dtype=np.uint8)
# BTI (variant.c:17-26) of the sixteen characters above
NT16_BTI = np.array([0, 0, 1, 0, 2, 0, 1, 0, 3, 0, 1, 0, 2, 0, 1, 0], dtype=np.int64)


def _upper(a: np.ndarray) -> np.ndarray:
    a = a.copy()
    m = (a >= ord("a")) & (a <= ord("z"))
    a[m] -= 32
    return a


def pileup_oracle(aln, refseq: np.ndarray, ref_start: int, start: int, end: int, ploidy, min_q: int = 13, min_m: int = 20) -> dict:
    """Candidate sites of [start, end) with every array crispgpu_pileup_run returns (+ the streams)."""
    P = aln.n_pools
    W = end - start
    ploidy = np.asarray(ploidy)
    ref = _upper(np.asarray(refseq, dtype=np.uint8))

    def ref_at(p):
        k = p - ref_start
        ok = (k >= 0) & (k < ref.shape[0])
        return np.where(ok, ref[np.clip(k, 0, max(ref.shape[0] - 1, 0))], ord("N")) if ref.shape[0] else np.full(p.shape, ord("N"))

    ind = np.zeros((W, P, 24), np.int64)
    depth = np.zeros((W, P), np.int64)
    mqc = np.zeros((W, 4), np.int64)
    filt = np.zeros((W, 2), np.int64)
    per_pool = []
    rec = aln.rec
    for j in range(P):
        r0, r1 = int(aln.pool_off[j]), int(aln.pool_off[j + 1])
        rr = rec[r0:r1]
        n = r1 - r0
        ml = rr["mlen"].astype(np.int64)
        tot = int(ml.sum())
        ridx = np.repeat(np.arange(n), ml)
        k = np.arange(tot) - np.repeat(np.cumsum(ml) - ml, ml)            # index inside the match block
        bi = rr["base_off"].astype(np.int64)[ridx] + rr["clip"].astype(np.int64)[ridx] + k
        code = (aln.seq4[bi >> 1] >> np.where(bi & 1, 0, 4)) & 15
        q = aln.qual[bi].astype(np.int64)
        p = rr["pos"].astype(np.int64)[ridx] + k
        ch = NT16[code]
        refc = ref_at(p)
        mism = (ch != refc) & (ch != ord("N"))                              # bamread.c:33 (the reference is upper-cased here)
        mm = np.bincount(ridx, weights=mism, minlength=n).astype(np.int64)
        maxmm = ml // 40 + 2                                                # variant.c:286, no gaps
        mq = rr["mapq"].astype(np.int64)
        mq_ok = mq >= min_m
        read_ok = mq_ok & (mm <= maxmm)                                     # variant.c:299
        read_ok_ref = mq_ok & (mm <= maxmm - 1)                             # variant.c:302
        inwin = (p >= start) & (p < end)
        x = (p - start)[inwin]
        rsel = ridx[inwin]
        np.add.at(depth[:, j], x, 1)                                        # variant.c:280
        mclass = np.where(mq < 10, 0, np.where(mq >= 40, 3, mq // 20 + 1))  # variant.c:282-283
        np.add.at(mqc, (x, mclass[rsel]), 1)
        f0 = (mq_ok & (mm > maxmm))[rsel]                                   # variant.c:288-289
        f1 = ~f0 & ((q[inwin] < min_q) & mq_ok[rsel])
        np.add.at(filt[:, 0], x[f0], 1)
        np.add.at(filt[:, 1], x[f1], 1)
        counted = (q >= min_q) & read_ok[ridx] & ((ch != refc) | read_ok_ref[ridx])
        cw = counted & inwin
        strand = rr["strand"].astype(np.int64)[ridx]
        allele = NT16_BTI[code] + 8 * (strand != 0)
        np.add.at(ind[:, j, :], ((p - start)[cw], allele[cw]), 1)
        per_pool.append(dict(p=p, counted=counted, code=code, q=q, ridx=ridx, rr=rr, k=k, strand=strand))
    counts = ind.sum(axis=1)
    refw = ref_at(np.arange(start, end))
    refcode = np.full(W, 255, np.int64)
    for c, ch in enumerate(b"ACGT"):
        refcode[refw == ch] = c
    tot = ind[:, :, 0:8] + ind[:, :, 8:16] + ind[:, :, 16:24]              # [W][P][8]
    score = np.where(tot >= 3, 2, np.where((tot >= 1) & (ploidy[None, :, None] == 2), tot, 0))
    nonref = np.ones((W, 8), bool)
    ok = refcode < 8
    nonref[np.nonzero(ok)[0], refcode[ok]] = False
    potential = (score * nonref[:, None, :]).sum(axis=(1, 2))               # variantcalls.c:76-92 (PICALL == 3)
    cand = np.nonzero((potential >= 2) & (refcode < 4))[0]
    n = cand.shape[0]
    al = np.zeros((n, 8), np.uint8)
    ct = np.zeros((n, 8), np.int32)
    for s, x in enumerate(cand):                                            # allelecounts.c:134-183
        c8 = (counts[x, 0:8] + counts[x, 8:16] + counts[x, 16:24]).tolist()
        a8 = list(range(8))
        r = int(refcode[x])
        if 1 <= r < 8:
            c8[0], c8[r] = c8[r], c8[0]
            a8[0], a8[r] = a8[r], a8[0]
        for i in range(1, 7):
            for jj in range(i + 1, 8):
                if c8[i] < c8[jj]:
                    c8[i], c8[jj] = c8[jj], c8[i]
                    a8[i], a8[jj] = a8[jj], a8[i]
        al[s], ct[s] = a8, c8
    # streams of the candidate sites: pool-major, BAM order inside a pool (host/crisp_shim.c)
    read_off = [0]
    reads = []
    alt_lists = [[[] for _ in range(5)] for _ in range(n)]
    where = {int(x): s for s, x in enumerate(cand)}
    site_reads = [[None] * P for _ in range(n)]
    for j in range(P):
        d = per_pool[j]
        sel = d["counted"] & np.isin(d["p"] - start, cand)
        order = np.nonzero(sel)[0]                                           # already (read, base) = BAM order per position
        xs = (d["p"] - start)[order]
        for x in np.unique(xs):
            s = where[int(x)]
            idx = order[xs == x]
            base = NT16_BTI[d["code"][idx]]
            val = ((d["q"][idx] + 33) & 0xFF) | (base << 8) | ((d["strand"][idx] != 0).astype(np.int64) << 10)
            site_reads[s][j] = val.astype(np.uint16)
            for kslot in range(5):
                if ct[s, kslot + 1] < 3:
                    continue
                m = (base == al[s, kslot + 1]) & (d["code"][idx] == (1 << base))   # characters are compared: allelecounts.c:55
                rr = d["rr"][d["ridx"][idx[m]]]
                off = rr["clip"].astype(np.int64) + d["k"][idx[m]]
                for o, a, st in zip(off, rr["mlen"], d["strand"][idx[m]]):
                    alt_lists[s][kslot].append((j, int(o), int(a), int(st != 0), 0))
    from crisp_b200.abi import ALTREAD_DTYPE
    for s in range(n):
        for j in range(P):
            v = site_reads[s][j]
            if v is not None:
                reads.append(v)
            read_off.append(read_off[-1] + (0 if v is None else v.shape[0]))
    alt_off = [0]
    alt = []
    for s in range(n):
        for kslot in range(5):
            alt.extend(alt_lists[s][kslot])
            alt_off.append(len(alt))
    return dict(
        n_sites=n, pos=(cand + start).astype(np.int32), refbase=refcode[cand].astype(np.uint8), maxvec_al=al, maxvec_ct=ct,
        counts=counts[cand].astype(np.int32), indcounts=ind[cand].astype(np.int32), readdepths=depth[cand].astype(np.int32),
        mqcounts=mqc[cand].astype(np.int32), filtered=filt[cand].astype(np.int32), read_off=np.array(read_off, np.uint32),
        reads=np.concatenate(reads) if reads else np.zeros(0, np.uint16), alt_off=np.array(alt_off, np.uint32),
        alt_reads=np.array(alt, dtype=ALTREAD_DTYPE) if alt else np.zeros(0, ALTREAD_DTYPE), potential=potential)
