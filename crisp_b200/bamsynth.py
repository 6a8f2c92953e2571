"""Synthetic multi-pool sequencing data as real files: reference FASTA + .fai and one coordinate-sorted BAM per pool.

The generative model is SURVEY.md §8d's (the same as crisp_b200.synth, which writes candidate batches directly): reads of
100 bases, 50/50 strand, MAPQ 60, base qualities iid from {20,25,30,35,40} with probabilities {.05,.10,.25,.35,.25},
sequencing errors at 10^(-Q/10) to a uniformly random other base, a variant planted every `spacing` bases with
population allele frequency 1/(P n) + Exp(0.03) clipped to 1, per-pool allele count ~ Binomial(n, AF), each read carrying
the variant with probability count/n.  A fraction of the planted variants are short deletions / insertions
(`indel_frac`), and `pe_frac` of the fragments are sequenced as overlapping mate pairs (what the reference's
OVERLAPPING_PE_READS logic, parsebam/variant.c:137-220, merges into its bidirectional strand class).

The files are what the reference's own front end consumes (samtools 0.1.18 BAM reader, bamsreader.c:124-185; .fai
reader, readfasta.c:33-108), so the same inputs drive `CRISP.binary` (the unmodified reference) and `CRISP.gpu` (the
reference's front end with the one call at variantcalls.c:115 bound to the B200 engine).  BGZF/BAM are written here
with numpy + zlib (no samtools needed on the GPU box); the .bai index, needed only for `--regions`, is built by
oracle/_ref/bamtool (the vendored library's bam_index_build).
"""
from __future__ import annotations

import os
import struct
import zlib

import numpy as np

READLEN = 100
QUALS = np.array([20, 25, 30, 35, 40], dtype=np.uint8)
QPROB = np.array([0.05, 0.10, 0.25, 0.35, 0.25])
BASES = np.frombuffer(b"ACGT", dtype=np.uint8)
NT16 = np.array([1, 2, 4, 8, 15], dtype=np.uint8)  # A C G T N in BAM's 4-bit code

_BGZF_EOF = bytes.fromhex("1f8b08040000000000ff0600424302001b0003000000000000000000")


def _bgzf_write(f, data: bytes, level: int = 1) -> None:
    """BGZF: independent gzip members of <= 0xff00 input bytes with the BC extra field (samtools/bgzf.c)."""
    for o in range(0, len(data), 0xFF00):
        chunk = data[o:o + 0xFF00]
        c = zlib.compressobj(level, zlib.DEFLATED, -15)
        body = c.compress(chunk) + c.flush()
        bsize = len(body) + 25  # total block size - 1
        f.write(struct.pack("<BBBBIBBHBBHH", 0x1F, 0x8B, 8, 4, 0, 0, 0xFF, 6, 0x42, 0x43, 2, bsize))
        f.write(body)
        f.write(struct.pack("<II", zlib.crc32(chunk) & 0xFFFFFFFF, len(chunk)))
    return None


def _reg2bin(beg: np.ndarray, end: np.ndarray) -> np.ndarray:
    """UCSC binning scheme of the BAM specification, vectorised (end exclusive)."""
    end = end - 1
    out = np.zeros(beg.shape, dtype=np.int64)
    done = np.zeros(beg.shape, dtype=bool)
    for shift, base in ((14, 4681), (17, 585), (20, 73), (23, 9), (26, 1)):
        hit = ~done & ((beg >> shift) == (end >> shift))
        out[hit] = base + (beg[hit] >> shift)
        done |= hit
    return out


def _pack_records(tid, pos, flag, mapq, cigars, seq, qual, names, mpos, isize):
    """BAM alignment records for reads that all have READLEN bases.  cigars: int array [n][3] of (len<<4|op), zero-padded;
    names: uint8 [n][L] NUL-terminated."""
    n = pos.shape[0]
    ncig = (cigars != 0).sum(axis=1).astype(np.int64)
    L = names.shape[1]
    reflen = np.zeros(n, dtype=np.int64)
    for k in range(cigars.shape[1]):
        op = cigars[:, k] & 0xF
        ln = cigars[:, k] >> 4
        reflen += np.where((cigars[:, k] != 0) & ((op == 0) | (op == 2) | (op == 3)), ln, 0)
    bins = _reg2bin(pos.astype(np.int64), pos.astype(np.int64) + reflen)
    size = 32 + L + 4 * ncig + (READLEN + 1) // 2 + READLEN  # without the leading block_size
    off = np.zeros(n + 1, dtype=np.int64)
    np.cumsum(size + 4, out=off[1:])
    out = np.zeros(int(off[-1]), dtype=np.uint8)
    packed = (NT16[seq[:, 0::2]] << 4) | NT16[seq[:, 1::2]]
    for nc in np.unique(ncig):
        m = np.nonzero(ncig == nc)[0]
        k = m.shape[0]
        rec = np.zeros(k, dtype=np.dtype([
            ("block_size", "<i4"), ("tid", "<i4"), ("pos", "<i4"), ("l_name", "u1"), ("mapq", "u1"), ("bin", "<u2"),
            ("ncig", "<u2"), ("flag", "<u2"), ("l_seq", "<i4"), ("mtid", "<i4"), ("mpos", "<i4"), ("isize", "<i4"),
            ("name", "u1", (L,)), ("cigar", "<u4", (int(nc),)), ("seq", "u1", ((READLEN + 1) // 2,)), ("qual", "u1", (READLEN,))]))
        rec["block_size"] = size[m]
        rec["tid"] = tid
        rec["pos"] = pos[m]
        rec["l_name"] = L
        rec["mapq"] = mapq[m] if isinstance(mapq, np.ndarray) else mapq
        rec["bin"] = bins[m]
        rec["ncig"] = nc
        rec["flag"] = flag[m]
        rec["l_seq"] = READLEN
        rec["mtid"] = np.where(mpos[m] >= 0, tid, -1)
        rec["mpos"] = mpos[m]
        rec["isize"] = isize[m]
        rec["name"] = names[m]
        rec["cigar"] = cigars[m, :int(nc)]
        rec["seq"] = packed[m]
        rec["qual"] = qual[m]
        raw = rec.view(np.uint8).reshape(k, -1)
        idx = off[m][:, None] + np.arange(raw.shape[1])[None, :]
        out[idx] = raw
    return out.tobytes()


def write_reference(path: str, name: str, seq: np.ndarray, width: int = 60) -> None:
    """FASTA + .fai (name, length, offset of the first base, bases per line, bytes per line)."""
    L = seq.shape[0]
    header = f">{name}\n".encode()
    with open(path, "wb") as f:
        f.write(header)
        body = seq.tobytes()
        f.write(b"\n".join(body[o:o + width] for o in range(0, L, width)) + b"\n")
    with open(path + ".fai", "w") as f:
        f.write(f"{name}\t{L}\t{len(header)}\t{width}\t{width + 1}\n")


def make_dataset(outdir: str, pools: int, poolsize: int, length: int, depth: float, seed: int = 1, spacing: int = 500,
                 indel_frac: float = 0.0, pe_frac: float = 0.0, chrom: str = "chr1", qual_values=None, qual_probs=None,
                 lowmapq_frac: float = 0.0, noisy_frac: float = 0.0, clip_frac: float = 0.0, n_frac: float = 0.0) -> dict:
    """Write ref.fa(.fai), pool_XX.bam and bams.list under outdir; return the planted truth.

    What the read filters of calculate_allelecounts (parsebam/variant.c:286-302) look at can be switched on: `lowmapq_frac` of
    the reads get a mapping quality from {0, 5, 15, 25, 37}, `noisy_frac` carry 3-8 extra mismatches (the MAX_MM filter),
    `clip_frac` of the gap-free single-end reads are soft-clipped at one or both ends, `n_frac` of the bases read 'N'."""
    os.makedirs(outdir, exist_ok=True)
    rng = np.random.default_rng(seed)
    qv = QUALS if qual_values is None else np.asarray(qual_values, dtype=np.uint8)
    qp = QPROB if qual_probs is None else np.asarray(qual_probs, dtype=float)
    ref_codes = rng.integers(0, 4, size=length, dtype=np.uint8)
    ref = BASES[ref_codes]
    fasta = os.path.join(outdir, "ref.fa")
    write_reference(fasta, chrom, ref)

    vpos = np.arange(spacing // 2, length - 2 * READLEN, spacing)
    nv = vpos.shape[0]
    af = np.minimum(1.0 / (pools * poolsize) + rng.exponential(0.03, size=nv), 1.0)
    alt = (ref_codes[vpos] + rng.integers(1, 4, size=nv)) % 4
    kind = np.zeros(nv, dtype=np.int8)  # 0 SNV, 1 deletion, 2 insertion
    if indel_frac > 0:
        u = rng.random(nv)
        kind[u < indel_frac] = 1
        kind[u < indel_frac / 2] = 2
    ilen = rng.integers(1, 4, size=nv)  # indel length 1..3
    acount = rng.binomial(poolsize, af[None, :], size=(pools, nv))

    header_text = f"@HD\tVN:1.0\tSO:coordinate\n@SQ\tSN:{chrom}\tLN:{length}\n".encode()
    bam_header = b"BAM\x01" + struct.pack("<i", len(header_text)) + header_text + struct.pack("<i", 1) + \
        struct.pack("<i", len(chrom) + 1) + chrom.encode() + b"\0" + struct.pack("<i", length)

    bam_paths = []
    for p in range(pools):
        nfrag = int(round(length * depth / READLEN))
        n_pe = int(round(nfrag * pe_frac / 2))  # each overlapping pair contributes two reads
        n_se = nfrag - 2 * n_pe
        start = rng.integers(0, length - READLEN - 8, size=n_se + 2 * n_pe)
        rev = rng.random(n_se + 2 * n_pe) < 0.5
        flag = np.where(rev, 16, 0).astype(np.int64)
        mpos = np.full(start.shape[0], -1, dtype=np.int64)
        isize = np.zeros(start.shape[0], dtype=np.int64)
        fragment = np.arange(start.shape[0])
        if n_pe:
            a = np.arange(n_se, n_se + n_pe)
            b = a + n_pe
            shift = rng.integers(20, 60, size=n_pe)  # mates overlap by 40-80 bases
            start[b] = np.minimum(start[a] + shift, length - READLEN - 8)
            fragment[b] = a
            # first mate forward, second mate reverse: 99 / 147
            flag[a] = 1 | 2 | 32 | 64
            flag[b] = 1 | 2 | 16 | 128
            mpos[a] = start[b]
            mpos[b] = start[a]
            isize[a] = start[b] + READLEN - start[a]
            isize[b] = -isize[a]
        n = start.shape[0]
        # which planted variant a read covers (at most one: spacing >= 2 * READLEN)
        v = np.clip((start + READLEN - 1 - spacing // 2) // spacing, 0, nv - 1)
        covers = (vpos[v] >= start + 2) & (vpos[v] <= start + READLEN - 6)
        # haplotype choice is per fragment, so overlapping mates agree
        hap_u = rng.random(n)[fragment]
        carries = covers & (hap_u < acount[p, v] / poolsize)
        offs = np.arange(READLEN)[None, :]
        cig = np.zeros((n, 3), dtype=np.int64)
        cig[:, 0] = (READLEN << 4) | 0
        refidx = start[:, None] + offs
        isdel = carries & (kind[v] == 1)
        isins = carries & (kind[v] == 2)
        k = (vpos[v] - start + 1)  # bases before the indel (the indel follows the anchor base vpos)
        if isdel.any():
            d = ilen[v]
            m = isdel
            refidx[m] = np.where(offs < k[m, None], refidx[m], refidx[m] + d[m, None])
            cig[m, 0] = (k[m] << 4) | 0
            cig[m, 1] = (d[m] << 4) | 2
            cig[m, 2] = ((READLEN - k[m]) << 4) | 0
        if isins.any():
            d = ilen[v]
            m = isins
            after = offs >= (k[m, None] + d[m, None])
            refidx[m] = np.where(offs < k[m, None], refidx[m], np.where(after, refidx[m] - d[m, None], refidx[m]))
            cig[m, 0] = (k[m] << 4) | 0
            cig[m, 1] = (d[m] << 4) | 1
            cig[m, 2] = ((READLEN - k[m] - d[m]) << 4) | 0
        seq = ref_codes[np.minimum(refidx, length - 1)]
        if isins.any():
            m = np.nonzero(isins)[0]
            ins = (offs >= k[m, None]) & (offs < k[m, None] + ilen[v[m], None])
            # inserted bases: a fixed pattern per variant so that every carrier shows the same allele
            pat = (alt[v[m], None] + offs - k[m, None]) % 4
            seq[m] = np.where(ins, pat, seq[m])
        snv = np.nonzero(carries & (kind[v] == 0))[0]
        seq[snv, vpos[v[snv]] - start[snv]] = alt[v[snv]]
        qual = qv[rng.choice(qv.shape[0], size=(n, READLEN), p=qp)]
        err = rng.random((n, READLEN)) < 10.0 ** (-qual.astype(np.float64) / 10.0)
        seq = np.where(err, (seq + rng.integers(1, 4, size=(n, READLEN))) % 4, seq).astype(np.uint8)
        mapq = np.full(n, 60, dtype=np.int64)
        pos = start.copy()
        if lowmapq_frac > 0:
            m = rng.random(n) < lowmapq_frac
            mapq[m] = rng.choice([0, 5, 15, 25, 37], size=int(m.sum()))
        if noisy_frac > 0:
            for r in np.nonzero(rng.random(n) < noisy_frac)[0]:
                at = rng.choice(READLEN, size=int(rng.integers(3, 9)), replace=False)
                seq[r, at] = (seq[r, at] + rng.integers(1, 4, size=at.shape[0])) % 4
        if clip_frac > 0:
            plain = (cig[:, 1] == 0) & (mpos < 0)
            for r in np.nonzero(plain & (rng.random(n) < clip_frac))[0]:
                lead = int(rng.integers(0, 11)) if rng.random() < 0.7 else 0
                trail = int(rng.integers(1, 11)) if (lead == 0 or rng.random() < 0.5) else 0
                ops = ([(lead << 4) | 4] if lead else []) + [((READLEN - lead - trail) << 4) | 0] + ([(trail << 4) | 4] if trail else [])
                cig[r, :] = 0
                cig[r, :len(ops)] = ops
                pos[r] = start[r] + lead
                seq[r, :lead] = rng.integers(0, 4, size=lead)
                if trail:
                    seq[r, READLEN - trail:] = rng.integers(0, 4, size=trail)
        if n_frac > 0:
            seq = np.where(rng.random((n, READLEN)) < n_frac, 4, seq).astype(np.uint8)
        order = np.argsort(pos, kind="stable")
        names = np.zeros((n, 10), dtype=np.uint8)
        digits = fragment.copy()
        for c in range(8, -1, -1):
            names[:, c] = 48 + digits % 10
            digits //= 10
        names[:, 0] = ord("r")
        recs = _pack_records(0, pos[order].astype(np.int32), flag[order], mapq[order], cig[order], seq[order], qual[order], names[order],
                             mpos[order], isize[order])
        path = os.path.join(outdir, f"pool_{p:03d}.bam")
        with open(path, "wb") as f:
            _bgzf_write(f, bam_header + recs)
            f.write(_BGZF_EOF)
        bam_paths.append(path)
    with open(os.path.join(outdir, "bams.list"), "w") as f:
        for path in bam_paths:
            f.write(path + "\n")
    return dict(fasta=fasta, bams=bam_paths, bamlist=os.path.join(outdir, "bams.list"), vpos=vpos, af=af, alt=alt, kind=kind,
                ilen=ilen, acount=acount, length=length, chrom=chrom, pools=pools, poolsize=poolsize)
