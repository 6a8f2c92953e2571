"""Pileup on the device: BAM decode on host threads, allele counting / candidate screen / site
packing on the GPU.

Pins (CPU, `-m "not gpu"`):
  * host/bam_decode.c against an independent Python parse of the same BAM files;
  * oracle/pileup_oracle.py (numpy restatement of calculate_allelecounts & co.) against the REFERENCE's own front end —
    the committed fixture tests/golden/pileup/frontend.npz (made from a CRISP.gpu dump by scripts/make_pileup_golden.py)
    and, where oracle/_ref is built, a live dump of a second, larger data set.
GPU (`-m gpu`): crispgpu_pileup_run against the fixture, against the restatement on adversarial alignment sets (clips,
N and IUPAC bases, lower-case reference, reads across the window edges, empty pools, long reads, diploid pools), its error
behaviour, window sharding, and BAM -> VCF: decode + device pileup + engine reproduce CRISP.binary's records.
"""
import gzip
import os
import re
import struct

import numpy as np
import pytest

from crisp_b200 import frontend as fe
from crisp_b200.abi import ALTREAD_DTYPE, EngineConfig
from crisp_b200.bamsynth import make_dataset
from crisp_b200.pileup import ALNREC_DTYPE, Alignments, decode_bams, make_window, pack_alignments, read_fasta
from frontend_io import read_dump
from oracle.pileup_oracle import pileup_oracle
from util import compare_results

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
GOLDEN = os.path.join(ROOT, "tests", "golden", "pileup", "frontend.npz")
REFBIN = os.path.join(ROOT, "oracle", "_ref", "CRISP.binary")
GPUBIN = os.path.join(ROOT, "oracle", "_ref", "CRISP.gpu")

SITE_FIELDS = ("pos", "refbase", "maxvec_al", "maxvec_ct", "counts", "indcounts", "readdepths", "mqcounts", "filtered", "read_off", "reads",
               "alt_off", "alt_reads")
HARD = dict(lowmapq_frac=0.15, noisy_frac=0.1, clip_frac=0.2, n_frac=0.01, qual_values=[5, 10, 12, 13, 20, 30, 40],
            qual_probs=[.05, .05, .05, .25, .2, .2, .2])


def golden():
    z = np.load(GOLDEN)
    aln = Alignments(int(z["n_pools"]), z["pool_off"], z["rec"], z["seq4"], z["qual"])
    want = {k: z["want_" + k] for k in SITE_FIELDS}
    return aln, z["ref"], int(z["length"]), int(z["poolsize"]), want


def assert_sites_equal(got, want, label=""):
    """got: dict or PileupSites; want: dict.  Every array bit for bit."""
    for k in SITE_FIELDS:
        g = np.asarray(got[k] if isinstance(got, dict) else getattr(got, k)).reshape(-1)
        w = np.asarray(want[k]).reshape(-1)
        assert g.shape == w.shape, f"{label}: {k} has {g.shape[0]} elements, expected {w.shape[0]}"
        if g.dtype.fields:
            same = g.tobytes() == w.astype(g.dtype).tobytes()
            assert same, f"{label}: {k} differs"
        else:
            bad = np.nonzero(g.astype(np.int64) != w.astype(np.int64))[0]
            assert bad.size == 0, f"{label}: {k} differs at {bad.size} places, first {bad[:5].tolist()}: {g[bad[:5]].tolist()} vs {w[bad[:5]].tolist()}"


def python_bam_records(path):
    raw = gzip.open(path).read()
    (lt,) = struct.unpack_from("<i", raw, 4)
    o = 8 + lt
    (nref,) = struct.unpack_from("<i", raw, o)
    o += 4
    for _ in range(nref):
        (ln,) = struct.unpack_from("<i", raw, o)
        o += 4 + ln + 4
    out = []
    while o < len(raw):
        bs, tid, pos, lname, mapq, _bin, ncig, flag, lseq, _mtid, mpos, tlen = struct.unpack_from("<iiiBBHHHiiii", raw, o)
        c0 = o + 36 + lname
        cig = struct.unpack_from(f"<{ncig}I", raw, c0)
        s0 = c0 + 4 * ncig
        seq = np.frombuffer(raw, np.uint8, (lseq + 1) // 2, s0)
        qual = np.frombuffer(raw, np.uint8, lseq, s0 + (lseq + 1) // 2)
        out.append(dict(tid=tid, pos=pos, mapq=mapq, flag=flag, cig=cig, seq=seq, qual=qual, mpos=mpos, tlen=tlen, lseq=lseq))
        o += 4 + bs
    return out


# ---------------------------------------------------------------- CPU: decoder and restatement ----------------------------------------------------------------

def test_decoder_matches_python_parse(tmp_path):
    ds = make_dataset(str(tmp_path), pools=3, poolsize=8, length=5000, depth=20, seed=5, indel_frac=0.2, **HARD)
    aln = decode_bams(ds["bams"], pinned=False, threads=3, allow_complex=True, paired=True)
    assert aln.stats["n_refs"] == 1 and aln.stats["first_ref_name"] == "chr1" and aln.stats["first_ref_len"] == 5000
    n_complex = 0
    for j, path in enumerate(ds["bams"]):
        recs = python_bam_records(path)
        simple = []
        for r in recs:
            ops = [c & 15 for c in r["cig"]]
            if any(op in (1, 2, 3, 6) for op in ops):
                n_complex += 1
                continue
            simple.append(r)
        got = aln.rec[aln.pool_off[j]:aln.pool_off[j + 1]]
        assert got.shape[0] == len(simple)
        for g, r, mp, isz in zip(got, simple, aln.mate_pos[aln.pool_off[j]:], aln.isize[aln.pool_off[j]:]):
            clip = (r["cig"][0] >> 4) if (r["cig"][0] & 15) == 4 else 0
            mlen = sum(c >> 4 for c in r["cig"] if (c & 15) == 0)
            assert (g["pos"], g["clip"], g["mlen"], g["mapq"], g["strand"]) == (r["pos"], clip, mlen, r["mapq"], 1 if r["flag"] & 16 else 0)
            assert g["base_off"] % 8 == 0 and (mp, isz) == (r["mpos"], r["tlen"])
            o = int(g["base_off"])
            assert np.array_equal(aln.qual[o:o + r["lseq"]], r["qual"]) and np.array_equal(aln.seq4[o // 2:o // 2 + (r["lseq"] + 1) // 2], r["seq"])
    assert n_complex > 0 and aln.stats["n_complex"] == n_complex
    with pytest.raises(RuntimeError, match="outside the device pileup's scope"):
        decode_bams(ds["bams"], pinned=False)


def test_decoder_region_and_flag_filter(tmp_path):
    ds = make_dataset(str(tmp_path), pools=2, poolsize=8, length=6000, depth=15, seed=6)
    whole = decode_bams(ds["bams"], pinned=False, threads=2)
    part = decode_bams(ds["bams"], pinned=False, threads=1, tid=0, start=2000, end=2500)
    for j in range(2):
        w = whole.rec[whole.pool_off[j]:whole.pool_off[j + 1]]
        keep = (w["pos"] + w["mlen"].astype(np.int64) > 2000) & (w["pos"] < 2500)
        p = part.rec[part.pool_off[j]:part.pool_off[j + 1]]
        assert np.array_equal(p["pos"], w["pos"][keep]) and np.array_equal(p["strand"], w["strand"][keep])
    assert part.stats["n_outside"] == whole.n_reads - part.n_reads
    rev_dropped = decode_bams(ds["bams"], pinned=False, filter_mask=16)   # any flag bit can be masked: drop reverse-strand reads
    assert rev_dropped.n_reads == int((whole.rec["strand"] == 0).sum()) and rev_dropped.stats["n_flag_filtered"] == int((whole.rec["strand"] == 1).sum())
    assert decode_bams(ds["bams"], pinned=False, tid=3).n_reads == 0


def _same_alignments(a, b):
    assert np.array_equal(a.pool_off, b.pool_off)
    for k in ("pos", "clip", "mlen", "mapq", "strand"):
        assert np.array_equal(a.rec[k], b.rec[k]), k
    for j in range(a.n_reads):
        oa, ob, n = int(a.rec["base_off"][j]), int(b.rec["base_off"][j]), int(a.rec["clip"][j]) + int(a.rec["mlen"][j])
        assert np.array_equal(a.qual[oa:oa + n], b.qual[ob:ob + n]) and np.array_equal(a.seq4[oa // 2:(oa + n + 1) // 2], b.seq4[ob // 2:(ob + n + 1) // 2])


def test_decoder_reaches_a_region_through_the_bai_index(tmp_path):
    """With a BAI index next to the files (built by the reference's own samtools, oracle/_ref/bamtool) a region decode inflates
    only the blocks the index names and returns exactly what the whole-file decode filtered to the region returns."""
    from crisp_b200 import frontend as fe
    bamtool = os.path.join(ROOT, "oracle", "_ref", "bamtool")
    if not os.path.exists(bamtool):
        pytest.skip("oracle/_ref/bamtool not built (needs /root/reference)")
    ds = make_dataset(str(tmp_path), pools=3, poolsize=8, length=120000, depth=12, seed=16)
    whole = decode_bams(ds["bams"], pinned=False, threads=2)
    assert whole.stats["inflated_bytes"] > 40 * 65280      # dozens of BGZF blocks per pool
    fe.index_bams(bamtool, ds)
    assert all(os.path.exists(b + ".bai") for b in ds["bams"])
    rng = np.random.default_rng(4)
    regions = [(0, 500), (0, 120000), (119000, 2**31 - 1), (60000, 60001), (16383, 16385), (32768, 49152)] + \
              [tuple(sorted(int(x) for x in rng.integers(0, 120000, 2))) for _ in range(12)]
    for start, end in regions:
        if end == start:
            end += 1
        plain = decode_bams(ds["bams"], pinned=False, threads=2, tid=0, start=start, end=end, use_index=False)
        fast = decode_bams(ds["bams"], pinned=False, threads=2, tid=0, start=start, end=end)
        _same_alignments(plain, fast)
        assert fast.stats["inflated_bytes"] <= plain.stats["inflated_bytes"]
        if end - start < 20000:
            assert fast.stats["inflated_bytes"] < plain.stats["inflated_bytes"] // 2, (start, end, fast.stats["inflated_bytes"])
        assert fast.stats["first_ref_name"] == plain.stats["first_ref_name"] and fast.stats["n_refs"] == plain.stats["n_refs"]
    # a region on a reference the files do not have, and one past the end of the data
    assert decode_bams(ds["bams"], pinned=False, tid=5, start=10, end=20).n_reads == 0
    assert decode_bams(ds["bams"], pinned=False, tid=0, start=10**8, end=10**8 + 5).n_reads == 0
    # the index of another file, a truncated one and garbage are ignored (whole-file decode), not trusted
    other = make_dataset(str(tmp_path / "other"), pools=1, poolsize=8, length=30000, depth=30, seed=17)
    fe.index_bams(bamtool, other)
    want = decode_bams(ds["bams"], pinned=False, tid=0, start=70000, end=71000, use_index=False)
    for bad in (open(other["bams"][0] + ".bai", "rb").read(), open(ds["bams"][0] + ".bai", "rb").read()[:40], b"BAI\x01" + bytes(range(200)), b"hello"):
        for b in ds["bams"]:
            open(b + ".bai", "wb").write(bad)
        _same_alignments(want, decode_bams(ds["bams"], pinned=False, tid=0, start=70000, end=71000))
    # x.bai beside x.bam is found as well
    for b in ds["bams"]:
        os.remove(b + ".bai")
    fe.index_bams(bamtool, ds)
    for b in ds["bams"]:
        os.rename(b + ".bai", b[:-4] + ".bai")
    got = decode_bams(ds["bams"], pinned=False, tid=0, start=70000, end=71000)
    _same_alignments(want, got)
    assert got.stats["inflated_bytes"] < want.stats["inflated_bytes"] // 2


def test_decoder_rejects_corrupt_files(tmp_path):
    ds = make_dataset(str(tmp_path), pools=1, poolsize=4, length=3000, depth=5, seed=2)
    raw = bytearray(open(ds["bams"][0], "rb").read())
    bad = str(tmp_path / "bad.bam")
    raw[len(raw) // 2] ^= 0x55   # inside a deflate stream: inflate or the CRC fails
    open(bad, "wb").write(raw)
    with pytest.raises(RuntimeError, match="status -3"):
        decode_bams([bad], pinned=False)
    open(bad, "wb").write(b"not a bam file at all, but long enough to hold a header")
    with pytest.raises(RuntimeError, match="status -3"):
        decode_bams([bad], pinned=False)
    with pytest.raises(RuntimeError, match="status -2"):
        decode_bams([str(tmp_path / "missing.bam")], pinned=False)


def test_restatement_matches_reference_fixture():
    aln, ref, length, poolsize, want = golden()
    got = pileup_oracle(aln, ref, 0, 0, length, np.full(aln.n_pools, poolsize))
    assert got["n_sites"] == want["pos"].shape[0] >= 30
    assert_sites_equal(got, want, "restatement vs reference front end (fixture)")


def _live_dump(tmp_path, poolsize, **kw):
    if not os.path.exists(GPUBIN):
        pytest.skip("oracle/_ref/CRISP.gpu not built (needs /root/reference)")
    ds = make_dataset(str(tmp_path), poolsize=poolsize, **kw)
    dump = str(tmp_path / "dump.bin")
    fe.run_program(GPUBIN, ds, str(tmp_path / "unused.vcf"), poolsize, [], env=dict(CRISPGPU_DUMP=dump, CRISPGPU_DUMP_ONLY=1, CRISPGPU_BATCH=10 ** 7))
    (cfg, batch, ex), = read_dump(dump)
    want = {k: np.asarray(getattr(batch, k)) for k in SITE_FIELDS if k not in ("readdepths", "mqcounts", "filtered")}
    want.update({k: ex[k] for k in ("readdepths", "mqcounts", "filtered")})
    return ds, cfg, batch, want


@pytest.mark.parametrize("poolsize,kw", [(20, dict(pools=6, length=12000, depth=60, seed=21, **HARD)),
                                         (2, dict(pools=12, length=8000, depth=12, seed=22, **HARD))], ids=["pooled", "diploid"])
def test_restatement_matches_reference_frontend_live(tmp_path, poolsize, kw):
    ds, cfg, batch, want = _live_dump(tmp_path, poolsize, **kw)
    aln = decode_bams(ds["bams"], pinned=False, threads=2)
    got = pileup_oracle(aln, read_fasta(ds["fasta"]), 0, 0, ds["length"], np.full(kw["pools"], poolsize))
    assert got["n_sites"] > 50
    assert_sites_equal(got, want, "restatement vs reference front end (live)")


# ---------------------------------------------------------------- GPU ----------------------------------------------------------------

def _engine(P, poolsize, **opts):
    from crisp_b200.engine import Engine
    return Engine(EngineConfig(ploidy=np.full(P, poolsize), options=opts), device=0)


@pytest.mark.gpu
def test_device_pileup_matches_reference_fixture():
    from oracle.loader import OracleEngine
    aln, ref, length, poolsize, want = golden()
    eng = _engine(aln.n_pools, poolsize)
    sites, res = eng.pileup_run(aln, make_window(0, 0, length, ref, want_streams=True))
    t = eng.timing()
    assert_sites_equal(sites, want, "device pileup vs reference front end (fixture)")
    assert t["n_positions"] == length and t["n_alignments"] == aln.n_reads and t["pileup_ms"] > 0 and t["pile_tile_ms"] > 0
    # the engine ran on the packed sites in device memory: same results as the checker on the reference's batch
    batch = sites.as_batch(aln.n_pools)
    compare_results(res, OracleEngine(eng.cfg).run(batch), label="pileup results")
    # and as the engine itself on the same batch shipped from the host
    compare_results(res, eng.run(batch), label="pileup vs host batch", tol=0.0)
    eng.close()


def _random_alignments(rng, P, L, depth, ploidy, readlen=(30, 150), long_frac=0.0, long_len=(200, 420), odd_ref=False):
    ref_codes = rng.integers(0, 4, size=L)
    ref = np.frombuffer(b"ACGT", np.uint8)[ref_codes].copy()
    lower = rng.random(L) < 0.1
    ref[lower] += 32
    ref[rng.random(L) < 0.01] = ord("N")
    ref[rng.random(L) < 0.003] = ord("R")
    if odd_ref:   # characters outside BAM's base alphabet: no read base equals them
        ref[rng.random(L) < 0.004] = ord("X")
        ref[rng.random(L) < 0.002] = ord("*")
    alphabet = "ACGT"
    pools = []
    for j in range(P):
        if j == 1 and P > 2:
            pools.append([])   # an empty pool
            continue
        n = int(L * depth / 80)
        starts = np.sort(rng.integers(-40, L - 10, size=n))
        reads = []
        for s in starts:
            ln = int(rng.integers(readlen[0], readlen[1] + 1)) if rng.random() >= long_frac else int(rng.integers(long_len[0], long_len[1]))
            lead = int(rng.integers(0, 9)) if rng.random() < 0.2 and ln > 20 else 0
            trail = int(rng.integers(0, 9)) if rng.random() < 0.2 and ln > 20 else 0
            ml = ln - lead - trail
            pos = int(s) + 40   # keep positions >= 0
            idx = np.clip(np.arange(pos, pos + ml), 0, L - 1)
            body = [alphabet[c] for c in ref_codes[idx]]
            errs = rng.random(ml) < (0.2 if rng.random() < 0.1 else 0.02)
            for k in np.nonzero(errs)[0]:
                body[k] = alphabet[(ref_codes[idx[k]] + int(rng.integers(1, 4))) % 4]
            for k in np.nonzero(rng.random(ml) < 0.01)[0]:
                body[k] = "NRYKMSW="[int(rng.integers(0, 8))]
            seq = "".join(alphabet[int(c)] for c in rng.integers(0, 4, size=lead)) + "".join(body) + "".join(alphabet[int(c)] for c in rng.integers(0, 4, size=trail))
            qual = rng.choice([2, 12, 13, 14, 25, 37, 41, 60], size=ln)
            reads.append(dict(pos=pos, seq=seq, qual=qual, clip=lead, mlen=ml, mapq=int(rng.choice([0, 9, 10, 19, 20, 39, 40, 60])), strand=int(rng.integers(0, 2))))
        pools.append(reads)
    # a planted variant every 40 bases in a few pools so that candidates exist at every depth
    for j, reads in enumerate(pools):
        for r in reads:
            for v in range(20, L + 40, 40):
                k = v - r["pos"]
                if 0 <= k < r["mlen"] and (v // 40 + j) % 3 == 0 and rng.random() < 0.5:
                    s = list(r["seq"])
                    s[r["clip"] + k] = alphabet[(ref_codes[min(v, L - 1)] + 1) % 4]
                    r["seq"] = "".join(s)
    return pack_alignments(pools), ref


@pytest.mark.gpu
@pytest.mark.parametrize("seed,P,L,depth,ploidy,kw", [
    (1, 5, 900, 25, 20, {}),
    (2, 3, 2500, 40, 8, dict(long_frac=0.1)),          # reads longer than one 128-base warp pass, tiles of 1024 positions crossed
    (3, 40, 300, 8, 2, {}),                            # diploid pools: the potential-variant score counts single reads
    (4, 1, 1500, 60, 30, {}),
    (5, 70, 200, 40, 4, dict(readlen=(10, 40))),       # more than two lanes-rounds of pools, short reads
    (6, 4, 3000, 30, 12, dict(long_frac=0.05, long_len=(600, 1500))),   # match blocks beyond the staged reference stretch: the general kernel path
    (7, 4, 2200, 30, 12, dict(odd_ref=True)),           # reference characters outside BAM's alphabet: staged, but not the 4-bit fast path
], ids=["pooled", "long_reads", "diploid", "one_pool", "many_pools", "very_long_reads", "odd_reference"])
def test_device_pileup_matches_restatement(seed, P, L, depth, ploidy, kw):
    rng = np.random.default_rng(seed)
    aln, ref = _random_alignments(rng, P, L, depth, ploidy, **kw)
    eng = _engine(P, ploidy)
    for (start, end, ref_start) in [(0, L, 0), (L // 3, 2 * L // 3, 0), (100, L - 50, 60)]:
        want = pileup_oracle(aln, ref[ref_start:], ref_start, start, end, np.full(P, ploidy))
        sites, res = eng.pileup_run(aln, make_window(0, start, end, ref[ref_start:], ref_start=ref_start, want_streams=True))
        assert want["n_sites"] > 0
        assert_sites_equal(sites, want, f"device pileup vs restatement [{start},{end}) ref from {ref_start}")
        assert res.n_sites == want["n_sites"]
    eng.close()


@pytest.mark.gpu
def test_device_pileup_windows_concatenate():
    """Results do not depend on how the region is cut into windows (the RNG key is (tid, pos, slot))."""
    aln, ref, length, poolsize, want = golden()
    eng = _engine(aln.n_pools, poolsize, chisq_permutation=1, max_iter=2000)
    whole_sites, whole = eng.pileup_run(aln, make_window(0, 0, length, ref, want_streams=True))
    pos, ct, delta = [], [], []
    for a, b in [(0, 1024), (1024, 1025), (1025, 3333), (3333, length)]:
        s, r = eng.pileup_run(aln, make_window(0, a, b, ref, want_streams=True))
        pos.append(s.pos), ct.append(r.ctpval), delta.append(r.delta)
    assert np.array_equal(np.concatenate(pos), whole_sites.pos)
    assert np.array_equal(np.concatenate(ct), whole.ctpval) and np.array_equal(np.concatenate(delta), whole.delta)
    # the same windows over alignments that went to the device once (crispgpu_pileup_upload + _set_window)
    eng.pileup_upload(aln, make_window(0, 0, 1024, ref))
    pos2, ct2 = [], []
    for k, (a, b) in enumerate([(0, 1024), (1024, 1025), (1025, 3333), (3333, length)]):
        if k:
            eng.pileup_set_window(make_window(0, a, b, None))
        s, r = eng.pileup_run_resident(fetch=True)
        pos2.append(s.pos), ct2.append(r.ctpval)
    assert np.array_equal(np.concatenate(pos2), whole_sites.pos) and np.array_equal(np.concatenate(ct2), whole.ctpval)
    # an empty window and a window without alignments
    s, r = eng.pileup_run(aln, make_window(0, 500, 500, ref))
    assert s.n_sites == 0 and r.n_sites == 0
    empty = Alignments(aln.n_pools, np.zeros(aln.n_pools + 1, np.uint32), np.zeros(0, ALNREC_DTYPE), np.zeros(0, np.uint8), np.zeros(0, np.uint8))
    s, r = eng.pileup_run(empty, make_window(0, 0, 2000, ref))
    assert s.n_sites == 0 and r.n_sites == 0
    eng.close()


@pytest.mark.gpu
def test_device_pileup_refuses_what_it_cannot_pile():
    from crisp_b200.engine import EngineError
    aln, ref, length, poolsize, _ = golden()
    eng = _engine(aln.n_pools, poolsize)
    win = lambda: make_window(0, 0, length, ref)  # noqa: E731
    bad = Alignments(aln.n_pools, aln.pool_off, aln.rec.copy(), aln.seq4, aln.qual)
    bad.rec["pos"][5], bad.rec["pos"][6] = bad.rec["pos"][6] + 7, bad.rec["pos"][5]
    with pytest.raises(EngineError, match="coordinate order"):
        eng.pileup_run(bad, win())
    bad = Alignments(aln.n_pools, aln.pool_off, aln.rec.copy(), aln.seq4, aln.qual)
    bad.rec["base_off"][10] += 4
    with pytest.raises(EngineError, match="multiple of 8"):
        eng.pileup_run(bad, win())
    bad = Alignments(aln.n_pools, aln.pool_off, aln.rec.copy(), aln.seq4, aln.qual)
    bad.rec["base_off"][-1] = aln.n_bases
    with pytest.raises(EngineError, match="past n_bases"):
        eng.pileup_run(bad, win())
    # overlapping mates: the reference merges them (variant.c:137-220); the device pileup refuses
    mate = np.full(aln.n_reads, -1, np.int32)
    isz = np.zeros(aln.n_reads, np.int32)
    mate[3] = aln.rec["pos"][3] + 30
    isz[3] = 130
    paired = Alignments(aln.n_pools, aln.pool_off, aln.rec, aln.seq4, aln.qual, mate, isz)
    with pytest.raises(EngineError, match="overlapping mates"):
        eng.pileup_run(paired, win())
    mate[3] = aln.rec["pos"][3] + 400   # a proper pair that does not overlap is fine
    isz[3] = 500
    sites, _ = eng.pileup_run(Alignments(aln.n_pools, aln.pool_off, aln.rec, aln.seq4, aln.qual, mate, isz), win())
    assert sites.n_sites > 0
    wrong_pools = Alignments(aln.n_pools - 1, aln.pool_off[:-1], aln.rec, aln.seq4, aln.qual)
    with pytest.raises(EngineError):
        eng.pileup_run(wrong_pools, win())
    eng.close()


NOMAPQ = HARD   # (kept name: every filter the front end applies is on)


@pytest.mark.gpu
@pytest.mark.parametrize("poolsize,extra,opts,kw", [
    (20, [], {}, dict(pools=6, length=12000, depth=60, seed=41, **HARD)),
    (2, [], {}, dict(pools=12, length=8000, depth=12, seed=42, **HARD)),
    (16, ["--chisquare", "0", "--perms", "3000"], dict(chisq_permutation=1, max_iter=3000), dict(pools=8, length=9000, depth=50, seed=43, **HARD)),
    (20, ["--EM", "0"], dict(em_flag=0), dict(pools=6, length=16000, depth=60, seed=44, clip_frac=0.2, lowmapq_frac=0.1, n_frac=0.005)),
], ids=["pooled", "diploid", "perm", "em0"])
def test_bam_to_vcf_without_the_reference(tmp_path, poolsize, extra, opts, kw):
    """BAM files -> host/bam_decode.c -> crispgpu_pileup_run -> host/vcf_writer.c == CRISP.binary on the same files (the
    reference's whole program: its own BAM reader, pileup, engine and printer), record for record, byte for byte; several
    device windows.  With --chisquare 0 the CT= field is a Monte-Carlo value (the reference seeds drand48 with the time)."""
    from crisp_b200.caller import call_bams, sample_id
    if not os.path.exists(REFBIN):
        pytest.skip("oracle/_ref/CRISP.binary not built")
    ds = make_dataset(str(tmp_path), poolsize=poolsize, **kw)
    ref_vcf = str(tmp_path / "ref.vcf")
    fe.run_program(REFBIN, ds, ref_vcf, poolsize, extra)
    want_lines = open(ref_vcf, "rb").read().split(b"\n")
    want = [w for w in want_lines if w and not w.startswith(b"#")]
    assert len(want) >= 5
    out = call_bams(ds["bams"], ds["fasta"], poolsize, out_vcf=str(tmp_path / "b200.vcf"), options=opts, window=3000, threads=4)
    got_lines = open(str(tmp_path / "b200.vcf"), "rb").read().split(b"\n")
    got = [g for g in got_lines if g and not g.startswith(b"#")]
    assert out["records"] == len(got)
    # same column header (pool names derived from the BAM paths as the reference does)
    assert [x for x in got_lines if x.startswith(b"#CHROM")] == [x for x in want_lines if x.startswith(b"#CHROM")]
    assert sample_id("/a/b/pool_003.sorted.bam") == "pool_003"
    if not opts.get("chisq_permutation"):
        assert got == want, next((f"first difference:\n{g.decode()}\n{w.decode()}" for g, w in zip(got, want) if g != w), f"{len(got)} records vs {len(want)}")
        return
    key = lambda line: int(line.split(b"\t")[1])  # noqa: E731
    gmap, wmap = {key(g): g for g in got}, {key(w): w for w in want}
    # an allele whose Monte-Carlo p-value sits at a FAST_FILTER threshold may be kept in one run and dropped in the other
    # (the reference's own runs differ from each other the same way): a missing record, or a record with one allele less
    ct = re.compile(rb"CT=[^;]*")
    differing = [k for k in set(gmap) & set(wmap) if ct.sub(b"CT=", gmap[k]) != ct.sub(b"CT=", wmap[k])]
    for k in differing:
        assert gmap[k].split(b"\t")[4] != wmap[k].split(b"\t")[4], f"record differs beyond its allele list:\n{gmap[k].decode()}\n{wmap[k].decode()}"
    assert len(set(gmap) ^ set(wmap)) + len(differing) <= 3, (sorted(set(gmap) ^ set(wmap)), differing)


@pytest.mark.gpu
def test_device_pileup_refuses_weighted_counts():
    """--weighted sums 1 - error probability per read into Qhighcounts, which the device pileup does not accumulate."""
    from crisp_b200.engine import EngineError
    aln, ref, length, poolsize, _ = golden()
    eng = _engine(aln.n_pools, poolsize, use_qv=1)
    with pytest.raises(EngineError, match="--weighted"):
        eng.pileup_run(aln, make_window(0, 0, length, ref))
    eng.close()
