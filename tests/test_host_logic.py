"""Host-side logic: synthetic generator, allele sort, batch slicing, region sharding and the rank-0 merge
(world_size-2 gloo run)."""
import os
import subprocess
import sys

import numpy as np
import pytest

from crisp_b200.abi import EngineConfig, concat_batches
from crisp_b200.shard import merge_results, partition_sites, shard_batch
from crisp_b200.synth import sort_alleles, synth_batch
from oracle.loader import RNG_PHILOX, OracleEngine
from util import compare_results

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def test_synth_is_deterministic_and_candidate_only():
    a = synth_batch(200, 8, 20, 60, seed=3, variant_frac=0.4, bidir_frac=0.1)
    b = synth_batch(200, 8, 20, 60, seed=3, variant_frac=0.4, bidir_frac=0.1)
    assert a.n_sites == b.n_sites and np.array_equal(a.reads, b.reads) and np.array_equal(a.indcounts, b.indcounts)
    ic = a.indcounts.reshape(a.n_sites, 8, 3, 8).sum(axis=2)
    for s in range(a.n_sites):
        nonref = [x for x in range(8) if x != a.refbase[s]]
        assert (ic[s][:, nonref] >= 3).any()  # variantcalls.c:85,92
    # counts are consistent with the packed stream
    ro = a.read_off.astype(np.int64)
    for s in (0, a.n_sites - 1):
        for i in range(8):
            w = a.reads[ro[s * 8 + i]:ro[s * 8 + i + 1]].astype(np.int64)
            cls = np.where((w >> 11) & 1, 2, (w >> 10) & 1)
            key = ((w >> 8) & 3) + 8 * cls
            assert np.array_equal(np.bincount(key, minlength=24), a.indcounts.reshape(a.n_sites, 8, 24)[s, i])


def test_sort_alleles_matches_reference_rule():
    # reference first, then exchange sort by count descending (allelecounts.c:155-187)
    counts = np.array([[5, 50, 7, 7, 0, 0, 3, 0], [9, 1, 9, 30, 2, 0, 0, 0]])
    al, ct = sort_alleles(counts, np.array([1, 3]))
    assert al[0].tolist()[:5] == [1, 2, 3, 0, 6] and ct[0].tolist()[:5] == [50, 7, 7, 5, 3]
    assert al[1].tolist()[:4] == [3, 0, 2, 4] or al[1].tolist()[:4] == [3, 2, 0, 4]
    assert ct[1].tolist()[:5] == [30, 9, 9, 2, 1]


def test_select_concat_roundtrip():
    b = synth_batch(80, 6, 16, 50, seed=9, variant_frac=0.5, indel_frac=0.4)
    idx = np.arange(b.n_sites)
    parts = [b.select(idx[:30]), b.select(idx[30:])]
    c = concat_batches(parts)
    for name in ("tid", "pos", "refbase", "maxvec_al", "maxvec_ct", "counts", "indcounts", "read_off", "reads", "alt_off"):
        assert np.array_equal(getattr(b, name), getattr(c, name)), name
    assert np.array_equal(b.alt_reads, c.alt_reads)


def test_partition_is_contiguous_and_balanced():
    rng = np.random.default_rng(0)
    tid = rng.integers(0, 3, 1000)
    pos = rng.integers(0, 10**6, 1000)
    parts = partition_sites(tid, pos, 8)
    allidx = np.concatenate(parts)
    assert sorted(allidx.tolist()) == list(range(1000))
    keys = tid[allidx].astype(np.int64) * 10**7 + pos[allidx]
    assert np.all(np.diff(keys) >= 0)
    assert max(len(p) for p in parts) - min(len(p) for p in parts) <= 1
    w = rng.random(1000)
    wp = partition_sites(tid, pos, 4, weight=w)
    loads = [w[p].sum() for p in wp]
    assert max(loads) - min(loads) < 0.05 * w.sum()


def test_sharded_run_equals_single_run():
    """Results do not depend on how sites are sharded (Philox streams are keyed by site, not by position in the batch)."""
    cfg = EngineConfig(ploidy=np.full(8, 20), options=dict(chisq_permutation=1, max_iter=300))
    b = synth_batch(90, 8, 20, 60, seed=4, variant_frac=0.5)
    orc = OracleEngine(cfg)
    whole = orc.run(b, mode=RNG_PHILOX)
    parts, idxs = [], []
    for r in range(3):
        sub, idx = shard_batch(b, 3, r)
        parts.append(orc.run(sub, mode=RNG_PHILOX))
        idxs.append(idx)
    merged = merge_results(parts, idxs, b.n_sites, 8)
    compare_results(merged, whole, tol=0.0, label="sharded")
    assert np.array_equal(merged.perm_k, whole.perm_k) and np.array_equal(merged.perm_hits, whole.perm_hits)


def test_two_rank_gloo_gather():
    """world_size-2 run of the N>1 host path on CPU (gloo): each rank processes its region with the checker standing
    in for the device, rank 0 merges in coordinate order."""
    script = os.path.join(ROOT, "tests", "_gloo_worker.py")
    env = dict(os.environ, PYTHONPATH=ROOT + os.pathsep + os.path.join(ROOT, "tests"))
    res = subprocess.run(
        [sys.executable, "-m", "torch.distributed.run", "--nnodes=1", "--nproc-per-node=2", "--master-addr", "127.0.0.1",
         "--master-port", "29611", script], env=env, capture_output=True, text=True, timeout=600)
    assert res.returncode == 0, res.stdout[-2000:] + res.stderr[-4000:]
    assert "GLOO_MERGE_OK" in res.stdout


def test_bench_reference_arm_prints_one_json_line():
    """bench.py --impl reference (the reference's own CPU code on the host cores) prints exactly one JSON line with the
    keys the driver reads; runs without a GPU."""
    import json
    res = subprocess.run([sys.executable, os.path.join(ROOT, "bench.py"), "--impl", "reference", "--steps", "1", "--warmup", "0",
                          "--cpu-sites-per-core", "4"], capture_output=True, text=True, timeout=600)
    assert res.returncode == 0, res.stderr[-2000:]
    lines = [l for l in res.stdout.splitlines() if l.strip()]
    assert len(lines) == 1
    d = json.loads(lines[0])
    assert d["impl"] == "reference" and d["unit"] == "sites/s" and d["higher_is_better"] is True and d["value"] > 0
    assert d["cpu_baseline"]["kind"] in ("reference", "port") and d["cpu_baseline"]["cores"] >= 1
    assert d["e2e"]["h2d_bytes_per_step"] == 0 and "workload" in d["config"]


def test_compact_formats_round_trip():
    """bins_from_reads / sparse_counts_from_dense are exact re-encodings (what k_em reads directly and k_expand_counts undoes)."""
    from crisp_b200.abi import bins_from_reads, reads_from_bins, sparse_counts_from_dense
    batch = synth_batch(60, 6, np.full(6, 20), 80, seed=5, variant_frac=0.4, bidir_frac=0.1, indel_frac=0.3)
    bo, bins = bins_from_reads(batch.read_off, batch.reads)
    back = reads_from_bins(bo, bins)
    ro = batch.read_off.astype(np.int64)
    assert back.size == batch.reads.size
    for c in range(ro.size - 1):
        assert np.array_equal(np.sort(batch.reads[ro[c]:ro[c + 1]]), back[ro[c]:ro[c + 1]])
        cell = bins[bo[c]:bo[c + 1]] & 0xFFFF
        assert np.all(cell[1:] > cell[:-1])            # ascending, distinct values
    off, ic = sparse_counts_from_dense(batch.indcounts)
    dense = np.zeros_like(batch.indcounts).reshape(-1, 24)
    cell = np.repeat(np.arange(off.size - 1), np.diff(off.astype(np.int64)))
    dense[cell, ic >> 27] = ic & 0x7FFFFFF
    assert np.array_equal(dense.reshape(-1), batch.indcounts)
    c = batch.compact()
    c.validate()
    assert c.reads is None and c.read_off is None and c.indcounts is None and c.nbytes() < batch.nbytes()
    # a multiplicity above 65535 is split over consecutive bins of the same value
    ro1 = np.array([0, 140000], np.uint32)
    bo1, b1 = bins_from_reads(ro1, np.full(140000, 0x1A1E, np.uint16))
    assert b1.tolist() == [(65535 << 16) | 0x1A1E, (65535 << 16) | 0x1A1E, (8930 << 16) | 0x1A1E] and bo1.tolist() == [0, 3]


def test_sharding_a_compact_batch():
    """Region shards of a compact batch are the compact forms of the dense batch's shards, and cover every site once."""
    from crisp_b200.shard import shard_batch
    batch = synth_batch(90, 6, np.full(6, 20), 80, seed=8, variant_frac=0.4, indel_frac=0.3)
    compact = batch.compact()
    seen = []
    for r in range(3):
        sub, idx = shard_batch(compact, 3, r)
        seen.append(idx)
        ref = batch.select(idx).compact()
        assert sub.reads is None and sub.indcounts is None
        for name in ("tid", "pos", "counts", "bin_off", "bins", "ic_off", "ic_sparse", "alt_off"):
            assert np.array_equal(getattr(sub, name), getattr(ref, name)), name
    assert np.array_equal(np.sort(np.concatenate(seen)), np.arange(batch.n_sites))


def test_results_pack_unpack_round_trip():
    """What a rank sends to rank 0 (shard.pack_results) comes back as the same results (flat byte buffer, no pickling)."""
    from crisp_b200.abi import _RESULT_FIELDS, Results
    from crisp_b200.shard import merge_results, pack_results, unpack_results
    rng = np.random.default_rng(4)
    n, P, calls = 37, 6, 11
    r = Results(n, P, max_calls=calls)
    r.n_calls = calls
    for name, dt, w in _RESULT_FIELDS:
        a = getattr(r, name)
        a[...] = (rng.random(a.shape) * 50).astype(dt)
    r.ac[...] = rng.integers(0, 20, size=r.ac.shape)
    r.qv[...] = rng.random(r.qv.shape)
    idx = rng.permutation(100)[:n]
    buf = pack_results(r, idx)
    assert buf.dtype == np.uint8
    back, idx2 = unpack_results(buf)
    assert np.array_equal(idx2, idx) and back.n_calls == calls and back.n_sites == n
    for name, _, _ in _RESULT_FIELDS:
        assert np.array_equal(getattr(back, name), getattr(r, name)), name
    assert np.array_equal(back.ac, r.ac) and np.array_equal(back.qv, r.qv)
    # an empty rank (no sites) packs and merges
    e = Results(0, P, max_calls=0)
    eb, ei = unpack_results(pack_results(e, np.zeros(0, np.int64)))
    assert eb.n_sites == 0 and ei.size == 0
    merged = merge_results([back, eb], [np.arange(n), ei], n, P)
    assert np.array_equal(merged.delta, r.delta) and merged.n_calls == calls


def test_em_trace_report_reads_the_committed_timeline():
    """scripts/em_trace.py on the task timeline committed under profiles/r2 (CRISPGPU_EM_TRACE of one k_em launch at the
    benchmark shape): the parser, the phase table and the schedule simulation run, and the figures DESIGN.md quotes hold."""
    root = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
    out = subprocess.run([sys.executable, os.path.join(root, "scripts", "em_trace.py"), os.path.join(root, "profiles", "r2", "em_trace_final.bin")],
                         capture_output=True, text=True, check=True).stdout
    assert "tasks 5271" in out and "planes (DMMA)" in out and "makespan" in out
    busy = float(out.split("busy task-time / (span x concurrent) = ")[1].split(";")[0])
    assert 0.85 < busy < 0.95
