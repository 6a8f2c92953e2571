"""Narrow transport form of the compact batch (crispgpu_batch.bins16 / ic16): the host packer against its numpy statement
(CPU), and the engine's results on it against the same batch shipped in the compact and in the per-read form (GPU)."""
import numpy as np
import pytest

from crisp_b200.abi import Batch, EngineConfig, narrow_arrays
from crisp_b200.engine import EngineError, NarrowBatch
from crisp_b200.synth import synth_batch
from util import compare_results

FIELDS = (("bins16", np.uint16), ("bin_len", np.uint8), ("bin_site", np.uint32), ("ic16", np.uint16), ("ic_len", np.uint8), ("ic_site", np.uint32))


def check_pack(batch):
    nb = NarrowBatch(batch, pinned=False)
    want = narrow_arrays(batch.compact())
    for name, dt in FIELDS:
        got = nb.array(name, dt, want[name].size)
        assert np.array_equal(got, want[name]), name
    # the site totals do not travel (recomputed on the device); alt-read records as 4-byte words
    c = nb.as_c()
    assert not c.counts and c.maxvec_al and c.maxvec_ct
    ar = batch.alt_reads
    if ar is not None and ar.size and int(ar["offset"].max()) < 1024 and int(ar["aligned"].max()) < 1024:
        a32 = nb.array("alt32", np.uint32, ar.size)
        assert not c.alt_reads
        assert np.array_equal(a32, ar["pool"].astype(np.uint32) | (ar["offset"].astype(np.uint32) << 11) | (ar["aligned"].astype(np.uint32) << 21) | (ar["strand"].astype(np.uint32) << 31))
    return nb, want


@pytest.mark.parametrize("kw", [dict(n=40, P=50, ploidy=48, depth=200, variant_frac=0.3), dict(n=30, P=7, ploidy=20, depth=3000, variant_frac=0.5, bidir_frac=0.2),
                                dict(n=25, P=33, ploidy=2, depth=8, variant_frac=0.5), dict(n=20, P=5, ploidy=10, depth=60, quals=range(14, 42))],
                         ids=["c2", "deep_bidir", "diploid_shallow", "manyquals"])
def test_narrow_packer_matches_numpy_statement(kw):
    kw = dict(kw)
    b = synth_batch(kw.pop("n"), kw.pop("P"), kw.pop("ploidy"), kw.pop("depth"), seed=5, **kw)
    nb, want = check_pack(b)
    c = b.compact()
    # the narrow entries say what the compact ones say: per cell, the reads per value and the counts per index add up
    cnt = (want["bins16"] >> 11).astype(np.int64) + 1
    assert cnt.sum() == int((c.bins >> 16).sum(dtype=np.int64)) and int((want["ic16"] & 0x7FF).sum(dtype=np.int64)) == int((c.ic_sparse & 0x7FFFFFF).sum(dtype=np.int64))
    assert want["bin_site"][-1] == want["bins16"].size == want["bin_len"].sum(dtype=np.int64)
    if kw.get("bidir_frac") is None and not kw.get("quals"):   # (bins of thousands of reads take many 32-read entries: the narrow form pays at ordinary depth)
        assert nb.nbytes() < c.nbytes()


def test_narrow_packer_splits_large_counts_and_refuses_overfull_cells():
    b = synth_batch(6, 3, 20, 6000, seed=2, variant_frac=1.0)   # thousands of reads per value: bins of 32, counts of 2047
    nb, want = check_pack(b)
    assert (want["bins16"] >> 11).max() == 31 and (want["ic16"] & 0x7FF).max() == 2047
    # a cell with more than 255 narrow entries cannot be described by a length byte
    deep = synth_batch(2, 2, 20, 40000, seed=3, variant_frac=1.0)
    with pytest.raises(EngineError, match="cannot be packed"):
        NarrowBatch(deep, pinned=False)


def test_narrow_packer_handles_empty_batches():
    b = synth_batch(4, 5, 10, 30, seed=1)
    empty = b.select(np.zeros(0, np.int64))
    nb = NarrowBatch(empty, pinned=False)
    assert nb.n_sites == 0


# ---------------------------------------------------------------- GPU ----------------------------------------------------------------

SCEN = [("c2", 300, 50, 48, 200, {}, dict(variant_frac=0.3)),
        ("bidir", 200, 10, 20, 100, {}, dict(variant_frac=0.4, bidir_frac=0.2)),
        ("deep", 60, 6, 20, 5000, {}, dict(variant_frac=0.5)),
        ("manyquals", 150, 10, 20, 100, {}, dict(variant_frac=0.4, bidir_frac=0.1, quals=range(14, 42))),
        ("indel", 200, 10, 20, 100, {}, dict(variant_frac=0.3, indel_frac=0.5, bidir_frac=0.05)),
        ("perm", 120, 10, 20, 100, dict(chisq_permutation=1, max_iter=2000), dict(variant_frac=0.4, with_qhigh=True)),
        ("em0", 200, 10, 20, 100, dict(em_flag=0), dict(variant_frac=0.4, with_stats=True, with_qhigh=True)),
        ("nobq", 150, 10, 20, 100, dict(use_base_qvs=0), dict(variant_frac=0.4)),
        ("diploid", 300, 40, 2, 10, {}, dict(variant_frac=0.5)),
        ("c3", 40, 100, 96, 60, {}, dict(variant_frac=0.5))]


@pytest.mark.gpu
@pytest.mark.parametrize("name,n,P,ploidy,depth,opts,kw", SCEN, ids=[s[0] for s in SCEN])
def test_engine_results_on_narrow_batch_equal_compact_and_reads(name, n, P, ploidy, depth, opts, kw):
    from crisp_b200.engine import Engine, pin_batch
    cfg = EngineConfig(ploidy=np.full(P, ploidy), options=opts)
    b = synth_batch(n, P, cfg.ploidy, depth, seed=21, **kw)
    eng = Engine(cfg, device=0)
    want = eng.run(b.compact())
    got = eng.run(NarrowBatch(b))
    t = eng.timing()
    compare_results(got, want, tol=0.0, label=f"{name}: narrow vs compact")
    if opts.get("chisq_permutation"):
        assert np.array_equal(got.perm_k, want.perm_k) and np.array_equal(got.perm_hits, want.perm_hits)
    assert t["h2d_bytes"] < pin_batch(b.compact()).nbytes() or n < 64
    # resident path: upload once, run twice (the widening kernels are part of every step)
    eng.upload(NarrowBatch(b))
    r1 = eng.run_resident(fetch=True)
    r2 = eng.run_resident(fetch=True)
    compare_results(r1, want, tol=0.0, label=f"{name}: narrow resident")
    compare_results(r2, want, tol=0.0, label=f"{name}: narrow resident again")
    # and against the per-read form (the likelihood sums run in another order there: 1e-9, not bit for bit)
    compare_results(got, eng.run(b), label=f"{name}: narrow vs reads", allow_flips=1)
    # the context goes back and forth between the forms
    compare_results(eng.run(b.compact()), want, tol=0.0, label=f"{name}: compact after narrow")
    empty = NarrowBatch(b.select(np.zeros(0, np.int64)))
    assert eng.run(empty).n_sites == 0
    eng.close()
