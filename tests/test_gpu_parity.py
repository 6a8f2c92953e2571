"""GPU parity tests (B200): the CUDA engine, called through the C ABI, against
  * the oracle restatement on seeded batches (every mode the reference has),
  * the unmodified reference (oracle/_ref, prebuilt; travels with the snapshot) incl. the VCF text of the
    reference's own printer fed with GPU results,
  * the committed golden vectors,
and through size-independent properties at the benchmark shape.
Tolerances (north_star): integer fields and per-pool allele counts bit-exact; p-values, likelihood-ratio statistics
and EM frequencies within 1e-9 relative; permutation p-values exact against the same counter-based streams and
within Monte-Carlo error against the reference's drand48 streams."""
import numpy as np
import pytest

from crisp_b200.abi import EngineConfig, concat_batches
from crisp_b200.engine import Engine, EngineError, pin_batch
from crisp_b200.shard import merge_results, shard_batch
from crisp_b200.synth import survey_demo_batch, synth_batch
from golden_io import golden_names, load_golden
from oracle.loader import RNG_DRAND48, RNG_PHILOX, OracleEngine, RefEngine, have_ref
from util import REL_TOL, compare_results, rel_err

pytestmark = pytest.mark.gpu

SCENARIOS = [
    # name, sites, P, ploidy, depth, options, synth kwargs
    ("default_c1", 400, 10, 20, 100, {}, dict(variant_frac=0.4, bidir_frac=0.1, second_alt_frac=0.3)),
    ("exome_c2", 160, 50, 48, 200, {}, dict(variant_frac=0.3)),
    ("chr20_c3", 48, 100, 96, 60, {}, dict(variant_frac=0.5)),
    ("lowcov_c5", 48, 500, 10, 10, {}, dict(variant_frac=0.5)),
    ("varpool", 300, 8, [10, 20, 30, 40, 10, 20, 30, 40], 120, {}, dict(variant_frac=0.4, bidir_frac=0.1)),
    ("manyquals", 200, 10, 20, 100, {}, dict(variant_frac=0.4, bidir_frac=0.1, quals=range(14, 42))),
    ("twelvequals_c2", 100, 50, 48, 200, {}, dict(variant_frac=0.3, quals=range(20, 42, 2))),
    ("deep_pool", 40, 4, 40, 3000, {}, dict(variant_frac=0.5)),
    ("big_poolsize_sparse", 60, 6, [150, 150, 150, 140, 140, 140], 400, {}, dict(variant_frac=0.6)),   # genotypes beyond j = 127
    ("big_poolsize_dense", 60, 6, 200, 400, {}, dict(variant_frac=0.6)),
    ("very_deep_sparse", 6, 2, [20, 24], 70000, {}, dict(variant_frac=1.0)),                            # > 65504 reads per pool: 16-bit bins fold
    ("very_deep_dense", 6, 2, 20, 70000, {}, dict(variant_frac=1.0)),
    ("pivot", 300, 12, 16, 80, dict(pivot_sample=5), dict(variant_frac=0.4)),
    ("indel", 300, 10, 20, 100, {}, dict(variant_frac=0.3, indel_frac=0.5, bidir_frac=0.05)),
    ("nobq", 200, 10, 20, 100, dict(use_base_qvs=0), dict(variant_frac=0.4)),
    ("refbias", 200, 10, 20, 100, dict(agilent_ref_bias=0.55), dict(variant_frac=0.4)),
    ("nofastfilter", 200, 10, 20, 100, dict(fast_filter=0), dict(variant_frac=0.3)),
    ("perm_c1", 150, 10, 20, 100, dict(chisq_permutation=1, max_iter=20000), dict(variant_frac=0.4, bidir_frac=0.05, with_qhigh=True)),
    ("perm_rare_c4", 60, 50, 48, 200, dict(chisq_permutation=1, max_iter=30000), dict(variant_frac=0.7, rare=True)),
    ("perm_c4_full_1e6", 5, 50, 48, 200, dict(chisq_permutation=1, max_iter=1000000), dict(variant_frac=1.0, rare=True)),      # BASELINE config 4 at its stated permutation count
    ("perm_c4_em0_ctpval6", 40, 50, 48, 200, dict(chisq_permutation=1, max_iter=20000, em_flag=0, thresh1=-6.0),
     dict(variant_frac=0.8, rare=True, with_stats=True, with_qhigh=True)),                                                   # ... and with --EM 0 --ctpval -6 (the only consumer of --ctpval)
    ("perm_c5", 40, 500, 10, 10, dict(chisq_permutation=1, max_iter=2000), dict(variant_frac=0.5)),
    ("perm_deep_wide", 60, 6, 20, 700, dict(chisq_permutation=1, max_iter=3000), dict(variant_frac=0.5)),          # > 255 reads per strand: 32-bit occupancy words in the tail
    ("lowcov_deep_wide", 100, 20, 2, 90, dict(chisq_permutation=1, max_iter=3000), dict(variant_frac=0.5, bidir_frac=0.1)),  # > 31 reads per stratum
    ("lowcovFET", 300, 40, 2, 10, dict(chisq_permutation=1, max_iter=2000), dict(variant_frac=0.5, bidir_frac=0.1)),
    ("diploid", 300, 40, 2, 10, {}, dict(variant_frac=0.5)),
    ("em0", 300, 10, 20, 100, dict(em_flag=0), dict(variant_frac=0.4, indel_frac=0.2, with_stats=True, with_qhigh=True)),
    ("em0_ser", 200, 10, 20, 100, dict(em_flag=0, ser=0.01), dict(variant_frac=0.4, with_stats=True)),
    ("weighted_amb", 200, 10, 20, 100, dict(use_qv=1, chisq_permutation=1, max_iter=500), dict(variant_frac=0.3, indel_frac=0.5, with_qhigh=True)),
    ("exact_p5", 200, 5, 20, 150, dict(chisq_permutation=1, max_iter=2000, exact_mode=1), dict(variant_frac=0.5, bidir_frac=0.05)),
    ("exact_p7_rare", 80, 7, 20, 40, dict(chisq_permutation=1, max_iter=2000, exact_mode=1), dict(variant_frac=0.6, rare=True)),
    ("exact_p2_indel", 200, 2, 40, 80, dict(chisq_permutation=1, max_iter=2000, exact_mode=1), dict(variant_frac=0.5, indel_frac=0.4)),
    ("exact_gate_p8", 60, 8, 20, 40, dict(chisq_permutation=1, max_iter=2000, exact_mode=1), dict(variant_frac=0.5)),  # 8 pools: Monte-Carlo only
    ("association", 200, 12, 20, 100, dict(association=1), dict(variant_frac=0.6)),
    ("association_indel", 200, 12, 20, 100, dict(association=1), dict(variant_frac=0.3, indel_frac=0.6)),
]
PHENOTYPES12 = [1, 1, 1, 1, 2, 2, 2, 2, 6, 6, 1, 2]  # 1 control, 2 case, 4 early onset (readmultiplebams.c --phenotypes)


def _cfg(P, ploidy, opts):
    ph = np.array(PHENOTYPES12, np.int32) if opts.get("association") else None
    return EngineConfig(ploidy=np.broadcast_to(np.asarray(ploidy), (P,)).copy(), options=opts, phenotypes=ph)


@pytest.mark.parametrize("name,n,P,ploidy,depth,opts,kw", SCENARIOS, ids=[s[0] for s in SCENARIOS])
def test_engine_matches_oracle(name, n, P, ploidy, depth, opts, kw):
    cfg = _cfg(P, ploidy, opts)
    batch = synth_batch(n, P, cfg.ploidy, depth, seed=11, **kw)
    want = OracleEngine(cfg).run(batch, mode=RNG_PHILOX)
    eng = Engine(cfg)
    got = eng.run(batch)
    eng.close()
    assert (want.status == 4).sum() > 0
    if opts.get("association"):
        # p = kf_gammaq(0.5, LLR ln 10): where the likelihood ratio is zero up to summation order (all subsets clamp to
        # the same allele frequency) its sign — and with it NaN vs 0 — is rounding noise in the reference itself
        compare_results(got, want, label=name, fp_skip=("assoc_p", "assoc_p1"))
        assert (want.assoc_llr != 0).sum() > 10
        for fp, fl in (("assoc_p", "assoc_llr"), ("assoc_p1", "assoc_llr1")):
            m = (np.abs(getattr(want, fl)) > 1e-9) & (np.abs(getattr(got, fl)) > 1e-9)
            assert rel_err(getattr(got, fp)[m], getattr(want, fp)[m]).max(initial=0) <= REL_TOL, fp
    elif P == 2 and opts.get("chisq_permutation"):
        # two pools: the reported chi-square has 0 degrees of freedom, kf_gammaq(0, z <= 1) = log(1 - exp(~1e-16)) is
        # rounding noise (NaN or about -16) in the reference itself; it feeds no decision in this mode
        compare_results(got, want, label=name, fp_skip=("chi2pval",))
    else:
        compare_results(got, want, label=name)
    if opts.get("chisq_permutation"):
        # same counter-based streams, same sequential stop rule: hit counts and stop indices are identical
        assert np.array_equal(got.perm_k, want.perm_k)
        assert np.array_equal(got.perm_hits, want.perm_hits)
    if opts.get("exact_mode") and P < 8:
        assert (want.perm_hits == -1).sum() > 20  # decided by the exact test


def _needs_ref():
    """Checked when the test runs: the session fixture builds oracle/_ref (where /root/reference exists) after collection."""
    if not have_ref():
        pytest.skip("oracle/_ref not present")


_VCF_SCENARIOS = [s for s in SCENARIOS if not s[5].get("chisq_permutation") and not s[5].get("association")
                  and s[0] not in ("very_deep_sparse", "very_deep_dense", "chr20_c3", "lowcov_c5")]


@pytest.mark.parametrize("name,n,P,ploidy,depth,opts,kw", _VCF_SCENARIOS, ids=[s[0] for s in _VCF_SCENARIOS])
def test_engine_matches_reference_and_vcf(name, n, P, ploidy, depth, opts, kw):
    """Against CRISP's own code: statistics within tolerance and — feeding the GPU results to the reference's
    unchanged print_pooledvariant — identical VCF records (modulo sites within tolerance of a print rounding)."""
    _needs_ref()
    cfg = _cfg(P, ploidy, opts)
    batch = synth_batch(n, P, cfg.ploidy, depth, seed=23, **kw)
    ref = RefEngine(cfg)
    want = ref.run(batch, want_vcf=True)
    eng = Engine(cfg)
    got = eng.run(batch)
    eng.close()
    compare_results(got, want, label=name)
    gv, wv = ref.replay_vcf(batch, got).splitlines(), want.vcf.splitlines()
    assert len(gv) == len(wv) and len(wv) > 0
    diff = [i for i, (a, b) in enumerate(zip(gv, wv)) if a != b]
    # a printed digit may differ only where a value sits within 1e-9 of a rounding boundary
    assert len(diff) <= max(1, len(wv) // 200), (len(diff), gv[diff[0]], wv[diff[0]])


@pytest.mark.parametrize("name", ["perm_c1", "lowcovFET", "weighted_amb"])
def test_permutation_results_print_the_reference_vcf(name):
    """Permutation mode through the reference's printer: the GPU's results, with only the Monte-Carlo p-value taken from the
    reference's own drand48 run (so that both print the same CT= digits), give the reference's records for every site whose
    call does not hinge on that p-value."""
    _needs_ref()
    sc = next(s for s in SCENARIOS if s[0] == name)
    _, n, P, ploidy, depth, opts, kw = sc
    cfg = _cfg(P, ploidy, opts)
    batch = synth_batch(n, P, cfg.ploidy, depth, seed=29, **kw)
    ref = RefEngine(cfg)
    want = ref.run(batch, want_vcf=True, seed=3)
    eng = Engine(cfg)
    got = eng.run(batch)
    eng.close()
    same_site = np.all(got.status == want.status, axis=1)   # decisions the Monte-Carlo noise did not flip
    assert same_site.mean() > 0.8
    got.ctpval[...] = want.ctpval
    gv = {l.split("\t")[0]: l for l in ref.replay_vcf(batch, got).splitlines()}
    wv = {l.split("\t")[0]: l for l in want.vcf.splitlines()}
    keys = [k for k in wv if same_site[int(k)]]
    assert len(keys) > 10
    diff = [k for k in keys if gv.get(k) != wv[k]]
    assert len(diff) <= max(1, len(keys) // 100), (len(diff), gv.get(diff[0]), wv[diff[0]])


def test_permutation_null_law_against_enumeration():
    """The sampler's null law against ground truth, not against a second Monte-Carlo run: three pools and eight alternate
    reads, so every placement of the forward and of the reverse alt reads (hypergeometric x hypergeometric, the reference's
    conditional null of contables.c:308-331) can be enumerated and the exact probability of `statistic <= observed + 1e-6`
    computed.  Tables are chosen with that probability in [5e-5, 4e-4]: the stop rule (10 hits, or 5 within 1000) then
    rarely fires, the test runs all 20 000 permutations, and the hit count is Poisson around 20 000 x p."""
    from itertools import product
    from math import comb, log10
    from util import batch_from_counts
    rng = np.random.default_rng(77)
    P, tables, exact = 3, [], []

    def placements(total, caps):
        out = []
        for x in product(*[range(min(total, int(c)) + 1) for c in caps]):
            if sum(x) == total:
                w = 1.0
                for xi, ci in zip(x, caps):
                    w *= comb(int(ci), xi)
                out.append((np.array(x), w))
        return out

    while len(tables) < 48:
        Nf, Nr = rng.integers(10, 22, P), rng.integers(10, 22, P)
        af, ar = np.zeros(P, np.int64), np.zeros(P, np.int64)
        af[0], ar[0] = rng.integers(3, 6), rng.integers(3, 6)      # the alt reads sit in pool 0 ...
        if rng.random() < 0.5:
            af[1] = 1                                               # ... sometimes with one stray read elsewhere
        if np.any(af > Nf) or np.any(ar > Nr):
            continue
        N = Nf + Nr
        stat = lambda x: sum(log10(comb(int(N[i]), int(x[i]))) for i in range(P))
        obs = stat(af + ar)
        pf, pr = placements(int(af.sum()), Nf), placements(int(ar.sum()), Nr)
        zf, zr = sum(w for _, w in pf), sum(w for _, w in pr)
        tail = sum(wf * wr for xf, wf in pf for xr, wr in pr if stat(xf + xr) <= obs + 1e-6) / (zf * zr)
        if 5e-5 <= tail <= 4e-4:
            ic = np.zeros((P, 24), np.int64)
            ic[:, 0], ic[:, 8] = Nf - af, Nr - ar     # reference allele A
            ic[:, 2], ic[:, 10] = af, ar              # alternate allele G
            tables.append(ic)
            exact.append(tail)
    batch = batch_from_counts(np.array(tables), np.zeros(len(tables), np.uint8))
    cfg = _cfg(P, 20, dict(chisq_permutation=1, max_iter=20000, fast_filter=0))
    eng = Engine(cfg)
    got = eng.run(batch)
    eng.close()
    exact = np.array(exact)
    k, h = got.perm_k[:, 0], got.perm_hits[:, 0]
    full = k >= 20000                                   # ran every permutation
    assert full.sum() >= 36, (k, h)
    expect = 20000 * exact[full]
    z_total = (h[full].sum() - expect.sum()) / np.sqrt(expect.sum())
    assert abs(z_total) < 4.0, (z_total, h[full].sum(), expect.sum())
    z_each = (h[full] - expect) / np.sqrt(np.maximum(expect, 1.0))
    assert np.abs(z_each).max() < 6.0, z_each
    # the early stops are consistent too: a table that stopped at 10 hits did so around 10 / p permutations
    for kk, hh, p in zip(k[~full], h[~full], exact[~full]):
        assert hh >= 5 and kk > 0.05 * hh / p, (kk, hh, p)


def test_deep_tables_beyond_the_packed_counters_are_refused():
    """More alternate reads of one stratum in one pool than the 10-bit occupancy fields of the low-coverage test hold: the
    engine reports CRISPGPU_ERR_UNSUPPORTED instead of sampling from a corrupted layout (the reference handles any depth)."""
    cfg = _cfg(4, 2, dict(chisq_permutation=1, max_iter=2000))
    batch = synth_batch(6, 4, cfg.ploidy, 6000, seed=3, variant_frac=1.0)
    eng = Engine(cfg)
    with pytest.raises(EngineError, match="packed occupancy"):
        eng.run(batch)
    eng.close()
    # the same depth is fine for pooled data (16-bit fields per strand)
    cfg = _cfg(4, 20, dict(chisq_permutation=1, max_iter=2000))
    batch = synth_batch(6, 4, cfg.ploidy, 6000, seed=3, variant_frac=1.0)
    want = OracleEngine(cfg).run(batch, mode=RNG_PHILOX)
    eng = Engine(cfg)
    got = eng.run(batch)
    eng.close()
    compare_results(got, want, label="deep_pooled_perm")
    assert np.array_equal(got.perm_k, want.perm_k) and np.array_equal(got.perm_hits, want.perm_hits)


def test_offsets_are_validated_when_asked():
    """tests/conftest.py switches CRISPGPU_VALIDATE on: a non-monotone offset array is an argument error, not a device fault."""
    cfg = _cfg(6, 20, {})
    batch = synth_batch(20, 6, cfg.ploidy, 40, seed=2, variant_frac=1.0)
    bad = batch.select(np.arange(batch.n_sites))
    bad.read_off = bad.read_off.copy()
    bad.read_off[7] = bad.read_off[9] + 5
    eng = Engine(cfg)
    with pytest.raises(EngineError, match="read_off decreases"):
        eng.run(bad)
    got = eng.run(batch)   # the context stays usable
    eng.close()
    assert got.n_sites == batch.n_sites


def test_compact_batch_without_any_bin():
    """A compact batch whose cells are all empty (every read filtered) must take the bins path with empty ranges — the
    dense-reads path has no read offsets to index."""
    from crisp_b200.abi import Batch
    cfg = _cfg(6, 20, {})
    b = synth_batch(30, 6, cfg.ploidy, 40, seed=2, variant_frac=1.0)
    n, P = b.n_sites, 6
    common = dict(tid=b.tid, pos=b.pos, refbase=b.refbase, maxvec_al=b.maxvec_al, maxvec_ct=b.maxvec_ct, counts=b.counts, indcounts=b.indcounts)
    no_reads = Batch(P, reads=np.zeros(0, np.uint16), read_off=np.zeros(n * P + 1, np.uint32), **common)
    no_bins = Batch(P, bins=np.zeros(0, np.uint32), bin_off=np.zeros(n * P + 1, np.uint32), **common)
    want = OracleEngine(cfg).run(no_reads, mode=RNG_PHILOX)
    eng = Engine(cfg)
    got = eng.run(no_bins)
    eng.close()
    compare_results(got, want, label="compact-without-bins")


@pytest.mark.parametrize("name", golden_names())
def test_engine_matches_golden(name):
    cfg, batch, want, _ = load_golden(name)
    eng = Engine(cfg)
    got = eng.run(batch)
    eng.close()
    perm = bool(cfg.options["chisq_permutation"])
    if not perm:
        compare_results(got, want, label=name)
        return
    # golden permutation p-values come from the reference's drand48 stream: compare within Monte-Carlo error and
    # leave the slots whose downstream decision depends on that p-value to the statistical check
    tested = want.perm_k > 0 if hasattr(want, "perm_k") else got.perm_k > 0
    k = np.maximum(got.perm_k, 1)
    pg, pw = 10.0 ** got.ctpval, 10.0 ** want.ctpval
    m = got.perm_k > 0
    # two independent Monte-Carlo estimates of the same p: each stops after ~10 hits, i.e. carries a relative error of about
    # 1/sqrt(10) when it stopped early; the difference of the two is judged against the sum of their binomial variances
    pm = np.maximum(pw, pg)
    sigma = np.sqrt(2 * pm * (1 - np.minimum(pw, pg)) / k) + 3.0 / k
    assert np.all(np.abs(pg - pw)[m] <= 6 * sigma[m]), np.max(np.abs(pg - pw)[m] / sigma[m])
    compare_results(got, want, fp_skip=("ctpval",), label=name, allow_flips=max(2, int(0.1 * m.sum())))


def test_permutation_pvalues_within_monte_carlo_error_of_reference():
    """Reference arm (drand48, seeded) vs engine (Philox) on sites that run all permutations: the hit counts of two
    independent Monte-Carlo runs of the same test differ by binomial noise only."""
    cfg = _cfg(10, 20, dict(chisq_permutation=1, max_iter=20000))
    batch = synth_batch(120, 10, cfg.ploidy, 100, seed=31, variant_frac=0.6)
    Checker = RefEngine if have_ref() else OracleEngine
    want = Checker(cfg).run(batch, seed=5) if have_ref() else Checker(cfg).run(batch, mode=RNG_DRAND48, seed=5)
    eng = Engine(cfg)
    got = eng.run(batch)
    eng.close()
    m = got.perm_k > 0
    assert m.sum() > 50
    pg = (got.perm_hits[m] + 1.0) / (got.perm_k[m] + 1.0)
    pw = 10.0 ** want.ctpval[m]
    kmin = np.minimum(got.perm_k[m], 20001)
    p = 0.5 * (pg + pw)
    sigma = np.sqrt(2 * p * (1 - p) / np.minimum(kmin, 1000)) + 3.0 / kmin
    z = np.abs(pg - pw) / sigma
    assert np.quantile(z, 0.95) < 3.5 and z.max() < 8, (np.quantile(z, 0.95), z.max())
    # floor of the estimator: log10(1/(MAXITER+2)) (contables.c:336)
    assert got.ctpval[m].min() >= np.log10(1.0 / 20002) - 1e-12


def test_empty_and_degenerate_batches():
    cfg = _cfg(6, 20, {})
    eng = Engine(cfg)
    b = survey_demo_batch()
    empty = b.select(np.zeros(0, np.int64))
    r = eng.run(empty)
    assert r.n_sites == 0 and r.n_calls == 0
    # a site whose pools have no reads at all, and a site with reads in one pool only
    one = synth_batch(20, 6, 20, 40, seed=2, variant_frac=1.0)
    ic = one.indcounts.reshape(one.n_sites, 6, 24).copy()
    ro = one.read_off.astype(np.int64)
    keep = np.zeros(one.reads.size, bool)
    for s in range(one.n_sites):
        keep[ro[s * 6]:ro[s * 6 + 1]] = True  # pool 0 only
    ic[:, 1:, :] = 0
    per = (ro[1:] - ro[:-1]).reshape(one.n_sites, 6).copy()
    per[:, 1:] = 0
    from crisp_b200.abi import Batch
    from crisp_b200.synth import sort_alleles
    counts = ic.sum(axis=1)
    al, ct = sort_alleles(counts.reshape(-1, 3, 8).sum(axis=1), one.refbase)
    ragged = Batch(6, tid=one.tid, pos=one.pos, refbase=one.refbase, maxvec_al=al.reshape(-1), maxvec_ct=ct.reshape(-1),
                   counts=counts.reshape(-1), indcounts=ic.reshape(-1), reads=one.reads[keep],
                   read_off=np.concatenate([[0], np.cumsum(per.reshape(-1))]).astype(np.uint32))
    want = OracleEngine(cfg).run(ragged, mode=RNG_PHILOX)
    got = eng.run(ragged)
    compare_results(got, want, label="ragged")
    eng.close()


@pytest.mark.parametrize("P,ploidy,depth,n", [(1, 20, 200, 50), (2, 20, 100, 50), (2048, 4, 6, 6)], ids=["one_pool", "two_pools", "max_pools_2048"])
def test_extreme_pool_counts(P, ploidy, depth, n):
    """One pool, two pools, and the largest supported pool count (64 rounds of 32 lanes) against the oracle; one pool
    more than that is refused at context creation, not mis-computed."""
    cfg = _cfg(P, ploidy, {})
    batch = synth_batch(n, P, cfg.ploidy, depth, seed=5, variant_frac=1.0)
    want = OracleEngine(cfg).run(batch, mode=RNG_PHILOX)
    eng = Engine(cfg)
    got = eng.run(batch)
    packed = eng.run(pin_batch(batch.compact()))
    eng.close()
    assert (want.status >= 3).sum() > 0
    # P <= 2: the chi-square has <= 0 degrees of freedom; kf_gammaq's value there is rounding noise in the reference itself
    skip = ("chi2pval", "ctpval") if P <= 2 else ()
    compare_results(got, want, label=f"P={P}", fp_skip=skip)
    compare_results(packed, got, tol=1e-12, label=f"P={P}-compact", fp_skip=skip)
    if P == 2048:
        with pytest.raises(EngineError):
            Engine(_cfg(P + 1, ploidy, {}))


def _random_cases():
    rng = np.random.default_rng(2024)
    cases = []
    for k in range(16):
        P = int(rng.choice([3, 7, 31, 32, 33, 64, 65]))
        ploidy = int(rng.choice([2, 5, 20, 31, 32, 63, 64, 127, 128, 130]))   # genotype counts around the 32/64/128 lane boundaries
        depth = int(rng.choice([5, 40, 300]))
        opts = [{}, dict(chisq_permutation=1, max_iter=1500), dict(em_flag=0), dict(use_base_qvs=0), dict(pivot_sample=max(1, P // 2)),
                dict(agilent_ref_bias=0.52), dict(fast_filter=0)][int(rng.integers(0, 7))]
        kw = dict(variant_frac=float(rng.choice([0.3, 0.7])), indel_frac=float(rng.choice([0.0, 0.3])), bidir_frac=float(rng.choice([0.0, 0.1])))
        if rng.random() < 0.3:
            kw["quals"] = range(12, 12 + int(rng.integers(17, 30)))           # more than 16 qualities: sparse likelihood path
        if opts.get("em_flag") == 0:
            kw.update(with_stats=True, with_qhigh=True)
        if opts.get("chisq_permutation"):
            kw.update(with_qhigh=True)
        cases.append((f"r{k}_P{P}_n{ploidy}_d{depth}_" + "_".join(sorted(opts)) or "default", P, ploidy, depth, opts, kw))
    return cases


@pytest.mark.parametrize("name,P,ploidy,depth,opts,kw", _random_cases(), ids=[c[0] for c in _random_cases()])
def test_random_shapes_match_oracle(name, P, ploidy, depth, opts, kw):
    """Pool counts and pool sizes around the warp and chunk boundaries, random options, dense and compact batches."""
    cfg = _cfg(P, ploidy, opts)
    batch = synth_batch(60, P, cfg.ploidy, depth, seed=77, **kw)
    want = OracleEngine(cfg).run(batch, mode=RNG_PHILOX)
    eng = Engine(cfg)
    got = eng.run(batch)
    packed = eng.run(pin_batch(batch.compact()))
    eng.close()
    compare_results(got, want, label=name)
    compare_results(packed, want, label=name + "-compact")
    if opts.get("chisq_permutation"):
        for r in (got, packed):
            assert np.array_equal(r.perm_k, want.perm_k) and np.array_equal(r.perm_hits, want.perm_hits)


def test_results_independent_of_batching_and_sharding():
    """Region shards (as at N GPUs) and arbitrary batch splits give bit-identical results, including the
    permutation test (streams are keyed by site)."""
    cfg = _cfg(10, 20, dict(chisq_permutation=1, max_iter=3000))
    batch = synth_batch(200, 10, cfg.ploidy, 80, seed=17, variant_frac=0.5)
    eng = Engine(cfg)
    whole = eng.run(batch)
    parts, idxs = [], []
    for r in range(4):
        sub, idx = shard_batch(batch, 4, r)
        parts.append(eng.run(sub))
        idxs.append(idx)
    merged = merge_results(parts, idxs, batch.n_sites, 10)
    compare_results(merged, whole, tol=0.0, label="sharded")
    assert np.array_equal(merged.perm_k, whole.perm_k) and np.array_equal(merged.perm_hits, whole.perm_hits)
    # twice the same batch: identical
    again = eng.run(batch)
    compare_results(again, whole, tol=0.0, label="repeat")
    eng.close()


@pytest.mark.parametrize("ploidy", [20, [2, 20, 2, 20, 2, 20, 2, 20]], ids=["pooled", "mixed-diploid"])
def test_candidate_screen_matches_oracle(ploidy):
    """crispgpu_screen_positions (k_candidates): sort_allelecounts and the potential-variant score, dense and sparse
    per-pool counts; and a batch submitted without maxvec gives the results of the batch with it."""
    from test_oracle import _screen_batch
    cfg, batch = _screen_batch(9, ploidy=ploidy, n=700)
    want = OracleEngine(cfg).screen(batch)
    eng = Engine(cfg)
    for form in (batch, batch.compact()):
        got = eng.screen_positions(form)
        for g, w in zip(got, want):
            assert np.array_equal(g, w)
    # the synthetic generator sorted with the unscrambled counts; re-sort for this batch, then drop maxvec
    batch.maxvec_al, batch.maxvec_ct = want[0].reshape(-1).copy(), want[1].reshape(-1).copy()
    with_maxvec = eng.run(batch)
    compare_results(eng.run(batch.without_maxvec()), with_maxvec, tol=0.0, label="device-sorted")
    # several pool sizes take the bin-by-bin likelihood path: same terms as the per-read histogram, other summation order
    compare_results(eng.run(pin_batch(batch.compact().without_maxvec())), with_maxvec, tol=0.0 if np.ndim(ploidy) == 0 else 1e-12, label="device-sorted-compact")
    compare_results(with_maxvec, OracleEngine(cfg).run(batch, mode=RNG_PHILOX), label="oracle")
    eng.close()


def test_context_reuse_across_different_batches():
    """A context carries device buffers, flags and the k_em instantiation from batch to batch; whatever came before
    (other qualities, bidirectional reads, far fewer or far more tasks, the other batch format), the results are those of
    a fresh context, bit for bit."""
    cfg = _cfg(10, 20, {})

    def fresh(batch):
        e = Engine(cfg)
        r = e.run(batch)
        e.close()
        return r

    a = synth_batch(600, 10, cfg.ploidy, 100, seed=3, variant_frac=0.3, indel_frac=0.2)
    b = synth_batch(600, 10, cfg.ploidy, 100, seed=4, variant_frac=0.3, indel_frac=0.2)      # same shape, other sites
    other_quals = synth_batch(600, 10, cfg.ploidy, 100, seed=5, variant_frac=0.3, quals=[22, 33])
    with_bidir = synth_batch(600, 10, cfg.ploidy, 100, seed=6, variant_frac=0.3, bidir_frac=0.1)
    few = a.select(np.arange(40))                                                              # sizes the rows for 40 sites ...
    many = synth_batch(3000, 10, cfg.ploidy, 100, seed=7, variant_frac=0.6)                   # ... then far more tasks arrive
    fa, fb = fresh(a), fresh(b)
    compare_results(fa, OracleEngine(cfg).run(a, mode=RNG_PHILOX), label="oracle")
    eng = Engine(cfg)
    seq = [("a", a, fa), ("a-again", a, fa), ("b", b, fb), ("quals", other_quals, fresh(other_quals)), ("quals-again", other_quals, None),
           ("bidir", with_bidir, fresh(with_bidir)), ("a-back", a, fa), ("few", few, fresh(few)), ("many", many, fresh(many)), ("many-again", many, None)]
    last = None
    for name, batch, want in seq:
        for form in (batch, pin_batch(batch.compact())):
            got = eng.run(form)
            compare_results(got, want if want is not None else last, tol=0.0, label=name)
        last = got if want is None else want
    eng.upload(b)
    compare_results(eng.run_resident(fetch=True), fb, tol=0.0, label="resident-first")
    compare_results(eng.run_resident(fetch=True), fb, tol=0.0, label="resident-again")
    eng.close()


def test_contexts_of_different_shapes_share_the_process():
    """Kernel attributes (dynamic shared-memory limits) are process-wide; contexts with small and large pool counts
    must be able to alternate, also from different threads."""
    import threading
    small, large = _cfg(10, 20, {}), _cfg(100, 96, {})
    bs = synth_batch(200, 10, small.ploidy, 100, seed=5, variant_frac=0.4)
    bl = synth_batch(24, 100, large.ploidy, 60, seed=6, variant_frac=0.5)
    ws, wl = OracleEngine(small).run(bs, mode=RNG_PHILOX), OracleEngine(large).run(bl, mode=RNG_PHILOX)
    es, el = Engine(small), Engine(large)
    for _ in range(2):
        compare_results(es.run(bs), ws, label="small")
        compare_results(el.run(bl), wl, label="large")
    out = {}

    def loop(name, eng, batch):
        out[name] = [eng.run(batch) for _ in range(6)]

    th = [threading.Thread(target=loop, args=("s", es, bs)), threading.Thread(target=loop, args=("l", el, bl))]
    for t in th:
        t.start()
    for t in th:
        t.join()
    for r in out["s"]:
        compare_results(r, ws, label="small-threaded")
    for r in out["l"]:
        compare_results(r, wl, label="large-threaded")
    es.close()
    el.close()


def test_arena_must_contain_every_array():
    cfg = _cfg(10, 20, {})
    batch = pin_batch(synth_batch(50, 10, cfg.ploidy, 100, seed=2, variant_frac=0.4).compact())
    eng = Engine(cfg)
    good = eng.run(batch)
    outside = pin_batch(batch, arena=False)
    outside._arena = batch._arena            # declares an arena its arrays do not live in
    with pytest.raises(EngineError, match="outside the declared arena"):
        eng.run(outside)
    compare_results(eng.run(batch), good, tol=0.0, label="after-error")
    eng.close()


def test_advance_between_submit_and_wait():
    """crispgpu_advance queues stage 2 early; results and error behaviour are those of a plain submit/wait."""
    cfg = _cfg(10, 20, {})
    batch = pin_batch(synth_batch(300, 10, cfg.ploidy, 100, seed=19, variant_frac=0.4).compact())
    e0, e1 = Engine(cfg), Engine(cfg)
    plain = e0.run(batch)
    t0, t1 = e0.submit(batch), e1.submit(batch)
    e1.advance(t1)
    e1.advance(t1)            # idempotent
    e0.advance(t0)
    a, b = e0.wait(t0), e1.wait(t1)
    compare_results(a, plain, tol=0.0, label="advance-0")
    compare_results(b, plain, tol=0.0, label="advance-1")
    with pytest.raises(EngineError):
        e0.advance(t0)        # nothing in flight any more
    e0.close()
    e1.close()


def test_submit_wait_equals_resident_path():
    cfg = _cfg(10, 20, {})
    batch = synth_batch(300, 10, cfg.ploidy, 100, seed=19, variant_frac=0.4)
    eng = Engine(cfg)
    a = eng.run(pin_batch(batch))
    eng.upload(batch)
    eng.run_resident(fetch=False)
    b = eng.run_resident(fetch=True)
    compare_results(b, a, tol=0.0, label="resident")
    t = eng.timing()
    assert t["launches"] >= 3 and t["n_tasks_em"] > 0 and t["em_ms"] > 0
    eng.close()
    # blocking_wait only changes how the host thread waits
    sleepy = Engine(_cfg(10, 20, dict(blocking_wait=1)))
    compare_results(sleepy.run(pin_batch(batch)), a, tol=0.0, label="blocking-wait")
    sleepy.close()


@pytest.mark.parametrize("opts,kw", [
    ({}, dict(variant_frac=0.2, bidir_frac=0.05)),
    ({}, dict(variant_frac=0.3, indel_frac=0.5, quals=range(14, 42))),
    (dict(chisq_permutation=1, max_iter=2000), dict(variant_frac=0.3)),
    (dict(use_base_qvs=0), dict(variant_frac=0.3)),
    (dict(em_flag=0), dict(variant_frac=0.3, with_stats=True, with_qhigh=True)),
], ids=["snp", "indel_manyquals", "perm", "nobq", "em0"])
def test_on_demand_read_fetch_equals_full_copy(opts, kw):
    """Pinned batches leave the packed reads in host memory and fetch only the sites that reach the quality-value
    likelihood (k_fetch_reads); pageable batches copy everything.  Same results; fewer bytes over PCIe."""
    cfg = _cfg(10, 20, opts)
    batch = synth_batch(1500, 10, cfg.ploidy, 100, seed=77, **kw)
    assert batch.reads.nbytes > (1 << 20)
    eng = Engine(cfg)
    full = eng.run(batch)
    t_full = eng.timing()
    lazy = eng.run(pin_batch(batch, arena=False))   # separate pinned arrays: the read array may stay on the host
    t_lazy = eng.timing()
    again = eng.run(batch)   # the device read buffer now holds a mix of stale and fresh sites: results must not care
    eng.close()
    assert (full.status == 4).sum() > 30
    want = OracleEngine(cfg).run(batch, mode=RNG_PHILOX)
    compare_results(lazy, want, label="lazy-vs-oracle")
    # the likelihood path (dense/sparse) is picked from the qualities seen, which may differ between the two scans
    compare_results(lazy, full, label="lazy-vs-full")
    compare_results(again, full, tol=0.0, label="full-repeat")
    if opts.get("em_flag", 1) >= 1 and opts.get("use_base_qvs", 1) == 1:
        assert t_lazy["h2d_bytes"] < t_full["h2d_bytes"]
        fetched = t_lazy["h2d_bytes"] - (t_full["h2d_bytes"] - batch.reads.nbytes)
        ro = batch.read_off.astype(np.int64)
        assert 0 < fetched <= batch.reads.nbytes
        per_site = 2 * (ro[10::10] - ro[:-1:10])
        assert fetched <= per_site[(lazy.status >= 3).any(axis=1)].sum()  # only sites with a slot that reached the caller
    else:
        # --usebq 0 / --EM 0: no kernel reads the packed reads, neither path moves them
        from crisp_b200.abi import _BATCH_ARRAYS
        total = sum(getattr(batch, name).nbytes for name, _, _ in _BATCH_ARRAYS if getattr(batch, name) is not None)
        assert t_lazy["h2d_bytes"] == t_full["h2d_bytes"] == total - batch.reads.nbytes


BINNED = ["default_c1", "exome_c2", "varpool", "manyquals", "very_deep_sparse", "very_deep_dense", "big_poolsize_sparse", "indel", "nobq",
          "perm_c1", "em0", "association"]


@pytest.mark.parametrize("name,n,P,ploidy,depth,opts,kw", [s for s in SCENARIOS if s[0] in BINNED], ids=BINNED)
def test_compact_read_format_matches_oracle(name, n, P, ploidy, depth, opts, kw):
    """The same scenarios shipped as (value, count) bins instead of one element per read (consumed as they are by k_em)."""
    cfg = _cfg(P, ploidy, opts)
    batch = synth_batch(n, P, cfg.ploidy, depth, seed=11, **kw)
    want = OracleEngine(cfg).run(batch, mode=RNG_PHILOX)
    compact = batch.binned()
    assert compact.reads is None
    if name != "manyquals":   # 28 quality values at 100x: more distinct values than reads to share them
        assert compact.bins.nbytes < batch.reads.nbytes
    eng = Engine(cfg)
    got = eng.run(compact)
    pinned = eng.run(pin_batch(compact))
    eng.upload(compact)
    resident = eng.run_resident(fetch=True)
    # ... and with the per-pool allele counts as their non-zero entries only (k_expand_counts)
    full = batch.compact()
    assert full.indcounts is None and full.ic_sparse.nbytes <= batch.indcounts.nbytes // 3   # at most 8 of 24 entries in use without indels
    sparse = eng.run(pin_batch(full, arena=False))
    t = eng.timing()
    # ... and packed into one pinned arena, moved with a single copy
    packed = pin_batch(full)
    one_copy = eng.run(packed)
    t_arena = eng.timing()
    eng.close()
    skip = ("assoc_p", "assoc_p1") if opts.get("association") else ()
    compare_results(got, want, label=name, fp_skip=skip)
    compare_results(pinned, got, tol=0.0, label=name + "-pinned")
    compare_results(resident, got, tol=0.0, label=name + "-resident")
    compare_results(sparse, got, tol=0.0, label=name + "-sparse-counts")
    compare_results(one_copy, got, tol=0.0, label=name + "-arena")
    assert t_arena["h2d_bytes"] == packed._arena[1] and full.nbytes() <= packed._arena[1] < full.nbytes() + 256 * 32
    used = full.nbytes() - (0 if opts.get("em_flag", 1) >= 1 and opts.get("use_base_qvs", 1) == 1 else full.bins.nbytes + full.bin_off.nbytes)
    assert t["h2d_bytes"] == used


@pytest.mark.parametrize("ploidy", [20, [20, 24]], ids=["dense", "sparse"])
def test_compact_read_format_splits_large_counts(ploidy):
    """Multiplicities above 65535 take several bins; both likelihood paths of k_em (one pool size: histogram + tensor-core
    contraction; several: bin-by-bin) add them up."""
    from crisp_b200.abi import bins_from_reads, reads_from_bins
    cfg = _cfg(2, ploidy, {})
    batch = synth_batch(4, 2, cfg.ploidy, 140000, seed=3, variant_frac=1.0, quals=[30])
    bo, bins = bins_from_reads(batch.read_off, batch.reads)
    assert (bins >> 16).max() == 65535 and np.array_equal(np.sort(reads_from_bins(bo, bins)[: batch.read_off[1]]), np.sort(batch.reads[: batch.read_off[1]]))
    want = OracleEngine(cfg).run(batch, mode=RNG_PHILOX)
    assert (want.status == 4).sum() >= 2
    eng = Engine(cfg)
    compare_results(eng.run(batch.binned()), want, label="split")
    eng.close()


@pytest.mark.parametrize("P,n,depth,sites", [(50, 48, 200, 600), (100, 96, 60, 300), (500, 10, 10, 160)],
                         ids=["c2_50x48", "c3_100x96", "c5_500x10"])
def test_benchmark_shape_properties(P, n, depth, sites):
    """BASELINE config 2, 3 and 5 shapes (pools x pool size at their depths): properties that hold at any size.
      * pool relabelling: permuting the pools permutes AC/QV and leaves the site statistics unchanged (1e-9);
      * strand swap: exchanging forward and reverse reads leaves delta, AF and AC unchanged and deltaf too;
      * a checksum of the integer outputs is reproducible run to run."""
    cfg = _cfg(P, n, {})
    batch = synth_batch(sites, P, n, depth, seed=101, variant_frac=0.25)
    eng = Engine(cfg)
    base = eng.run(batch)
    assert (base.status == 4).sum() > 30
    # --- pool permutation ---
    perm = np.random.default_rng(1).permutation(P)
    N = batch.n_sites
    ro = batch.read_off.astype(np.int64)
    lens = (ro[1:] - ro[:-1]).reshape(N, P)
    starts = ro[:-1].reshape(N, P)
    new_lens = lens[:, perm]
    src = np.concatenate([np.arange(starts[s, p], starts[s, p] + lens[s, p]) for s in range(N) for p in perm])
    alt = batch.alt_reads.copy()
    inv = np.argsort(perm)
    alt["pool"] = inv[alt["pool"]]
    ao = batch.alt_off.astype(np.int64)
    order = np.concatenate([ao[k] + np.argsort(alt["pool"][ao[k]:ao[k + 1]], kind="stable") for k in range(N * 5)]) if alt.size else np.zeros(0, np.int64)
    from crisp_b200.abi import Batch
    pb = Batch(P, tid=batch.tid, pos=batch.pos, refbase=batch.refbase, maxvec_al=batch.maxvec_al, maxvec_ct=batch.maxvec_ct,
               counts=batch.counts, indcounts=batch.indcounts.reshape(N, P, 24)[:, perm].reshape(-1), reads=batch.reads[src],
               read_off=np.concatenate([[0], np.cumsum(new_lens.reshape(-1))]).astype(np.uint32), alt_off=batch.alt_off,
               alt_reads=alt[order])
    pr = eng.run(pb)
    assert np.array_equal(pr.status, base.status) and np.array_equal(pr.varpools, base.varpools)
    for f in ("paf", "ctpval", "af_ml", "delta", "deltaf", "hwe"):
        assert rel_err(getattr(pr, f), getattr(base, f)).max() <= REL_TOL, f
    for o in np.nonzero(base.status.reshape(-1) == 4)[0]:
        assert np.array_equal(pr.ac[pr.call_row.reshape(-1)[o]], base.ac[base.call_row.reshape(-1)[o]][perm])
    # --- strand swap ---
    w = batch.reads.astype(np.uint16)
    sw = (w ^ np.uint16(1 << 10))
    ic = batch.indcounts.reshape(N, P, 3, 8)[:, :, [1, 0, 2], :]
    ct = batch.counts.reshape(N, 3, 8)[:, [1, 0, 2], :]
    alt2 = batch.alt_reads.copy()
    alt2["strand"] ^= 1
    sb = Batch(P, tid=batch.tid, pos=batch.pos, refbase=batch.refbase, maxvec_al=batch.maxvec_al, maxvec_ct=batch.maxvec_ct,
               counts=ct.reshape(-1), indcounts=ic.reshape(-1), reads=sw, read_off=batch.read_off, alt_off=batch.alt_off, alt_reads=alt2)
    sr = eng.run(sb)
    assert np.array_equal(sr.status, base.status) and np.array_equal(sr.allelecount, base.allelecount)
    for f in ("paf", "ctpval", "af_ml", "delta", "deltaf", "hwe"):
        assert rel_err(getattr(sr, f), getattr(base, f)).max() <= REL_TOL, f
    # --- reproducibility checksum (a shared-memory race would show up as run-to-run differences) ---
    for _ in range(3):
        again = eng.run(batch)
        for f in ("paf", "ctpval", "af_ml", "delta", "deltaf", "hwe"):
            assert np.array_equal(getattr(again, f), getattr(base, f)), f
        assert np.array_equal(again.allelecount, base.allelecount) and np.array_equal(again.varpools, base.varpools)
    again = eng.run(batch)
    assert int(again.status.astype(np.int64).sum() * 31 + again.allelecount.sum()) == int(base.status.astype(np.int64).sum() * 31 + base.allelecount.sum())
    assert np.array_equal(again.delta, base.delta)
    eng.close()


def test_abi_error_paths():
    """Misuse of the C ABI is reported through status codes and crispgpu_last_error, never by crashing."""
    import ctypes as C
    from crisp_b200.abi import CResults
    from crisp_b200.engine import EngineError
    cfg = _cfg(6, 20, {})
    eng = Engine(cfg)
    batch = synth_batch(40, 6, 20, 60, seed=5, variant_frac=0.5)
    with pytest.raises(EngineError):
        eng.run_resident()              # nothing uploaded
    t = eng.submit(batch)
    with pytest.raises(EngineError):
        eng.submit(batch)               # one batch in flight per context
    with pytest.raises(EngineError):
        eng.wait(t + 7)                 # unknown ticket
    res = eng.wait(t)
    assert res.n_sites == batch.n_sites
    with pytest.raises(EngineError):
        eng.wait(t)                     # already consumed
    cres = CResults()
    assert eng.lib.crispgpu_wait(None, 0, C.byref(cres)) == -1
    assert b"" == eng.lib.crispgpu_last_error(eng.h)[:0]
    # the context is still usable after the errors
    again = eng.run(batch)
    compare_results(again, res, tol=0.0, label="after-errors")
    eng.close()
