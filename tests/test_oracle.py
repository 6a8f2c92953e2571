"""CPU tests of the checkers themselves: the C restatement (oracle/crisp_oracle.c) against
  (1) the committed golden vectors produced by the unmodified reference (tests/golden/, scripts/make_golden.py),
  (2) the reference itself when oracle/_ref is present (built from /root/reference here; travels prebuilt to the GPU box),
  (3) the known-answer values recorded in SURVEY.md §8c.
The reference ships no tests of its own; these pin the oracle."""
import ctypes as C

import numpy as np
import pytest

from crisp_b200.abi import EngineConfig
from crisp_b200.synth import survey_demo_batch, synth_batch
from golden_io import golden_names, load_golden
from oracle.loader import MODE_CALLER, MODE_SLOTS, RNG_DRAND48, RNG_PHILOX, OracleEngine, RefEngine, have_ref
from util import FP_FIELDS, INT_FIELDS, compare_results

def needs_ref(fn):
    """Skip at run time, not at collection: the session fixture builds oracle/_ref (where /root/reference exists) only after
    the tests have been collected, so a collection-time check skips everything on a clean checkout's first run."""
    import functools

    @functools.wraps(fn)
    def wrapper(*a, **kw):
        if not have_ref():
            pytest.skip("oracle/_ref not built (needs /root/reference)")
        return fn(*a, **kw)
    return wrapper

# SURVEY.md §8c level 2: values printed by the reference's own functions
KATS = [
    ("ncr", (10, 3), 2.0791812460476247), ("ncr", (100, 50), 29.00385390977171), ("ncr", (200, 7), 12.358676364700507),
    ("fet", (3, 50, 10, 60), -0.86482538043645185), ("fet", (0, 100, 8, 90), -2.6702592500383115),
    ("minreads_pvalue", (100, 2, 0.05), -0.92715117745995934), ("minreads_pvalue", (57, 4, 0.05), -0.073356255354241187),
    ("binomial_pvalue", (100, 7, 0.01), -4.1482302106869877), ("kf_lgamma", (4.5,), 2.4537365708424428),
    ("kf_gammaq", (2.5, 10.0), -6.684827300476976), ("kf_gammaq", (2.5, 0.5), -0.038152879313638192),
]
SIGS = {
    "ncr": [C.c_int, C.c_int], "fet": [C.c_int] * 4, "minreads_pvalue": [C.c_int, C.c_int, C.c_double],
    "binomial_pvalue": [C.c_int, C.c_int, C.c_double], "kf_lgamma": [C.c_double], "kf_gammaq": [C.c_double, C.c_double],
}


def _cfg(P=6, n=20, **opts):
    return EngineConfig(ploidy=np.full(P, n), options=opts)


@pytest.mark.parametrize("fn,args,want", KATS)
def test_scalar_known_answers(fn, args, want):
    orc = OracleEngine(_cfg())
    got = orc.scalar(fn, C.c_double, SIGS[fn])(*args)
    assert got == pytest.approx(want, rel=1e-14, abs=1e-15)
    if have_ref():
        assert RefEngine(_cfg()).scalar(fn, C.c_double, SIGS[fn])(*args) == got


def test_gammaq_chi2_tail_kat():
    # kf_gammaq(4.5, 27.85)/ln 10 = -8.04787816035231
    orc = OracleEngine(_cfg())
    g = orc.scalar("kf_gammaq", C.c_double, SIGS["kf_gammaq"])(4.5, 27.85)
    assert g / np.log(10) == pytest.approx(-8.04787816035231, rel=1e-13)


def test_survey_demo_site():
    """The six-pool site of SURVEY.md §8c level 3, values printed by the reference during the survey."""
    r = OracleEngine(_cfg()).run(survey_demo_batch(), mode=RNG_DRAND48)
    assert r.status[0, 0] == 4 and r.varalleles[0] == 1
    assert r.paf[0, 0] == 1.3999860001399986
    assert r.ctpval[0, 0] == -5.0082859818506558 and r.chi2pval[0, 0] == -5.0082859818506558
    assert r.af_ml[0, 0] == 0.012768332219397762
    assert r.delta[0, 0] == 5.3010728026494967
    assert r.deltaf[0, 0] == -2.0304825902872743
    assert r.hwe[0, 0] == 0
    assert r.ac[0].tolist() == [1, 0, 0, 0, 0, 0]
    np.testing.assert_allclose(r.qv[0], [10.9095566534, 28.0831570167, 28.0946013699, 28.0825585559, 28.1000096436, 0.865157209598], rtol=1e-11)


def test_survey_demo_permutation_drand48():
    # CHISQ_PERMUTATION=1, srand48(12345): 4 hits in all 20000 permutations -> log10(5/20001)
    r = OracleEngine(_cfg(chisq_permutation=1)).run(survey_demo_batch(), mode=RNG_DRAND48, seed=12345)
    assert r.ctpval[0, 0] == -3.6021034186048251
    assert r.perm_k[0, 0] == 20001 and r.perm_hits[0, 0] == 4


@pytest.mark.parametrize("name", golden_names())
def test_oracle_reproduces_golden(name):
    cfg, batch, want, seed = load_golden(name)
    got = OracleEngine(cfg).run(batch, mode=RNG_DRAND48, seed=seed)
    for f in INT_FIELDS:
        assert np.array_equal(getattr(got, f), getattr(want, f)), f
    for f in FP_FIELDS:
        assert np.array_equal(getattr(got, f), getattr(want, f)), f  # bit-exact: same libm, same operation order
    assert got.n_calls == want.n_calls
    assert np.array_equal(got.ac, want.ac) and np.array_equal(got.qv, want.qv)


SCENARIOS = [
    ("default", 10, 20, 100, {}, dict(variant_frac=0.4, bidir_frac=0.1, second_alt_frac=0.3)),
    ("perm", 10, 20, 100, dict(chisq_permutation=1, max_iter=1000), dict(variant_frac=0.4, with_qhigh=True)),
    ("lowcov", 30, 2, 10, dict(chisq_permutation=1, max_iter=1000), dict(variant_frac=0.5, bidir_frac=0.1)),
    ("indel", 10, 20, 100, {}, dict(variant_frac=0.3, indel_frac=0.5, bidir_frac=0.05)),
    ("em0", 10, 20, 100, dict(em_flag=0), dict(variant_frac=0.4, indel_frac=0.2, with_stats=True, with_qhigh=True)),
    ("varpool", 8, [10, 20, 30, 40, 10, 20, 30, 40], 120, {}, dict(variant_frac=0.4, bidir_frac=0.1)),
    ("nobq", 10, 20, 100, dict(use_base_qvs=0), dict(variant_frac=0.4)),
    ("weighted_amb", 10, 20, 100, dict(use_qv=1, chisq_permutation=1, max_iter=500), dict(variant_frac=0.3, indel_frac=0.5, with_qhigh=True)),
]


@needs_ref
@pytest.mark.parametrize("name,P,ploidy,depth,opts,kw", SCENARIOS)
def test_oracle_equals_reference(name, P, ploidy, depth, opts, kw):
    cfg = EngineConfig(ploidy=np.broadcast_to(np.asarray(ploidy), (P,)).copy(), options=opts)
    batch = synth_batch(120, P, cfg.ploidy, depth, seed=5, **kw)
    ref = RefEngine(cfg)
    want = ref.run(batch, seed=99, want_vcf=True, want_genll=True)
    got = OracleEngine(cfg).run(batch, mode=RNG_DRAND48, seed=99, want_genll=True)
    # bit-exact, except where the batch carries an aggregated real-valued decrement (sum of per-read error
    # probabilities) that the reference subtracts read by read: same value up to summation order
    compare_results(got, want, tol=1e-12 if opts.get("use_qv") else 0.0, label=name)
    assert np.array_equal(got.genll, want.genll)
    # the results block carries everything the unchanged printer needs: replayed VCF text is identical
    assert ref.replay_vcf(batch, got) == want.vcf


@needs_ref
def test_slot_walk_equals_newCRISPcaller():
    """The harness's slot walk (which exposes per-slot state) gives the same records as calling the reference's
    newCRISPcaller itself (SNP-only sites)."""
    cfg = _cfg(P=10)
    batch = synth_batch(150, 10, cfg.ploidy, 100, seed=8, variant_frac=0.5, bidir_frac=0.1, second_alt_frac=0.3)
    ref = RefEngine(cfg)
    a = ref.run(batch, mode=MODE_SLOTS, want_vcf=True)
    b = ref.run(batch, mode=MODE_CALLER, want_vcf=True)
    assert a.vcf == b.vcf and len(a.vcf.splitlines()) > 20
    assert np.array_equal(a.varalleles, b.varalleles)


def test_philox_sampler_matches_drand48_distribution():
    """The counter-based sampler draws from the same null law as the reference's drand48 shuffle: compare hit
    frequencies over many sites (binomial tolerance), and the stop rule bookkeeping."""
    cfg = _cfg(P=8, chisq_permutation=1, max_iter=400)
    batch = synth_batch(300, 8, cfg.ploidy, 60, seed=21, variant_frac=0.5)
    orc = OracleEngine(cfg)
    a = orc.run(batch, mode=RNG_DRAND48, seed=7)
    b = orc.run(batch, mode=RNG_PHILOX)
    tested = a.perm_k > 0
    assert np.array_equal(tested, b.perm_k > 0) and tested.sum() > 100
    # per-slot Monte-Carlo p estimates agree within 5 binomial sigmas of the pooled permutation count
    pa = (a.perm_hits[tested] + 1) / (a.perm_k[tested] + 1)
    pb = (b.perm_hits[tested] + 1) / (b.perm_k[tested] + 1)
    k = np.minimum(a.perm_k[tested], b.perm_k[tested])
    p = 0.5 * (pa + pb)
    sigma = np.sqrt(2 * p * (1 - p) / k) + 2.0 / k
    assert np.all(np.abs(pa - pb) <= 6 * sigma)
    assert abs(pa.mean() - pb.mean()) < 0.02


@needs_ref
@pytest.mark.parametrize("kw", [dict(variant_frac=0.6), dict(variant_frac=0.3, indel_frac=0.6)], ids=["snp", "indel"])
def test_association_statistic_matches_reference_log_line(kw):
    """calculate_LRstatistic (crisp/LRstatistic.c:48-69) writes only an `ASSOC` line to stdout; the harness parses it.
    The oracle must agree to the printed precision (%0.2f for p and LLR, %0.3f for the odds ratio)."""
    P = 12
    ph = np.array([1, 1, 1, 1, 2, 2, 2, 2, 6, 6, 1, 2], np.int32)
    cfg = EngineConfig(ploidy=np.full(P, 20), options=dict(association=1), phenotypes=ph)
    batch = synth_batch(150, P, 20, 100, seed=3, **kw)
    want = RefEngine(cfg).run(batch)
    got = OracleEngine(cfg).run(batch, mode=RNG_DRAND48)
    m = want.assoc_llr != 0
    assert m.sum() > 40
    for f, tol in (("assoc_p", 0.00501), ("assoc_llr", 0.00501), ("assoc_or", 0.000501), ("assoc_p1", 0.00501), ("assoc_llr1", 0.00501), ("assoc_or1", 0.000501)):
        a, b = getattr(want, f)[m], getattr(got, f)[m]
        fin = np.isfinite(a) & np.isfinite(b)
        assert np.array_equal(np.isnan(a), np.isnan(b)), f
        big = np.abs(a) > 1e4  # %0.3f of a huge odds ratio carries fewer significant digits than a double compare needs
        assert np.all(np.abs(a - b)[fin & ~big] <= tol), f
    # everything else is unaffected by the association pass
    compare_results(got, want, tol=0.0, label="assoc", fp_skip=("assoc_p", "assoc_llr", "assoc_or", "assoc_p1", "assoc_llr1", "assoc_or1"))


@needs_ref
@pytest.mark.parametrize("kw", [
    dict(variant_frac=0.5, bidir_frac=0.1, second_alt_frac=0.3),
    dict(variant_frac=0.3, indel_frac=0.6, bidir_frac=0.05, with_qhigh=True, with_stats=True),
], ids=["snp", "indel"])
def test_host_shim_round_trip(kw):
    """host/crisp_shim.c is the packer a CRISP maintainer links in at variantcalls.c:115.  Rebuild the reference's
    struct VARIANT + read lists from a batch, pack them with the shim, and get the same batch back; the copy-back half
    of the shim (crisp_shim_fill_variant) is what test_oracle_equals_reference's VCF replay goes through."""
    cfg = _cfg(P=8)
    batch = synth_batch(120, 8, cfg.ploidy, 80, seed=41, **kw)
    ref = RefEngine(cfg)
    bad, msg = ref.shim_roundtrip(batch)
    assert bad == 0, msg
    # compact mode of the packer: the same read lists come back as the (value, count) bins bins_from_reads derives
    from crisp_b200.abi import bins_from_reads
    both = batch.compact()
    both.reads, both.read_off, both.indcounts = batch.reads, batch.read_off, batch.indcounts   # the harness rebuilds the reference's structures from these
    bad, msg = ref.shim_roundtrip(both)
    assert bad == 0, msg
    assert np.array_equal(bins_from_reads(batch.read_off, batch.reads)[1], both.bins)


# ---- exact 2 x k test (exact_mode; specification FET/contables_exact.c:62-117) ------------------------------------------
def _exact_tables(seed, count):
    rng = np.random.default_rng(seed)
    out = []
    for _ in range(count):
        size = int(rng.integers(2, 8))
        N = rng.integers(0, 60, size).astype(np.int32)
        if rng.random() < 0.2:
            N[int(rng.integers(0, size))] = int(rng.integers(500, 2000))
        A = np.minimum(N, rng.binomial(N, rng.choice([0.01, 0.05, 0.2]))).astype(np.int32)
        while A.sum() >= 40:
            A = A // 2
        out.append((N, A))
    return out


def _exact_call(engine, N, A):
    fn = engine.scalar("exact_2xk", C.c_double, [C.POINTER(C.c_int), C.POINTER(C.c_int), C.c_int])
    return fn(N.ctypes.data_as(C.POINTER(C.c_int)), A.ctypes.data_as(C.POINTER(C.c_int)), len(N))


def _exact_brute(N, A):
    """(strict, with ties): log10 of the summed hypergeometric probabilities of all tables with the observed margins that
    are less likely / no more likely than the observed one, in exact integer arithmetic.  The reference compares
    rounded log10 sums without an epsilon (contables_exact.c:70), so tables tied with the observed one — including the
    observed one itself — are counted or not depending on rounding; its answer lies between the two."""
    from itertools import product
    from math import comb, log10
    ones, reads = int(A.sum()), int(N.sum())
    obs = 1
    for n, a in zip(N, A):
        obs *= comb(int(n), int(a))
    strict = ties = 0
    for t in product(*[range(min(int(n), ones) + 1) for n in N[:-1]]):
        last = ones - sum(t)
        if last < 0 or last > N[-1]:
            continue
        w = comb(int(N[-1]), last)
        for n, a in zip(N, t):
            w *= comb(int(n), a)
        if w < obs:
            strict += w
        if w <= obs:
            ties += w
    den = log10(comb(reads, ones))
    return (log10(strict) - den if strict else -np.inf), log10(ties) - den


def test_exact_2xk_against_brute_force():
    orc = OracleEngine(_cfg())
    rng = np.random.default_rng(5)
    equal_hi = 0
    for _ in range(60):
        size = int(rng.integers(2, 5))
        N = rng.integers(1, 14, size).astype(np.int32)
        A = rng.binomial(N, 0.25).astype(np.int32)
        lo, hi = _exact_brute(N, A)
        got = _exact_call(orc, N, A)
        if got < -1 - 1e-9 and lo == -np.inf and got < hi - 1e-9:
            continue  # nothing qualified after rounding: the reference's "-1" sentinel minus the normaliser
        assert lo - 1e-10 <= got <= hi + 1e-10, (N, A, lo, got, hi)
        equal_hi += abs(got - hi) < 1e-10
    assert equal_hi > 20


@needs_ref
def test_exact_2xk_equals_reference():
    """compute_pvalue_exact of the unmodified FET/contables_exact.c (linked into oracle/_ref) on random small tables:
    the restatement uses the same term tables and acceptance order, so the doubles are identical."""
    orc, ref = OracleEngine(_cfg()), RefEngine(_cfg())
    for N, A in _exact_tables(3, 300):
        assert _exact_call(orc, N, A) == _exact_call(ref, N, A), (N, A)
    # outside the gate (>= 70 alt reads, >= 8 columns) both report "not applicable" as 1.0
    big = (np.full(4, 100, np.int32), np.full(4, 20, np.int32))
    assert _exact_call(orc, *big) == 1.0 and _exact_call(ref, *big) == 1.0


def test_exact_mode_replaces_monte_carlo_for_small_tables():
    cfg_mc = _cfg(P=5, chisq_permutation=1, max_iter=2000)
    cfg_ex = _cfg(P=5, chisq_permutation=1, max_iter=2000, exact_mode=1)
    batch = synth_batch(120, 5, cfg_mc.ploidy, 50, seed=31, variant_frac=0.5)
    mc = OracleEngine(cfg_mc).run(batch, mode=RNG_PHILOX)
    ex = OracleEngine(cfg_ex).run(batch, mode=RNG_PHILOX)
    exact = ex.perm_hits == -1
    assert exact.sum() > 50 and np.all(ex.perm_k[exact] == 0)
    other = ~exact
    assert np.array_equal(ex.perm_k[other], mc.perm_k[other]) and np.array_equal(ex.perm_hits[other], mc.perm_hits[other])
    # the Monte-Carlo test is on the stranded table, the exact one on the unstranded table: related, not equal
    m = exact & (mc.perm_hits >= 5)
    assert m.sum() > 50
    assert np.corrcoef(ex.ctpval[m], mc.ctpval[m])[0, 1] > 0.5


# ---- candidate screen (variantcalls.c:62-92) ------------------------------------------------
def _screen_batch(seed, P=8, ploidy=20, n=400, **kw):
    cfg = EngineConfig(ploidy=np.broadcast_to(np.asarray(ploidy), (P,)).copy(), options={})
    batch = synth_batch(n, P, cfg.ploidy, 60, seed=seed, variant_frac=0.5, indel_frac=0.3, second_alt_frac=0.4, bidir_frac=0.1, **kw)
    # sort ties and reference alleles other than the most frequent one: scramble some sites
    rng = np.random.default_rng(seed)
    n = batch.n_sites
    batch.counts = batch.counts.copy()
    c = batch.counts.reshape(-1, 24)
    tie = rng.random(n) < 0.3
    c[tie, 1] = c[tie, 2]; c[tie, 9] = c[tie, 10]; c[tie, 17] = c[tie, 18]
    batch.refbase = rng.integers(0, 4, n).astype(np.uint8)
    return cfg, batch


@needs_ref
def test_sort_allelecounts_equals_reference():
    """o_sort_allelecounts against the reference's own sort_allelecounts (linked into oracle/_ref), ties included."""
    for seed in (1, 2):
        cfg, batch = _screen_batch(seed)
        al, ct, _ = OracleEngine(cfg).screen(batch)
        ral, rct = RefEngine(cfg).sort_allelecounts(batch)
        assert np.array_equal(al, ral) and np.array_equal(ct, rct)
        assert (ct[:, 1:-1] >= ct[:, 2:]).all() and np.array_equal(al[:, 0], batch.refbase)


def test_potential_variant_score():
    """variantcalls.c:76-92 (PICALL 3): +2 per (non-reference allele, pool) with >= 3 reads; diploid pools with 1-2 reads
    add the read count.  Checked against a direct numpy evaluation."""
    for ploidy in (20, 2, [2, 20, 2, 20, 2, 20, 2, 20]):
        cfg, batch = _screen_batch(5, ploidy=ploidy)
        _, _, pot = OracleEngine(cfg).screen(batch)
        ic = batch.indcounts.reshape(batch.n_sites, 8, 3, 8).sum(axis=2)            # [site][pool][allele]
        dip = (cfg.ploidy == 2)[None, :, None]
        score = np.where(ic >= 3, 2, np.where((ic >= 1) & dip, ic, 0))
        nonref = np.arange(8)[None, None, :] != batch.refbase[:, None, None]
        assert np.array_equal(pot, (score * nonref).sum(axis=(1, 2)))
