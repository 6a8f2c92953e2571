"""Ad-hoc GPU parity probe (development aid): engine vs oracle on several scenarios, printing every difference."""
import sys, os, time, traceback
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, os.path.join(os.path.dirname(os.path.dirname(os.path.abspath(__file__))), "tests"))
import numpy as np
from crisp_b200.abi import EngineConfig
from crisp_b200.engine import Engine
from crisp_b200.synth import synth_batch, survey_demo_batch
from oracle.loader import OracleEngine, RNG_PHILOX, RNG_DRAND48, build
from util import compare_results

build()
SCEN = [
    ("demo", None, 6, 20, 100, {}, {}),
    ("default", 300, 10, 20, 100, {}, dict(variant_frac=0.4, bidir_frac=0.1, second_alt_frac=0.3)),
    ("c2small", 200, 50, 48, 200, {}, dict(variant_frac=0.3)),
    ("varpool", 300, 8, [10, 20, 30, 40, 10, 20, 30, 40], 120, {}, dict(variant_frac=0.4, bidir_frac=0.1)),
    ("pivot", 300, 12, 16, 80, dict(pivot_sample=5), dict(variant_frac=0.4)),
    ("indel", 300, 10, 20, 100, {}, dict(variant_frac=0.3, indel_frac=0.5, bidir_frac=0.05)),
    ("nobq", 200, 10, 20, 100, dict(use_base_qvs=0), dict(variant_frac=0.4)),
    ("refbias", 200, 10, 20, 100, dict(agilent_ref_bias=0.55), dict(variant_frac=0.4)),
    ("perm", 150, 10, 20, 100, dict(chisq_permutation=1, max_iter=2000), dict(variant_frac=0.4, bidir_frac=0.05, with_qhigh=True)),
    ("lowcov", 300, 40, 2, 10, dict(chisq_permutation=1, max_iter=2000), dict(variant_frac=0.5, bidir_frac=0.1)),
    ("diploid", 300, 40, 2, 10, {}, dict(variant_frac=0.5)),
    ("em0", 300, 10, 20, 100, dict(em_flag=0), dict(variant_frac=0.4, indel_frac=0.2, with_stats=True, with_qhigh=True)),
    ("manyquals", 200, 10, 20, 100, {}, dict(variant_frac=0.4, bidir_frac=0.1, quals=range(14, 42))),
    ("bidir_dense", 200, 10, 20, 100, {}, dict(variant_frac=0.4, bidir_frac=0.2)),
    ("c3small", 60, 100, 96, 60, {}, dict(variant_frac=0.5)),
    ("c5small", 60, 500, 10, 10, {}, dict(variant_frac=0.5)),
]
only = sys.argv[1:]
SCALE = float(os.environ.get("GPU_CHECK_SCALE", "1"))  # shrink the scenarios, e.g. under compute-sanitizer
fails = 0
for name, n, P, ploidy, depth, opts, kw in SCEN:
    if only and name not in only:
        continue
    try:
        cfg = EngineConfig(ploidy=np.broadcast_to(np.asarray(ploidy), (P,)).copy(), options=opts)
        b = survey_demo_batch() if n is None else synth_batch(max(8, int(n * SCALE)), P, cfg.ploidy, depth, seed=11, **kw)
        t = time.time(); want = OracleEngine(cfg).run(b, mode=RNG_PHILOX); to = time.time() - t
        eng = Engine(cfg)
        t = time.time(); got = eng.run(b); tg = time.time() - t
        tm = eng.timing()
        print(f"[{name}] sites={b.n_sites} oracle {to:.2f}s gpu {tg:.3f}s status={np.bincount(got.status.ravel(), minlength=5).tolist()} want={np.bincount(want.status.ravel(), minlength=5).tolist()} timing={ {k: round(v,3) if isinstance(v,float) else v for k,v in tm.items()} }")
        skip = ()
        try:
            flips = compare_results(got, want, label=name, allow_flips=0)
            if opts.get("chisq_permutation"):
                assert np.array_equal(got.perm_k, want.perm_k) and np.array_equal(got.perm_hits, want.perm_hits), "perm k/hits differ"
            print(f"   OK")
        except AssertionError as e:
            fails += 1
            print("   FAIL:", str(e)[:3000])
        eng.close()
    except Exception:
        fails += 1
        traceback.print_exc()
print("FAILS", fails)
