"""Generate tests/golden/*.npz from the UNMODIFIED reference (oracle/_ref, built from /root/reference).

Run in the build container (where /root/reference exists):  python scripts/make_golden.py
Each fixture holds a small seeded batch, the engine options, and what the reference itself computed on it
(per-slot statistics, per-pool allele counts / qualities, and the VCF text of print_pooledvariant), so that the
oracle restatement and the CUDA engine can be pinned where /root/reference is absent.
"""
import os
import sys

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
from crisp_b200.abi import _BATCH_ARRAYS, _RESULT_FIELDS, EngineConfig  # noqa: E402
from crisp_b200.synth import survey_demo_batch, synth_batch  # noqa: E402
from oracle.loader import RefEngine, build  # noqa: E402

SCENARIOS = {
    # name: (n_sites, P, ploidy, depth, options, synth kwargs, drand48 seed)
    "survey_demo": (None, 6, 20, 100, {}, {}, 1),
    "pooled_default": (60, 8, 20, 80, {}, dict(variant_frac=0.5, bidir_frac=0.1, second_alt_frac=0.3), 1),
    "pooled_indel": (60, 8, 20, 80, {}, dict(variant_frac=0.3, indel_frac=0.5, bidir_frac=0.05), 1),
    "pooled_perm_drand48": (40, 8, 20, 80, dict(chisq_permutation=1, max_iter=2000), dict(variant_frac=0.5, with_qhigh=True), 4242),
    "diploid_lowcov_drand48": (60, 24, 2, 10, dict(chisq_permutation=1, max_iter=2000), dict(variant_frac=0.5, bidir_frac=0.1), 4242),
    "legacy_em0": (60, 8, 20, 80, dict(em_flag=0), dict(variant_frac=0.5, indel_frac=0.2, with_stats=True, with_qhigh=True), 1),
    "variable_poolsize": (60, 6, [10, 16, 24, 10, 16, 24], 90, {}, dict(variant_frac=0.5, bidir_frac=0.1), 1),
    "pivot_sample": (60, 10, 16, 70, dict(pivot_sample=4), dict(variant_frac=0.5), 1),
    "exome_c2_shape": (12, 50, 48, 200, {}, dict(variant_frac=0.5), 1),
}


def main():
    build()
    out_dir = os.path.join(ROOT, "tests", "golden")
    os.makedirs(out_dir, exist_ok=True)
    for name, (n, P, ploidy, depth, opts, kw, seed) in SCENARIOS.items():
        cfg = EngineConfig(ploidy=np.broadcast_to(np.asarray(ploidy), (P,)).copy(), options=opts)
        batch = survey_demo_batch() if n is None else synth_batch(n, P, cfg.ploidy, depth, seed=20240 + len(name), **kw)
        ref = RefEngine(cfg)
        res = ref.run(batch, seed=seed, want_vcf=True)
        data = {"ploidy": cfg.ploidy, "drand48_seed": np.int64(seed)}
        data["option_names"] = np.array(sorted(opts), dtype="U32")
        data["option_values"] = np.array([float(opts[k]) for k in sorted(opts)], dtype=np.float64)
        for an, _, _ in _BATCH_ARRAYS:
            a = getattr(batch, an)
            if a is not None:
                data["b_" + an] = a
        for rn, _, _ in _RESULT_FIELDS:
            data["r_" + rn] = getattr(res, rn)
        data["r_ac"] = res.ac
        data["r_qv"] = res.qv
        data["vcf"] = np.array(res.vcf)
        path = os.path.join(out_dir, name + ".npz")
        np.savez_compressed(path, **data)
        print(f"{name}: sites={batch.n_sites} calls={res.n_calls} vcf_lines={len(res.vcf.splitlines())} {os.path.getsize(path)/1024:.0f} KiB")


if __name__ == "__main__":
    main()
