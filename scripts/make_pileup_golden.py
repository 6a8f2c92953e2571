"""Generate tests/golden/pileup/frontend.npz: a small multi-pool alignment set and what the REFERENCE's own front end
(oracle/_ref/CRISP.gpu in dump mode = the unmodified parsebam/ objects up to the call at variantcalls.c:115) makes of it —
candidate positions, allele counts, sorted allele vectors, read depths, mapping-quality classes, filtered-read counters and
the packed read / alt-read streams.  Run here (needs /root/reference for oracle/_ref); the fixture travels to the GPU box.

    python scripts/make_pileup_golden.py
"""
import os
import sys
import tempfile

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "tests"))

from crisp_b200 import bamsynth, frontend as fe  # noqa: E402
from crisp_b200.pileup import decode_bams, read_fasta  # noqa: E402
from frontend_io import read_dump  # noqa: E402

POOLS, POOLSIZE = 4, 12
KW = dict(pools=POOLS, poolsize=POOLSIZE, length=5000, depth=50, seed=33, spacing=300, lowmapq_frac=0.12, noisy_frac=0.08, clip_frac=0.2,
          n_frac=0.004, qual_values=[5, 10, 12, 13, 20, 30, 40], qual_probs=[.04, .04, .04, .38, .2, .15, .15])


def main():
    d = tempfile.mkdtemp()
    ds = bamsynth.make_dataset(d, **KW)
    dump = os.path.join(d, "dump.bin")
    fe.run_program(os.path.join(ROOT, "oracle", "_ref", "CRISP.gpu"), ds, os.path.join(d, "unused.vcf"), POOLSIZE, [],
                   env=dict(CRISPGPU_DUMP=dump, CRISPGPU_DUMP_ONLY=1, CRISPGPU_BATCH=1000000))
    (cfg, batch, ex), = read_dump(dump)
    aln = decode_bams(ds["bams"], pinned=False, threads=2)
    out = dict(n_pools=POOLS, poolsize=POOLSIZE, length=ds["length"], ref=read_fasta(ds["fasta"]), pool_off=aln.pool_off, rec=aln.rec,
               seq4=aln.seq4, qual=aln.qual)
    for k in ("pos", "refbase", "maxvec_al", "maxvec_ct", "counts", "indcounts", "read_off", "reads", "alt_off", "alt_reads"):
        out["want_" + k] = np.asarray(getattr(batch, k))
    for k in ("readdepths", "mqcounts", "filtered"):
        out["want_" + k] = ex[k]
    path = os.path.join(ROOT, "tests", "golden", "pileup", "frontend.npz")
    np.savez_compressed(path, **out)
    print(path, os.path.getsize(path), "bytes;", batch.n_sites, "candidate sites,", aln.n_reads, "alignments")


if __name__ == "__main__":
    main()
