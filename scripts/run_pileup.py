"""Run the device pileup on a synthetic multi-pool BAM set (profiling / quick measurements).

    python scripts/run_pileup.py --pools 10 --poolsize 20 --length 100000 --depth 100 --steps 5
"""
import argparse
import os
import sys
import tempfile
import time

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)

from crisp_b200.abi import EngineConfig  # noqa: E402
from crisp_b200.bamsynth import make_dataset  # noqa: E402
from crisp_b200.engine import Engine  # noqa: E402
from crisp_b200.pileup import decode_bams, make_window, read_fasta  # noqa: E402


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--pools", type=int, default=10)
    ap.add_argument("--poolsize", type=int, default=20)
    ap.add_argument("--length", type=int, default=100000)
    ap.add_argument("--depth", type=float, default=100)
    ap.add_argument("--steps", type=int, default=5)
    ap.add_argument("--threads", type=int, default=os.cpu_count() or 1)
    a = ap.parse_args()
    with tempfile.TemporaryDirectory() as d:
        t0 = time.perf_counter()
        ds = make_dataset(d, pools=a.pools, poolsize=a.poolsize, length=a.length, depth=a.depth, seed=1)
        t1 = time.perf_counter()
        aln = decode_bams(ds["bams"], threads=a.threads)
        t2 = time.perf_counter()
        ref = read_fasta(ds["fasta"])
    print(f"dataset {t1 - t0:.1f} s, decode {t2 - t1:.3f} s ({aln.stats['compressed_bytes'] / (t2 - t1) / 1e6:.0f} MB/s compressed, {aln.n_reads / (t2 - t1) / 1e6:.2f} M alignments/s), "
          f"{aln.n_reads} alignments, {aln.n_bases} bases")
    eng = Engine(EngineConfig(ploidy=np.full(a.pools, a.poolsize), options=dict(minimal_results=1)), device=0)
    win = make_window(0, 0, a.length, ref)
    n, res = eng.pileup_run(aln, win, want_sites=False)
    t = eng.timing()
    print(f"pileup_run: {n} candidate sites, {int((res.varalleles >= 1).sum())} called; h2d {t['h2d_ms']:.3f} ms pileup {t['pileup_ms']:.3f} ms (tile {t['pile_tile_ms']:.3f}) "
          f"evaluate {t['evaluate_ms']:.3f} em {t['em_ms']:.3f} total {t['total_ms']:.3f} ms")
    eng.pileup_upload(aln, win)
    for _ in range(a.steps):
        eng.pileup_run_resident(False)
        t = eng.timing()
        print(f"resident: pileup {t['pileup_ms']:.3f} ms, tile {t['pile_tile_ms']:.3f} ms = {t['pile_bytes'] / t['pile_tile_ms'] / 1e6:.0f} GB/s of {t['pile_bytes'] / 1e6:.1f} MB, total {t['total_ms']:.3f} ms")
    eng.close()


if __name__ == "__main__":
    main()
