"""Batch VCF writer (host/vcf_writer.c) against the reference's own printer.

CPU: the batches the reference's front end packs from BAM files (CRISP.gpu in dump mode) + the unmodified reference engine's
results on them, formatted by crisp_vcf_format, are byte for byte the records CRISP.binary writes for the same files —
pooled, diploid and --EM 0 output layouts, multi-allelic sites, LowMQ / LowDepth filters, IUPAC reference bases.
The GPU counterpart (device pileup + engine + this writer == CRISP.binary) is tests/test_pileup.py.
"""
import os
from types import SimpleNamespace

import numpy as np
import pytest

from crisp_b200 import frontend as fe
from crisp_b200.bamsynth import make_dataset
from crisp_b200.pileup import read_fasta
from crisp_b200.vcf import format_records
from frontend_io import read_dump
from oracle.loader import RefEngine

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
REFBIN = os.path.join(ROOT, "oracle", "_ref", "CRISP.binary")
GPUBIN = os.path.join(ROOT, "oracle", "_ref", "CRISP.gpu")

NOISY = dict(lowmapq_frac=0.15, noisy_frac=0.1, clip_frac=0.2, n_frac=0.0,  # (the replay harness cannot tell an N from an A in the packed read stream)
             qual_values=[5, 10, 12, 13, 20, 30, 40], qual_probs=[.05, .05, .05, .25, .2, .2, .2])
CASES = {
    "pooled": (20, [], dict(pools=6, length=12000, depth=60, seed=51, **NOISY)),
    "pooled_lowmq": (20, [], dict(pools=5, length=12000, depth=120, seed=52, **dict(NOISY, lowmapq_frac=0.35))),
    "diploid": (2, [], dict(pools=12, length=8000, depth=12, seed=53, **NOISY)),
    "em0": (20, ["--EM", "0"], dict(pools=6, length=16000, depth=60, seed=54, clip_frac=0.2, lowmapq_frac=0.1)),
    "shallow": (20, [], dict(pools=4, length=9000, depth=6, seed=55, spacing=150)),   # LowDepth records, pools without data
}


def sites_of(batch, ex):
    P = batch.n_pools
    return SimpleNamespace(pos=batch.pos, refbase=batch.refbase, counts=batch.counts, indcounts=batch.indcounts,
                           readdepths=ex["readdepths"].reshape(-1, P), mqcounts=ex["mqcounts"].reshape(-1, 4))


@pytest.mark.parametrize("name", list(CASES))
def test_batch_writer_prints_the_reference_records(name, tmp_path):
    if not (os.path.exists(REFBIN) and os.path.exists(GPUBIN)):
        pytest.skip("oracle/_ref programs not built (need /root/reference)")
    poolsize, extra, kw = CASES[name]
    ds = make_dataset(str(tmp_path), poolsize=poolsize, **kw)
    ref_vcf = str(tmp_path / "ref.vcf")
    fe.run_program(REFBIN, ds, ref_vcf, poolsize, extra)
    want = [line for line in open(ref_vcf, "rb").read().split(b"\n") if line and not line.startswith(b"#")]
    assert len(want) >= 3
    dump = str(tmp_path / "dump.bin")
    fe.run_program(GPUBIN, ds, str(tmp_path / "unused.vcf"), poolsize, extra, env=dict(CRISPGPU_DUMP=dump, CRISPGPU_DUMP_ONLY=1, CRISPGPU_BATCH=64))
    ref = read_fasta(ds["fasta"])
    got = []
    for cfg, batch, ex in read_dump(dump):
        res = RefEngine(cfg).run(batch, seed=5)
        text = format_records(sites_of(batch, ex), res, cfg.ploidy, ds["chrom"], ref, chrom_len=ds["length"], em_flag=cfg.options["em_flag"],
                              varpoolsize=cfg.options["varpoolsize"], threads=3)
        got += [line for line in text.split(b"\n") if line]
    assert len(got) == len(want), f"{len(got)} records, CRISP.binary wrote {len(want)}"
    for g, w in zip(got, want):
        assert g == w, f"record differs:\n{g.decode()}\n{w.decode()}"
    if name == "pooled_lowmq":
        assert any(b"LowMQ" in w for w in want)
    if name == "shallow":
        assert any(b"LowDepth" in w for w in want)


def test_batch_writer_threads_and_empty_batches(tmp_path):
    """The text does not depend on the thread count; a batch without called sites gives no text."""
    if not os.path.exists(GPUBIN):
        pytest.skip("oracle/_ref programs not built (need /root/reference)")
    poolsize, extra, kw = CASES["pooled"]
    ds = make_dataset(str(tmp_path), poolsize=poolsize, **kw)
    dump = str(tmp_path / "dump.bin")
    fe.run_program(GPUBIN, ds, str(tmp_path / "unused.vcf"), poolsize, extra, env=dict(CRISPGPU_DUMP=dump, CRISPGPU_DUMP_ONLY=1, CRISPGPU_BATCH=10 ** 6))
    (cfg, batch, ex), = read_dump(dump)
    res = RefEngine(cfg).run(batch, seed=5)
    ref = read_fasta(ds["fasta"])
    texts = [format_records(sites_of(batch, ex), res, cfg.ploidy, ds["chrom"], ref, chrom_len=ds["length"], threads=t) for t in (1, 2, 7, 64)]
    assert all(t == texts[0] for t in texts) and texts[0].count(b"\n") == int((res.varalleles >= 1).sum()) > 0
    res.varalleles[:] = 0
    assert format_records(sites_of(batch, ex), res, cfg.ploidy, ds["chrom"], ref, chrom_len=ds["length"]) == b""
