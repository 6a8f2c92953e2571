"""Drive the reference's own program on BAM files: `CRISP.binary` (unmodified reference) and `CRISP.gpu` (the reference's
front end whose one call into the statistics engine, variantcalls.c:115, is bound to libcrispgpu.so by
host/crisp_frontend.c).  Both are built by oracle/Makefile from /root/reference and travel prebuilt to the GPU box; the
callers (tests/, bench.py) pass their paths in.  Region sharding is the reference's own mechanism: one process per
contiguous `--regions chr:start-end` slice of the indexed BAMs, VCF bodies concatenated in coordinate order
(README.md:51, bamsreader.c:147-173).
"""
from __future__ import annotations

import json
import os
import re
import subprocess
import time
from concurrent.futures import ThreadPoolExecutor


def crisp_args(ds: dict, vcf: str, poolsize: int, extra=(), region: str | None = None) -> list:
    args = ["--bams", ds["bamlist"], "--ref", ds["fasta"], "--VCF", vcf, "--poolsize", str(poolsize)]
    if region:
        args += ["--regions", region]
    return args + [str(x) for x in extra]


def run_program(binary: str, ds: dict, vcf: str, poolsize: int, extra=(), region: str | None = None, env: dict | None = None,
                timeout: float = 3600.0) -> float:
    """One process; returns its wall time in seconds.  The reference's main() always returns 1 (readmultiplebams.c:278)."""
    e = dict(os.environ)
    if env:
        e.update({k: str(v) for k, v in env.items()})
    t0 = time.perf_counter()
    p = subprocess.run([binary] + crisp_args(ds, vcf, poolsize, extra, region), stdout=subprocess.DEVNULL, stderr=subprocess.PIPE,
                       env=e, timeout=timeout)
    dt = time.perf_counter() - t0
    if p.returncode not in (0, 1):
        raise RuntimeError(f"{os.path.basename(binary)} exited {p.returncode}: {p.stderr.decode(errors='replace')[-800:]}")
    return dt


def index_bams(bamtool: str, ds: dict) -> None:
    for b in ds["bams"]:
        if not os.path.exists(b + ".bai"):
            subprocess.run([bamtool, "index", b], check=True)


def regions(ds: dict, n: int) -> list:
    """n equal contiguous slices of the (single-chromosome) synthetic reference, 1-based inclusive as samtools parses them."""
    L = ds["length"]
    edges = [round(L * k / n) for k in range(n + 1)]
    return [f"{ds['chrom']}:{edges[k] + 1}-{edges[k + 1]}" for k in range(n)]


def run_sharded(binary: str, ds: dict, vcf: str, poolsize: int, nproc: int, extra=(), env: dict | None = None,
                per_shard_env=None) -> dict:
    """nproc concurrent `--regions` shards (needs .bai).  Wall time = until the slowest shard ends.  The shard VCFs are
    concatenated (header of the first, bodies in region order) into `vcf`."""
    regs = regions(ds, nproc)
    parts = [f"{vcf}.shard{k}" for k in range(nproc)]

    def one(k):
        e = dict(env or {})
        if per_shard_env:
            e.update(per_shard_env(k))
        return run_program(binary, ds, parts[k], poolsize, extra, regs[k], e)

    t0 = time.perf_counter()
    with ThreadPoolExecutor(nproc) as ex:
        times = list(ex.map(one, range(nproc)))
    wall = time.perf_counter() - t0
    with open(vcf, "w") as out:
        for k, p in enumerate(parts):
            with open(p) as f:
                for line in f:
                    if line.startswith("#") and k > 0:
                        continue
                    out.write(line)
            os.remove(p)
    return dict(wall_s=wall, shard_s=times, nproc=nproc)


def read_vcf(path: str) -> list:
    with open(path) as f:
        return [line.rstrip("\n").split("\t") for line in f if line.strip() and not line.startswith("#")]


_CT = re.compile(r"CT=([^;]*)")


def compare_vcf(got: list, want: list, monte_carlo: bool = False) -> dict:
    """Record-by-record comparison of two VCF bodies keyed by (chrom, pos, ref).  Returns
    {same, only_got, only_want, differ: [(key, column, got, want)]}.  With monte_carlo=True the CT= field (permutation
    p-value: the reference draws from drand48 seeded with the time of day, readmultiplebams.c:252) is compared
    separately: `ct` lists (key, got, want) and everything else must match."""
    def key(r):
        return (r[0], int(r[1]), r[3])

    g = {key(r): r for r in got}
    w = {key(r): r for r in want}
    res = dict(same=0, only_got=sorted(set(g) - set(w)), only_want=sorted(set(w) - set(g)), differ=[], ct=[])
    for k in sorted(set(g) & set(w)):
        a, b = list(g[k]), list(w[k])
        if monte_carlo:
            ca, cb = _CT.search(a[7]), _CT.search(b[7])
            res["ct"].append((k, ca.group(1) if ca else None, cb.group(1) if cb else None))
            a[7], b[7] = _CT.sub("CT=", a[7]), _CT.sub("CT=", b[7])
        if a == b:
            res["same"] += 1
            continue
        for c, (x, y) in enumerate(zip(a, b)):
            if x != y:
                res["differ"].append((k, c, x, y))
                break
        else:
            res["differ"].append((k, -1, len(a), len(b)))
    return res


def read_stats(path: str) -> dict:
    with open(path) as f:
        return json.loads(f.read())
