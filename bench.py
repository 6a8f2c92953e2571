#!/usr/bin/env python
"""bench.py — candidate sites/sec of the CRISP per-site statistics engine on B200 (BASELINE.json metric).

  python bench.py --gpus N --steps K --warmup W            (N>1: launched under torch.distributed.run)
  python bench.py --impl reference --gpus N --steps K --warmup W   the reference's own CPU code on the host cores

A "step" is one pass of the hot path (k_evaluate -> k_em -> k_finalize; + permutation kernels when the workload
enables them) over one batch of synthetic candidate sites of the named shape.  At N GPUs the sites are sharded by
genomic region, one shard per rank, no data-path collective: weak scaling, each rank owns a batch of
`--sites` candidates.  `value` is device-timed with the inputs resident in HBM; `e2e` is the same metric through the
C-ABI call a front end makes (crispgpu_submit/crispgpu_wait) with pinned host buffers, host<->device copies inside
the timed region.  Prints ONE JSON line on rank 0.
"""
from __future__ import annotations

import argparse
import json
import os
import subprocess
import sys
import threading
import time

import numpy as np

ROOT = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, ROOT)

WORKLOADS = {
    # BASELINE.json configs[1]: 50 pools x poolsize 48, synthetic exome at 200x per pool (default flags: --EM 1,
    # asymptotic contingency test).  variant_frac: planted variants every 500 bp vs ~1.3 % of positions being
    # sequencing-error candidates at this depth -> ~15 % of candidates carry a variant.
    "exome50x48": dict(P=50, ploidy=48, depth=200, options={}, synth=dict(variant_frac=0.15)),
    "c1_10x20_perm": dict(P=10, ploidy=20, depth=100, options=dict(chisq_permutation=1, max_iter=20000), synth=dict(variant_frac=0.3)),
    "c3_100x96": dict(P=100, ploidy=96, depth=60, options={}, synth=dict(variant_frac=0.15)),
    "c4_perm_stress": dict(P=50, ploidy=48, depth=200, options=dict(chisq_permutation=1, max_iter=1000000), synth=dict(variant_frac=0.8, rare=True)),
    "c5_500x10": dict(P=500, ploidy=10, depth=10, options={}, synth=dict(variant_frac=0.15)),
    # config 5 with the permutation test on: pool size 10 dispatches to the stranded test, pool size 2 to lowcovFET
    "c5_500x10_perm": dict(P=500, ploidy=10, depth=10, options=dict(chisq_permutation=1, max_iter=20000), synth=dict(variant_frac=0.3)),
    "c5p_500x2_lowcov": dict(P=500, ploidy=2, depth=10, options=dict(chisq_permutation=1, max_iter=20000), synth=dict(variant_frac=0.5)),
    # config 4 exactly as BASELINE.json words it: --perms 1,000,000 with --ctpval -6; --ctpval is consumed only by the
    # --EM 0 caller (crisp/oldcrispmethod.c:170), so this is that path: em_flag=0, thresh1=-6
    "c4_perm_stress_em0": dict(P=50, ploidy=48, depth=200, options=dict(chisq_permutation=1, max_iter=1000000, em_flag=0, thresh1=-6.0),
                               synth=dict(variant_frac=0.8, rare=True, with_stats=True, with_qhigh=True)),
    # config 2 with a 40-valued quality spectrum (older instruments): the likelihood takes the sparse path, bins barely compress
    "exome50x48_q40": dict(P=50, ploidy=48, depth=200, options={}, synth=dict(variant_frac=0.15, quals=range(13, 53))),
}

# the other BASELINE configs, measured at reduced size inside the default run: (workload, sites, reference sites per core)
SIDE_CONFIGS = [
    ("c1_10x20_perm", 1024, 24), ("c3_100x96", 4096, 16), ("c4_perm_stress", 1024, 1), ("c4_perm_stress_em0", 1024, 1),
    ("c5_500x10", 1024, 16), ("c5_500x10_perm", 512, 2), ("c5p_500x2_lowcov", 512, 2), ("exome50x48_q40", 2048, 64),
]


def make_batch(wl: dict, n_sites: int, seed: int, tid: int = 0):
    """Seeded synthetic candidate sites of the workload's shape, generated in chunks to bound host memory."""
    from crisp_b200.abi import concat_batches
    from crisp_b200.synth import synth_batch

    parts, have, chunk, k = [], 0, 1024, 0
    while have < n_sites:
        b = synth_batch(chunk, wl["P"], wl["ploidy"], wl["depth"], seed=seed * 1000 + k, tid=tid, pos_start=1000 + k * chunk * 75, **wl["synth"])
        parts.append(b)
        have += b.n_sites
        k += 1
    full = concat_batches(parts)
    return full.select(np.arange(n_sites)) if full.n_sites > n_sites else full


class ClockSampler:
    """nvidia-smi clocks / throttle reasons sampled during the timed region (B200_PROFILING.md recipe)."""

    def __init__(self, index: int):
        self.index, self.rows, self.proc = index, [], None

    def start(self):
        q = "clocks.sm,clocks.max.sm,power.draw,clocks_event_reasons.hw_slowdown,clocks_event_reasons.hw_thermal_slowdown,clocks_event_reasons.sw_thermal_slowdown,clocks_event_reasons.sw_power_cap"
        try:
            self.proc = subprocess.Popen(["nvidia-smi", "-i", str(self.index), f"--query-gpu={q}", "--format=csv,noheader,nounits", "-lms", "100"],
                                         stdout=subprocess.PIPE, stderr=subprocess.DEVNULL, text=True)
            self.t = threading.Thread(target=self._read, daemon=True)
            self.t.start()
        except Exception:
            self.proc = None

    def _read(self):
        for line in self.proc.stdout:
            self.rows.append([x.strip() for x in line.split(",")])

    def stop(self) -> dict:
        if self.proc is None:
            return {"sm_mhz": None, "sm_max_mhz": None, "reasons": ["nvidia-smi unavailable"]}
        self.proc.terminate()
        try:
            self.proc.wait(timeout=5)
        except Exception:
            self.proc.kill()
        sm, mx, reasons = [], [], set()
        for r in self.rows:
            try:
                sm.append(float(r[0])); mx.append(float(r[1]))
            except Exception:
                continue
            for name, v in zip(("hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown", "sw_power_cap"), r[3:7]):
                if v.lower().startswith("active"):
                    reasons.add(name)
        return {"sm_mhz": float(np.median(sm)) if sm else None, "sm_max_mhz": max(mx) if mx else None, "reasons": sorted(reasons), "samples": len(sm)}


def _ref_worker(args):
    cfg_ploidy, options, arrays, P, seed = args
    from crisp_b200.abi import Batch, EngineConfig
    from oracle.loader import OracleEngine, RefEngine, have_ref
    cfg = EngineConfig(ploidy=cfg_ploidy, options=options)
    batch = Batch(P, **arrays)
    t0 = time.perf_counter()
    if have_ref():
        RefEngine(cfg).run(batch, seed=seed)
    else:
        OracleEngine(cfg).run(batch, mode=0, seed=seed)
    return time.perf_counter() - t0, batch.n_sites


def run_reference_sample(wl: dict, batch, cores: int, sites_per_core: int):
    """The reference's own CPU implementation (oracle/_ref; oracle port if it was never built) on `cores` processes,
    each owning a contiguous region of the sample — the reference's --regions parallelism."""
    import multiprocessing as mp
    from crisp_b200.abi import _BATCH_ARRAYS
    from oracle.loader import have_ref

    n = min(batch.n_sites, cores * sites_per_core)
    per = n // cores
    jobs = []
    ploidy = np.full(wl["P"], wl["ploidy"], np.int32)
    for c in range(cores):
        sub = batch.select(np.arange(c * per, (c + 1) * per))
        arrays = {name: getattr(sub, name) for name, _, _ in _BATCH_ARRAYS if getattr(sub, name) is not None}
        jobs.append((ploidy, wl["options"], arrays, wl["P"], 1 + c))
    ctx = mp.get_context("fork")
    t0 = time.perf_counter()
    with ctx.Pool(cores) as pool:
        out = pool.map(_ref_worker, jobs)
    wall = time.perf_counter() - t0
    busy = max(o[0] for o in out)
    sites = sum(o[1] for o in out)
    return sites / busy, sites, busy, wall, ("reference" if have_ref() else "port")


def host_cores() -> int:
    try:
        return len(os.sched_getaffinity(0))
    except Exception:
        return os.cpu_count() or 1


def emit(line: dict) -> None:
    """Write the one JSON line to the real stdout (everything else — NCCL banners, make, warnings — goes to stderr)."""
    os.write(_REAL_STDOUT, (json.dumps(line) + "\n").encode())


_REAL_STDOUT = 1


def perm_placements(batch, results, lowcov: bool) -> float:
    """Ball placements the permutation kernels performed: sum over tested (site, slot) of executed permutations x the
    alternate reads of the slot's allele (forward + reverse; + bidirectional in the low-coverage test) — SURVEY.md §8d K3."""
    k = results.perm_k.astype(np.int64)
    s_idx, slot = np.nonzero(k > 0)
    if s_idx.size == 0:
        return 0.0
    P = batch.n_pools
    ic = batch.indcounts.reshape(batch.n_sites, P, 24).sum(axis=1)
    al = results.allele[s_idx, slot].astype(np.int64)
    ones = ic[s_idx, al] + ic[s_idx, al + 8] + (ic[s_idx, al + 16] if lowcov else 0)
    return float((k[s_idx, slot] * ones).sum())


def rooflines(name, wl, batch, shipped, results, kt, steps, n_em, fmt, hbm_peak, hbm_src, fp64_peak, int32_peak, traffic=None, ncu_static=None):
    """Roofline objects of the three kernel groups of one workload; `dominant` names the largest share of the step."""
    traffic, ncu_static = traffic or {}, ncu_static or {}
    P, J = wl["P"], wl["ploidy"] + 1
    em_ms, ev_ms, pm_ms = kt["em_ms"] / steps, kt["evaluate_ms"] / steps, kt["permute_ms"] / steps
    iters = results.em_iters[results.em_iters > 0]
    # algorithmic work of the EM kernel: per iteration 3 planes x sum_i (n_i+1) genotype evaluations at
    # 35 flop + 3 P logs at 30 flop
    ge_flops = float(iters.sum()) * (35.0 * 3 * P * J + 30.0 * 3 * P)
    em_sites = np.unique(np.nonzero(results.em_iters > 0)[0])
    if fmt == "bins" and shipped.bin_off is not None:   # the likelihood reads the sites' (value,count) bins, 4 B each
        bo = shipped.bin_off.astype(np.int64)
        em_bytes = float(sum(4 * (bo[(s + 1) * P] - bo[s * P]) for s in em_sites)) + n_em * (96.0 * P + 12.0 * P + 128.0)
    else:
        ro = batch.read_off.astype(np.int64)
        em_bytes = float(sum(2 * (ro[(s + 1) * P] - ro[s * P]) for s in em_sites)) + n_em * (96.0 * P + 12.0 * P + 128.0)
    ev_bytes = batch.n_sites * (96.0 * P + 40.0 + 5 * 110.0)
    out = {}
    out["k_em"] = {"kernel": "k_em", "bound": "fp64", "achieved": ge_flops / (em_ms * 1e-3) / 1e12 if em_ms > 0 else None, "peak": fp64_peak,
                   "unit": "TFLOP/s", "frac": (ge_flops / (em_ms * 1e-3) / 1e12 / fp64_peak) if em_ms > 0 and fp64_peak > 0 else None,
                   "traffic": traffic.get("k_em"), "algorithmic_bytes": em_bytes, "peak_source": "DFMA loop measured live (crispgpu_measure_fp64_tflops)",
                   "ms_per_launch": em_ms, "tasks_per_launch": n_em, "mean_em_iterations": float(iters.mean()) if iters.size else 0.0,
                   "algorithmic_hbm_GBps": em_bytes / (em_ms * 1e-3) / 1e9 if em_ms > 0 else None,
                   "ncu_committed_capture": ncu_static.get("k_em") or None}
    out["k_evaluate"] = {"kernel": "k_evaluate", "bound": "hbm", "achieved": ev_bytes / (ev_ms * 1e-3) / 1e9 if ev_ms > 0 else None, "peak": hbm_peak, "unit": "GB/s",
                         "frac": (ev_bytes / (ev_ms * 1e-3) / 1e9 / hbm_peak) if ev_ms > 0 else None, "traffic": traffic.get("k_evaluate"), "algorithmic_bytes": ev_bytes,
                         "peak_source": hbm_src, "ms_per_launch": ev_ms, "sites_per_launch": batch.n_sites, "ncu_committed_capture": ncu_static.get("k_evaluate") or None}
    if wl["options"].get("chisq_permutation") and pm_ms > 0:
        lowcov = wl["ploidy"] <= 2
        placements = perm_placements(batch, results, lowcov)
        # algorithmic INT32 work per ball placement: one Philox4x32-10 draw amortised over its four words
        # (10 rounds x 4 multiplies / 4 = 10) + the bounded-integer reduction (2) + the search of the pool among P cumulative
        # capacities (ceil log2 P) + the occupancy test and update (4)
        ops = 16 + int(np.ceil(np.log2(max(P, 2))))
        ach = placements * ops / (pm_ms * 1e-3) / 1e12
        out["k_permute"] = {"kernel": "k_permute_lowcov(+tail)" if lowcov else "k_permute_stranded(+tail)", "bound": "int32", "achieved": ach, "peak": int32_peak,
                            "unit": "TIOP/s", "frac": ach / int32_peak if int32_peak > 0 else None, "traffic": None,
                            "peak_source": "IMAD loop measured live (crispgpu_measure_int32_tops)", "ms_per_launch": pm_ms,
                            "placements_per_launch": placements, "int32_ops_per_placement": ops,
                            "placements_per_s": placements / (pm_ms * 1e-3), "permutations_per_launch": float(results.perm_k[results.perm_k > 0].sum()),
                            "ncu_committed_capture": ncu_static.get("k_permute") or None}
    share = {"k_em": em_ms, "k_evaluate": ev_ms, "k_permute": pm_ms if "k_permute" in out else 0.0}
    dominant = max(share, key=share.get)
    return out, dominant


def time_resident(eng, shipped, steps, warmup, stream, flush, barrier):
    """K resident steps, CUDA events per step on the engine's stream (L2 flushed in between, outside the events)."""
    import torch
    eng.upload(shipped)
    for _ in range(warmup):
        eng.run_resident(fetch=False)
    barrier()
    evs = [(torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)) for _ in range(steps)]
    kt = {"evaluate_ms": 0.0, "em_ms": 0.0, "permute_ms": 0.0, "filter_ms": 0.0, "legacy_ms": 0.0}
    launches = 0
    with torch.cuda.stream(stream):
        for e0, e1 in evs:
            if flush is not None:
                flush.fill_(1)
            e0.record()
            eng.run_resident(fetch=False)
            e1.record()
            t = eng.timing()
            for k in kt:
                kt[k] += t[k]
            launches += t["launches"]
    barrier()
    ms = sum(e0.elapsed_time(e1) for e0, e1 in evs)
    return ms, kt, launches, eng.timing()


def time_e2e(cfg_e2e, pinned, steps, warmup, nctx, local, barrier):
    """K batches through crispgpu_submit / crispgpu_wait from pinned host memory: one host thread per engine context, steps
    dealt round-robin; wall clock between two barriers.  Returns (seconds, timing of one context's last batch)."""
    import torch
    from crisp_b200.engine import Engine
    nctx = max(1, min(nctx, steps))
    streams = [torch.cuda.Stream(device=local) for _ in range(nctx)]
    engs = [Engine(cfg_e2e, device=local, stream=st.cuda_stream) for st in streams]
    for e in engs:
        for _ in range(max(1, warmup // 2)):
            e.run(pinned)
    share = [steps // nctx + (1 if k < steps % nctx else 0) for k in range(nctx)]
    gate = threading.Barrier(nctx + 1)
    errors = []

    def front_end(e, nbatches):
        try:
            torch.cuda.set_device(local)
            gate.wait()
            for _ in range(nbatches):
                e.wait(e.submit(pinned), copy=False)
        except Exception as exc:  # noqa: BLE001
            errors.append(exc)

    workers = [threading.Thread(target=front_end, args=(e, nb)) for e, nb in zip(engs, share)]
    for w in workers:
        w.start()
    barrier()
    gate.wait()
    w0 = time.perf_counter()
    for w in workers:
        w.join()
    barrier()
    secs = time.perf_counter() - w0
    if errors:
        raise errors[0]
    t = engs[0].timing()
    for e in engs:
        e.close()
    return secs, t, nctx


def side_config(name, sites, ref_per_core, local, peaks3, cores, want_cpu):
    """One of the other BASELINE configs at reduced size: resident and end-to-end rate, roofline of the kernel that dominates
    its step, and the reference on the host cores over a bounded sample."""
    import torch
    from crisp_b200.abi import EngineConfig
    from crisp_b200.engine import Engine, pin_batch
    wl = WORKLOADS[name]
    hbm_peak, hbm_src, fp64_peak, int32_peak = peaks3
    steps, warmup = 3, 3
    t0 = time.perf_counter()
    batch = make_batch(wl, sites, seed=11)
    cfg = EngineConfig(ploidy=np.full(wl["P"], wl["ploidy"]), options=wl["options"])
    shipped = batch.compact()
    pinned = pin_batch(shipped)
    stream = torch.cuda.Stream(device=local)
    eng = Engine(cfg, device=local, stream=stream.cuda_stream)
    flush = torch.empty(256 << 20, dtype=torch.uint8, device=f"cuda:{local}")
    ms, kt, launches, t_last = time_resident(eng, shipped, steps, warmup, stream, flush, torch.cuda.synchronize)
    results = eng.run(batch)
    eng.close()
    secs, t_e2e, nctx = time_e2e(cfg, pinned, steps, warmup, 2, local, torch.cuda.synchronize)
    roofs, dominant = rooflines(name, wl, batch, shipped, results, kt, steps, t_last["n_tasks_em"], "bins", hbm_peak, hbm_src, fp64_peak, int32_peak)
    out = {"workload": f"{name}: {wl['P']} pools x poolsize {wl['ploidy']}, {wl['depth']}x per pool, options {wl['options'] or 'defaults'}",
           "sites": batch.n_sites, "steps": steps, "value": batch.n_sites * steps / (ms * 1e-3), "unit": "sites/s", "ms_per_step": ms / steps,
           "e2e": {"value": batch.n_sites * steps / secs, "unit": "sites/s", "h2d_bytes_per_step": int(t_e2e["h2d_bytes"]), "d2h_bytes_per_step": int(t_e2e["d2h_bytes"]),
                   "contexts": nctx},
           "kernel_ms_per_step": {k: v / steps for k, v in kt.items()},
           "tasks": {"em_tasks": t_last["n_tasks_em"], "perm_tasks": t_last["n_tasks_perm"], "called_alleles": int((results.status == 4).sum())},
           "roofline": roofs[dominant]}
    if want_cpu:
        v, nsites, busy, wall, kind = run_reference_sample(wl, batch, cores, ref_per_core)
        out["cpu_baseline"] = {"value": v, "unit": "sites/s", "cores": cores, "kind": kind,
                               "sample": f"{nsites} sites of the same batch, {cores} processes x {ref_per_core} sites, {busy:.1f} s busy"}
    out["wall_s"] = round(time.perf_counter() - t0, 1)
    del flush
    torch.cuda.empty_cache()
    return out


def narrow_pack_us(batch) -> float:
    """Host cost of crispgpu_batch_pack_narrow per site (compact batch -> narrow arena), one core."""
    from crisp_b200.engine import NarrowBatch
    c = batch.compact()
    NarrowBatch(c, pinned=False)
    t0 = time.perf_counter()
    for _ in range(5):
        NarrowBatch(c, pinned=False)
    return 1e6 * (time.perf_counter() - t0) / 5 / batch.n_sites


def bam_end_to_end(cores: int):
    """BASELINE configs[0] as real files: 10 pools x poolsize 20, 100 kb at 100x, written as BAMs; the reference's own program
    (CRISP.binary: one process, and one process per core over --regions shards — its only parallelism, README.md:51) beside
    the same front end bound to the engine (CRISP.gpu, host/crisp_frontend.c).  Candidate sites = sites that reach
    variantcalls.c:115.  The programs are the prebuilt oracle/_ref binaries (compiled from /root/reference)."""
    import tempfile
    from crisp_b200 import frontend as fe
    from crisp_b200.bamsynth import make_dataset
    ref = os.path.join(ROOT, "oracle", "_ref")
    refbin, gpubin, bamtool = (os.path.join(ref, x) for x in ("CRISP.binary", "CRISP.gpu", "bamtool"))
    if not all(os.path.exists(x) for x in (refbin, gpubin, bamtool)):
        return {"unavailable": "oracle/_ref programs not built (they need /root/reference at build time)"}
    out = {"workload": "c1 as BAM files: 10 pools x poolsize 20, 100 kb at 100x per pool, --EM 1 (default flags)"}
    with tempfile.TemporaryDirectory() as d:
        ds = make_dataset(d, pools=10, poolsize=20, length=100000, depth=100, seed=1)
        fe.index_bams(bamtool, ds)
        stats = os.path.join(d, "stats.json")
        t_gpu = fe.run_program(gpubin, ds, os.path.join(d, "gpu.vcf"), 20, env=dict(CRISPGPU_COMPACT=1, CRISPGPU_STATS=stats))
        st = fe.read_stats(stats)
        sites = st["sites"]
        t_one = fe.run_program(refbin, ds, os.path.join(d, "ref.vcf"), 20)
        sh = fe.run_sharded(refbin, ds, os.path.join(d, "ref_sharded.vcf"), 20, cores)
        cmp = fe.compare_vcf(fe.read_vcf(os.path.join(d, "gpu.vcf")), fe.read_vcf(os.path.join(d, "ref.vcf")))
        out.update({"candidate_sites": sites, "vcf_records": st["records"],
                    "vcf_identical_to_reference": not cmp["differ"] and not cmp["only_got"] and not cmp["only_want"],
                    "reference_1_process": {"wall_s": round(t_one, 3), "sites_per_s": sites / t_one},
                    "reference_regions": {"processes": cores, "wall_s": round(sh["wall_s"], 3), "sites_per_s": sites / sh["wall_s"]},
                    "frontend_plus_engine": {"wall_s": round(t_gpu, 3), "sites_per_s": sites / t_gpu, "processes": 1,
                                             "pack_s": st["pack_s"], "engine_s": st["engine_s"], "device_s": st["device_s"], "print_s": st["print_s"],
                                             "batches": st["batches"], "h2d_bytes": st["h2d_bytes"], "d2h_bytes": st["d2h_bytes"]},
                    "note": "the front end (BAM decode + pileup, unchanged reference C, one process) is what remains of the wall time once the engine runs on the GPU"})
        try:
            out["device_pileup"] = device_pileup_leg(ds, cores, sites, t_one, sh["wall_s"], ref_vcf=os.path.join(d, "ref.vcf"))
        except Exception as exc:  # noqa: BLE001
            out["device_pileup"] = {"error": repr(exc)[:300]}
    return out


def device_pileup_leg(ds, cores, frontend_sites, ref_wall_1, ref_wall_regions, local=0, steps=10, ref_vcf=None):
    """SURVEY.md section 8f items 1-3 on the same BAM files: host/bam_decode.c (host threads) -> crispgpu_pileup_run (allele counting,
    candidate screen and site packing on the device, then the engine) — no reference code in the loop.  Reports the decoder's
    rate, the pileup's resident and end-to-end rates (positions/s, candidate sites/s) and the HBM roofline of k_pile_tile."""
    import json as _json
    from crisp_b200.abi import EngineConfig
    from crisp_b200.engine import Engine
    from crisp_b200.pileup import decode_bams, make_window, read_fasta
    P, L = ds["pools"], ds["length"]
    t0 = time.perf_counter()
    aln = decode_bams(ds["bams"], threads=cores)
    t_dec = time.perf_counter() - t0
    ref = read_fasta(ds["fasta"])
    eng = Engine(EngineConfig(ploidy=np.full(P, ds["poolsize"]), options=dict(minimal_results=1)), device=local)
    win = make_window(0, 0, L, ref)
    for _ in range(3):
        n_sites, res = eng.pileup_run(aln, win, want_sites=False)
    t0 = time.perf_counter()
    for _ in range(steps):
        n_sites, res = eng.pileup_run(aln, win, want_sites=False)
    e2e = (time.perf_counter() - t0) / steps
    t_e2e = eng.timing()
    eng.pileup_upload(aln, win)
    for _ in range(3):
        eng.pileup_run_resident(False)
    acc = {"pileup_ms": 0.0, "pile_tile_ms": 0.0, "total_ms": 0.0}
    for _ in range(steps):
        eng.pileup_run_resident(False)
        t = eng.timing()
        for k in acc:
            acc[k] += t[k] / steps
    eng.close()
    peaks = {}
    try:
        peaks = _json.load(open(os.path.join(ROOT, "MEASURED_PEAKS.json")))
    except Exception:
        pass
    hbm = float(peaks.get("hbm_gbs", 6650.0))
    gbs = t["pile_bytes"] / (acc["pile_tile_ms"] * 1e-3) / 1e9
    called = int((res.varalleles >= 1).sum())
    return {"what": "BAM files -> host/bam_decode.c -> crispgpu_pileup_run (pileup, screen, packing, engine on the device); same files as above",
            "alignments": aln.n_reads, "bases": aln.n_bases, "positions": L, "candidate_sites": int(n_sites), "candidate_sites_reference_frontend": frontend_sites,
            "called_sites": called,
            "decode": {"wall_s": round(t_dec, 4), "threads": cores, "compressed_MBps": aln.stats["compressed_bytes"] / t_dec / 1e6, "alignments_per_s": aln.n_reads / t_dec,
                       "inflate_s": round(aln.stats["inflate_s"], 4), "parse_s": round(aln.stats["parse_s"], 4), "read_s": round(aln.stats["read_s"], 4)},
            "e2e": {"ms_per_window": e2e * 1e3, "positions_per_s": L / e2e, "candidate_sites_per_s": n_sites / e2e, "h2d_bytes": int(t_e2e["h2d_bytes"]), "d2h_bytes": int(t_e2e["d2h_bytes"]),
                    "what": "crispgpu_pileup_run from pinned host memory: H2D of the alignments + pileup + engine + D2H of the results, one context, synchronous"},
            "resident": {"pileup_ms": acc["pileup_ms"], "pile_tile_ms": acc["pile_tile_ms"], "total_ms": acc["total_ms"], "positions_per_s": L / (acc["total_ms"] * 1e-3),
                         "candidate_sites_per_s": n_sites / (acc["total_ms"] * 1e-3)},
            "roofline": {"kernel": "k_pile_tile", "bound": "hbm", "achieved": gbs, "peak": hbm, "unit": "GB/s", "frac": gbs / hbm, "traffic": None,
                         "algorithmic_bytes": int(t["pile_bytes"]), "ms_per_launch": acc["pile_tile_ms"],
                         "bytes_definition": "16 per alignment + 1.5 per base (4-bit base + quality byte) + 32 per (position, pool) cell of the count table written"},
            "bam_to_vcf": bam_to_vcf_leg(ds, cores, ref_wall_1, ref_wall_regions, ref_vcf)}


def bam_to_vcf_leg(ds, cores, ref_wall_1, ref_wall_regions, ref_vcf):
    """The whole program without the reference: crisp_b200.caller.call_bams (decode on host threads, pileup + engine on the device,
    VCF text by host/vcf_writer.c, file written) against CRISP.binary's wall time on the same files; records compared."""
    import tempfile
    from crisp_b200.caller import call_bams
    with tempfile.TemporaryDirectory() as d:
        path = os.path.join(d, "b200.vcf")
        r0 = call_bams(ds["bams"], ds["fasta"], ds["poolsize"], out_vcf=path, threads=cores)   # first call: pins the alignment arrays
        r = call_bams(ds["bams"], ds["fasta"], ds["poolsize"], out_vcf=path, threads=cores)
        body = lambda p: [x for x in open(p, "rb").read().split(b"\n") if x and not x.startswith(b"#")]  # noqa: E731
        same = None if not ref_vcf else body(path) == body(ref_vcf)
    return {"wall_s": round(r["wall_s"], 4), "first_call_wall_s": round(r0["wall_s"], 4), "decode_s": round(r["decode_s"], 4), "fasta_s": round(r["fasta_s"], 4), "engine_create_s": round(r["engine_create_s"], 4),
            "device_s": round(r["device_s"], 4), "format_s": round(r["format_s"], 4), "teardown_s": round(r["teardown_s"], 4),
            "records": r["records"], "records_identical_to_CRISP.binary": same,
            "vs_reference_1_process": ref_wall_1 / r["wall_s"], "vs_reference_regions": ref_wall_regions / r["wall_s"],
            "what": "BAM files read to VCF file written with no reference code in the loop (decode threads = host cores; one device window; the engine's destruction afterwards is `teardown_s`); `wall_s` is the second call of the process, which decodes into the pinned arrays the first call allocated (`first_call_wall_s` includes pinning them; the CUDA context exists in both), against CRISP.binary's whole process on the same files: "
                    "one process, and one process per core over --regions shards"}


def main():
    global _REAL_STDOUT
    # keep stdout clean for the single JSON line: libraries that print to fd 1 are routed to stderr
    sys.stdout.flush()
    _REAL_STDOUT = os.dup(1)
    os.dup2(2, 1)
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=40)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--impl", default="b200", choices=["b200", "reference"])
    ap.add_argument("--workload", default="exome50x48", choices=sorted(WORKLOADS))
    ap.add_argument("--sites", type=int, default=8192, help="candidate sites per GPU per step")
    ap.add_argument("--cpu-sites-per-core", type=int, default=512)
    ap.add_argument("--no-cpu-baseline", action="store_true")
    ap.add_argument("--no-configs", action="store_true", help="skip the reduced-size runs of the other BASELINE configs")
    ap.add_argument("--no-bam", action="store_true", help="skip the BAM -> VCF run of config 1 through the reference's front end")
    ap.add_argument("--no-arena", action="store_true", help="end-to-end leg: one pinned allocation per array instead of one arena for the batch")
    ap.add_argument("--contexts", type=int, default=0,
                    help="engine contexts per GPU in the end-to-end leg, one host thread each (batches in flight); 0 = 6, fewer when the ranks of a box would oversubscribe its cores")
    ap.add_argument("--format", "--e2e-format", dest="e2e_format", default="narrow", choices=["narrow", "bins", "reads"],
                    help="batch format of both legs: compact in its narrow transport form (16-bit bins and counts), compact (32-bit read bins + sparse allele counts), or dense arrays with one element per read")
    args = ap.parse_args()
    wl = WORKLOADS[args.workload]
    rank = int(os.environ.get("RANK", "0"))
    world = int(os.environ.get("WORLD_SIZE", "1"))
    local = int(os.environ.get("LOCAL_RANK", "0"))
    config = {"workload": f"{args.workload}: {wl['P']} pools x poolsize {wl['ploidy']}, {wl['depth']}x per pool, options {wl['options'] or 'defaults (--EM 1, asymptotic CT)'}",
              "sites_per_gpu_per_step": args.sites, "sharding": "genomic region per rank, no collective; rank 0 merges the per-rank results in coordinate order (timed: `merge`)",
              "l2_policy": ("inputs larger than L2 (batch > 126 MB)" if args.e2e_format == "reads" and args.sites * (2 * wl['P'] * wl['depth'] + 96 * wl['P']) > 126e6
                            else "L2 flushed between steps (256 MB fill, outside the per-step CUDA events)"),
              "synthetic": {k: (list(v) if isinstance(v, range) else v) for k, v in wl["synth"].items()}}

    if args.impl == "reference":
        if rank != 0:
            return 0
        from oracle.loader import build
        build()
        cores = host_cores()
        batch = make_batch(wl, cores * args.cpu_sites_per_core, seed=1)
        for _ in range(max(0, min(args.warmup, 1))):
            run_reference_sample(wl, batch, cores, max(1, args.cpu_sites_per_core // 8))
        vals, t_total = [], 0.0
        for _ in range(args.steps):
            v, sites, busy, wall, kind = run_reference_sample(wl, batch, cores, args.cpu_sites_per_core)
            vals.append(v)
            t_total += busy
        value = float(np.mean(vals))
        line = {"impl": "reference", "metric": "candidate sites/sec", "value": value, "unit": "sites/s", "n_gpus": args.gpus, "steps": args.steps,
                "warmup": args.warmup, "ms_per_step": 1e3 * t_total / args.steps, "higher_is_better": True, "scaling": "weak", "vs_baseline": None,
                "dtype": "f64", "data": "synthetic", "config": config,
                "cpu_baseline": {"value": value, "unit": "sites/s", "cores": cores, "kind": kind,
                                 "sample": f"{cores} processes x {args.cpu_sites_per_core} sites of the workload, one contiguous region each (the reference's --regions parallelism); sites/s = sites / slowest process; mean of {args.steps} steps"},
                "e2e": {"value": value, "unit": "sites/s", "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0}}
        emit(line)
        return 0

    import torch
    import torch.distributed as dist
    from crisp_b200.abi import EngineConfig
    from crisp_b200.engine import Engine, NarrowBatch, measure_fp64_tflops, measure_int32_tops, pin_batch
    from crisp_b200 import build as cbuild

    if "CRISPGPU_LIB" not in os.environ:
        cbuild.build()
    if not torch.cuda.is_available():
        raise SystemExit("bench.py needs a CUDA device: there is no CPU fallback")
    torch.cuda.set_device(local)
    if world > 1:
        dist.init_process_group("nccl", device_id=torch.device("cuda", local))
    want_ctx = args.contexts if args.contexts > 0 else max(2, min(6, (os.cpu_count() or 8) // max(1, world)))
    # spinning waits are faster, but only while every waiting thread has a core of its own
    oversubscribed = world * (want_ctx + 1) > (os.cpu_count() or 8) // 2
    cfg = EngineConfig(ploidy=np.full(wl["P"], wl["ploidy"]), options=wl["options"])
    cfg_e2e = EngineConfig(ploidy=np.full(wl["P"], wl["ploidy"]), options=dict(wl["options"], blocking_wait=1 if oversubscribed else 0, minimal_results=1))   # as host/crisp_frontend.c asks for them
    # region shard of this rank: its own contiguous run of candidate sites (tid = rank)
    batch = make_batch(wl, args.sites, seed=1 + rank, tid=rank)
    compact_batch = batch.compact()
    # the form both legs consume
    if args.e2e_format == "narrow":
        shipped = pinned = NarrowBatch(batch)
    else:
        shipped = compact_batch if args.e2e_format == "bins" else batch
        pinned = pin_batch(shipped, arena=not args.no_arena)
    stream = torch.cuda.Stream(device=local)
    eng = Engine(cfg, device=local, stream=stream.cuda_stream)
    flush = torch.empty(256 << 20, dtype=torch.uint8, device=f"cuda:{local}") if "flushed" in config["l2_policy"] else None

    def barrier():
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize()

    sampler = ClockSampler(local)
    if rank == 0:
        sampler.start()
    # ---------------- device-resident throughput ----------------
    ms, kt, launches, t_last = time_resident(eng, shipped, args.steps, args.warmup, stream, flush, barrier)
    # ---------------- end to end through the C ABI from pinned host memory ----------------
    # The call sequence a front end makes: crispgpu_submit (asynchronous: H2D + screening kernels) on one context while
    # the previous batch finishes on another (crispgpu_wait: caller kernels + D2H).  Every step copies its inputs from
    # pinned host memory and reads its results back to the host.
    e2e_s, t_e2e, nctx = time_e2e(cfg_e2e, pinned, args.steps, args.warmup, want_ctx, local, barrier)
    clocks = sampler.stop() if rank == 0 else None
    # the same leg in the other batch format (fewer steps: the per-read arrays are 7x the bytes)
    other_fmt = "reads" if args.e2e_format != "reads" else "bins"
    other_steps = max(4, args.steps // 4)
    pinned_other = pin_batch(batch if other_fmt == "reads" else compact_batch, arena=not args.no_arena)
    e2e_other_s, t_e2e_other, _ = time_e2e(cfg_e2e, pinned_other, other_steps, args.warmup, want_ctx, local, barrier)
    del pinned_other
    e2e_wide = None
    if args.e2e_format == "narrow":   # the 32-bit compact form, for comparison
        pinned_wide = pin_batch(compact_batch, arena=not args.no_arena)
        ws, wt, _ = time_e2e(cfg_e2e, pinned_wide, args.steps, args.warmup, want_ctx, local, barrier)
        e2e_wide = (ws, wt)
        del pinned_wide
    results = eng.run(batch)

    # ---------------- coordinate-order merge of the per-rank results on rank 0 ----------------
    merge = None
    if world > 1:
        from crisp_b200.shard import gather_results
        idx = np.arange(batch.n_sites) + rank * batch.n_sites
        gather_results(results, idx, batch.n_sites * world, wl["P"])   # warm: communicator set-up, allocator
        barrier()
        g0 = time.perf_counter()
        merged = gather_results(results, idx, batch.n_sites * world, wl["P"])
        g1 = time.perf_counter()
        if rank == 0:
            # the merged order is (tid, pos): every rank owns one contiguous region, regions ordered by rank
            assert merged.n_sites == batch.n_sites * world and np.array_equal(merged.varalleles[:batch.n_sites], results.varalleles)
            merge = {"ms_per_batch": (g1 - g0) * 1e3, "sites": batch.n_sites * world, "called_alleles": int(merged.n_calls),
                     "note": "rank 0, once per batch of all ranks: every rank's results as one flat byte tensor (two small collectives on the control plane), then "
                             "crisp_b200.shard.merge_results scatters them into coordinate order"}

    times = torch.tensor([ms, e2e_s * 1e3, e2e_other_s * 1e3, (e2e_wide[0] if e2e_wide else 0.0) * 1e3], dtype=torch.float64, device=f"cuda:{local}")
    per_rank = [[float(times[0]) / args.steps, float(times[1]) / args.steps]]
    if world > 1:
        gathered_t = [torch.zeros_like(times) for _ in range(world)]
        dist.all_gather(gathered_t, times)
        per_rank = [[float(g[0]) / args.steps, float(g[1]) / args.steps] for g in gathered_t]
        dist.all_reduce(times, op=dist.ReduceOp.MAX)
    ms_max, e2e_ms_max, e2e_other_ms_max, e2e_wide_ms_max = float(times[0]), float(times[1]), float(times[2]), float(times[3])
    total_sites = batch.n_sites * world
    value = total_sites * args.steps / (ms_max * 1e-3)
    e2e_value = total_sites * args.steps / (e2e_ms_max * 1e-3)
    e2e_other_value = total_sites * other_steps / (e2e_other_ms_max * 1e-3)

    if rank == 0:
        peaks = {}
        try:
            peaks = json.load(open(os.path.join(ROOT, "MEASURED_PEAKS.json")))
        except Exception:
            pass
        hbm_peak = float(peaks.get("hbm_gbs", 6650.0))
        hbm_src = "MEASURED_PEAKS.json hbm_gbs (burst copy)" if "hbm_gbs" in peaks else "fallback 6650 GB/s (B200_PROFILING.md)"
        fp64_peak = measure_fp64_tflops(local)
        int32_peak = measure_int32_tops(local)
        # DRAM traffic per launch from the committed ncu --set full capture of this very configuration (null otherwise)
        traffic, ncu_static = {}, {}
        try:
            tj = json.load(open(os.path.join(ROOT, "profiles", "r2", "traffic.json"))).get(args.workload + (":bins" if args.e2e_format != "reads" else ""), {})
            if tj.get("sites") == batch.n_sites:
                traffic = {k: v["dram_read_bytes"] + v["dram_write_bytes"] for k, v in tj.items() if isinstance(v, dict)}
                ncu_static = {k: {m: v[m] for m in ("issue_active_pct", "fp64_pipe_pct", "warp_instructions", "registers") if m in v} for k, v in tj.items() if isinstance(v, dict)}
        except Exception:
            pass
        roofs, dominant = rooflines(args.workload, wl, batch, compact_batch if args.e2e_format == "narrow" else shipped, results, kt, args.steps, t_last["n_tasks_em"],
                                    "bins" if args.e2e_format == "narrow" else args.e2e_format,
                                    hbm_peak, hbm_src, fp64_peak, int32_peak, traffic, ncu_static)
        second = max((k for k in roofs if k != dominant), key=lambda k: roofs[k]["ms_per_launch"])
        fmt_label = {"narrow": "compact, narrow transport form (crispgpu_batch_pack_narrow): reads pre-binned by the host packer as 16-bit (value, count) entries with one length byte per site x pool cell, per-pool allele counts as 16-bit non-zero entries; widened on the device (k_unpack_bins16, k_narrow_offsets, k_expand_counts16) into the compact arrays k_em consumes",
                     "bins": "compact (pre-binned by the host packer): reads as (value,count) bins per site x pool (consumed as they are by k_em), per-pool allele counts as non-zero entries (k_expand_counts)",
                     "reads": "one 16-bit element per read (the format the reference arm consumes); sites that reach the likelihood fetched on demand (k_fetch_reads)"}
        line = {"metric": "candidate sites/sec", "value": value, "unit": "sites/s", "n_gpus": world, "steps": args.steps, "warmup": args.warmup,
                "ms_per_step": ms_max / args.steps, "higher_is_better": True, "scaling": "weak", "vs_baseline": None, "dtype": "f64",
                "data": "synthetic", "config": config, "clocks": clocks,
                "value_definition": "kernels only, inputs resident in HBM (CUDA events around every step); `e2e` is H2D + kernels + D2H through the C ABI",
                "e2e": {"value": e2e_value, "unit": "sites/s", "h2d_bytes_per_step": int(t_e2e["h2d_bytes"]), "d2h_bytes_per_step": int(t_e2e["d2h_bytes"]),
                        "ms_per_step": e2e_ms_max / args.steps, "h2d_ms": t_e2e["h2d_ms"], "d2h_ms": t_e2e["d2h_ms"],
                        "fetched_bytes_per_step": int(t_e2e["fetched_bytes"]), "input_format": fmt_label[args.e2e_format],
                        "single_context_device_ms": round(float(t_e2e["total_ms"]), 4),
                        "stage_ms_last_batch": {k: round(float(t_e2e[k]), 4) for k in ("h2d_ms", "evaluate_ms", "permute_ms", "filter_ms", "em_ms", "d2h_ms", "total_ms")},
                        "mode": f"{nctx} engine contexts per GPU, one host thread each looping crispgpu_submit/crispgpu_wait on pinned buffers; steps dealt round-robin; " + ("blocking waits (more waiting threads than spare cores)" if oversubscribed else "spinning waits")},
                "e2e_" + other_fmt: {"value": e2e_other_value, "unit": "sites/s", "steps": other_steps, "h2d_bytes_per_step": int(t_e2e_other["h2d_bytes"]),
                                     "d2h_bytes_per_step": int(t_e2e_other["d2h_bytes"]), "fetched_bytes_per_step": int(t_e2e_other["fetched_bytes"]),
                                     "ms_per_step": e2e_other_ms_max / other_steps, "input_format": fmt_label[other_fmt]},
                **({"e2e_bins": {"value": total_sites * args.steps / (e2e_wide_ms_max * 1e-3), "unit": "sites/s", "steps": args.steps, "h2d_bytes_per_step": int(e2e_wide[1]["h2d_bytes"]),
                                 "d2h_bytes_per_step": int(e2e_wide[1]["d2h_bytes"]), "ms_per_step": e2e_wide_ms_max / args.steps, "input_format": fmt_label["bins"]}} if e2e_wide else {}),
                "gpu_launches": int(launches),
                "per_rank_ms_per_step": {"resident": [round(x[0], 4) for x in per_rank], "e2e": [round(x[1], 4) for x in per_rank]},
                "roofline": roofs[dominant], "roofline_secondary": roofs[second],
                "kernel_ms_per_step": {k: v / args.steps for k, v in kt.items()},
                "tasks": {"sites": batch.n_sites, "em_tasks": t_last["n_tasks_em"], "called_alleles": int((results.status == 4).sum()), "perm_tasks": t_last["n_tasks_perm"]}}
        if merge:
            line["merge"] = merge
        eng.close()
        del flush
        torch.cuda.empty_cache()
        if world == 1 and not args.no_cpu_baseline:
            from oracle.loader import RefEngine, build, have_ref
            build()
            cores = host_cores()
            v, sites, busy, wall, kind = run_reference_sample(wl, batch, cores, args.cpu_sites_per_core)
            line["cpu_baseline"] = {"value": v, "unit": "sites/s", "cores": cores, "kind": kind,
                                    "sample": f"{sites} sites of the same batch, {cores} processes x {args.cpu_sites_per_core} sites (contiguous regions), {busy:.1f} s busy"}
            if have_ref():
                # what packing a site costs the host in either batch format: the production packer (host/crisp_shim.c) on the
                # reference's own read lists, one core
                sub = batch.select(np.arange(min(256, batch.n_sites)))
                r = RefEngine(cfg)
                line["host_pack"] = {"reads_us_per_site": 1e6 * r.shim_pack_seconds(sub, False) / sub.n_sites,
                                     "compact_us_per_site": 1e6 * r.shim_pack_seconds(sub, True) / sub.n_sites,
                                     "narrow_extra_us_per_site": narrow_pack_us(sub), "cores": 1, "sites": sub.n_sites,
                                     "what": "crisp_shim_append_site alone (walks of the reference's read lists + packing), per candidate site of this workload, and what crispgpu_batch_pack_narrow adds on top of the compact form; not inside `e2e`, whose batches are packed before the timed region"}
            if not args.no_configs:
                line["configs"] = {}
                for name, nsites, per_core in SIDE_CONFIGS:
                    if name == args.workload:
                        continue
                    try:
                        line["configs"][name] = side_config(name, nsites, per_core, local, (hbm_peak, hbm_src, fp64_peak, int32_peak), cores, True)
                    except Exception as exc:  # noqa: BLE001 — a side run must not take the headline down
                        line["configs"][name] = {"error": repr(exc)[:300]}
            if not args.no_bam:
                try:
                    line["bam_e2e"] = bam_end_to_end(cores)
                except Exception as exc:  # noqa: BLE001
                    line["bam_e2e"] = {"error": repr(exc)[:300]}
        emit(line)
    else:
        eng.close()
    if world > 1:
        dist.destroy_process_group()
    return 0


if __name__ == "__main__":
    sys.exit(main())
