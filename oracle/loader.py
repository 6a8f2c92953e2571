"""TEST INFRASTRUCTURE ONLY — ctypes loaders for the two CPU checkers.

  RefEngine     oracle/_ref/libcrispref.so : the UNMODIFIED reference engine (built by oracle/Makefile from
                /root/reference) driven by oracle/ref_harness.c
  OracleEngine  oracle/liboracle.so        : oracle/crisp_oracle.c, the plain-C restatement

Only tests/, __graft_entry__.smoke() and bench.py's cpu_baseline / --impl reference legs import this module;
nothing under crisp_b200/ does.
"""
from __future__ import annotations

import ctypes as C
import os
import subprocess
import sys

import numpy as np

from crisp_b200.abi import _RESULT_FIELDS, Batch, CBatch, Config, EngineConfig, Results

HERE = os.path.dirname(os.path.abspath(__file__))
REF_LIB = os.path.join(HERE, "_ref", "libcrispref.so")
ORACLE_LIB = os.path.join(HERE, "liboracle.so")

MODE_SLOTS, MODE_CALLER = 0, 1
RNG_DRAND48, RNG_PHILOX = 0, 1


class HostResults(C.Structure):
    _fields_ = (
        [("n_sites", C.c_int32), ("n_calls", C.c_int32), ("max_calls", C.c_int32), ("reserved", C.c_int32)]
        + [(name, C.c_void_p) for name, _, _ in _RESULT_FIELDS]
        + [("ac", C.c_void_p), ("qv", C.c_void_p), ("genll", C.c_void_p), ("genll_stride", C.c_int32),
           ("reserved2", C.c_int32)]
    )


def build(force: bool = False) -> None:
    """Compile liboracle.so and, when /root/reference is present, oracle/_ref/libcrispref.so."""
    args = ["make", "-s", "-C", HERE]
    if force:
        args.append("-B")
    # make's own messages must not reach stdout: bench.py prints exactly one JSON line there
    subprocess.run(args + ["all"], check=True, stdout=sys.stderr)


def _host_results(res: Results, genll: np.ndarray | None) -> HostResults:
    h = HostResults()
    h.n_sites = res.n_sites
    h.max_calls = res.max_calls
    for name, _, _ in _RESULT_FIELDS:
        setattr(h, name, getattr(res, name).ctypes.data)
    h.ac = res.ac.ctypes.data
    h.qv = res.qv.ctypes.data
    if genll is not None:
        h.genll = genll.ctypes.data
        h.genll_stride = genll.shape[-1]
    return h


class _CpuEngine:
    lib_path = ""
    prefix = ""

    def __init__(self, cfg: EngineConfig):
        if not os.path.exists(self.lib_path):
            raise FileNotFoundError(f"{self.lib_path} not built (run `make -C oracle`)")
        self.lib = C.CDLL(self.lib_path)
        self.cfg = cfg
        self._ccfg = cfg.to_c()
        create = getattr(self.lib, self.prefix + "_create")
        create.restype = C.c_void_p
        create.argtypes = [C.POINTER(Config)]
        self.h = create(C.byref(self._ccfg))
        if not self.h:
            raise RuntimeError("oracle create failed")
        self._run = getattr(self.lib, self.prefix + "_run")
        self._run.restype = C.c_int
        self._run.argtypes = [C.c_void_p, C.POINTER(CBatch), C.POINTER(HostResults), C.c_int, C.c_long,
                              C.c_char_p, C.c_size_t, C.POINTER(C.c_size_t)]
        self._err = getattr(self.lib, self.prefix + "_last_error")
        self._err.restype = C.c_char_p
        self._err.argtypes = [C.c_void_p]

    def run(self, batch: Batch, mode: int = MODE_SLOTS, seed: int = 1, want_vcf: bool = False,
            want_genll: bool = False):
        res = Results(batch.n_sites, batch.n_pools)
        genll = None
        if want_genll:
            genll = np.zeros((batch.n_sites, 5, 4, batch.n_pools, int(self.cfg.ploidy.max()) + 1))
        h = _host_results(res, genll)
        cb = batch.as_c()
        cap = (1 << 16) + batch.n_sites * (512 + 64 * batch.n_pools) if want_vcf else 0
        buf = C.create_string_buffer(cap) if cap else None
        ln = C.c_size_t(0)
        rc = self._run(self.h, C.byref(cb), C.byref(h), mode, seed, buf, cap, C.byref(ln))
        if rc != 0:
            raise RuntimeError(f"{self.prefix}_run failed rc={rc}: {self._err(self.h).decode()}")
        res.n_calls = h.n_calls
        res.ac = res.ac[: res.n_calls]
        res.qv = res.qv[: res.n_calls]
        res.vcf = buf.value.decode() if buf is not None else None
        res.genll = genll
        return res


class RefEngine(_CpuEngine):
    lib_path = REF_LIB
    prefix = "crisp_ref"

    def replay_vcf(self, batch: Batch, res: Results, readdepths=None, mqcounts=None) -> str:
        """VCF records of `res` printed by the reference's print_pooledvariant.  readdepths [n][P] / mqcounts [n][4]: the front
        end's own counters for the DP and MQS fields (default: rebuilt from the allele counts)."""
        st = self.lib.crisp_ref_set_printer_state
        st.restype = None
        st.argtypes = [C.c_void_p, C.c_void_p, C.c_void_p]
        rd = None if readdepths is None else np.ascontiguousarray(readdepths, np.int32)
        mq = None if mqcounts is None else np.ascontiguousarray(mqcounts, np.int32)
        st(self.h, None if rd is None else rd.ctypes.data, None if mq is None else mq.ctypes.data)
        try:
            return self._replay_vcf(batch, res)
        finally:
            st(self.h, None, None)

    def _replay_vcf(self, batch: Batch, res: Results) -> str:
        fn = self.lib.crisp_ref_replay_vcf
        fn.restype = C.c_int
        fn.argtypes = [C.c_void_p, C.POINTER(CBatch), C.POINTER(HostResults), C.c_char_p, C.c_size_t,
                       C.POINTER(C.c_size_t)]
        full = Results(res.n_sites, res.n_pools, max_calls=max(res.n_calls, 1))
        for name, _, _ in _RESULT_FIELDS:
            getattr(full, name)[...] = getattr(res, name)
        full.ac[: res.n_calls] = res.ac[: res.n_calls]
        full.qv[: res.n_calls] = res.qv[: res.n_calls]
        h = _host_results(full, None)
        h.n_calls = res.n_calls
        cb = batch.as_c()
        cap = (1 << 16) + batch.n_sites * (512 + 64 * batch.n_pools)
        buf = C.create_string_buffer(cap)
        ln = C.c_size_t(0)
        rc = fn(self.h, C.byref(cb), C.byref(h), buf, cap, C.byref(ln))
        if rc != 0:
            raise RuntimeError("replay failed")
        return buf.value.decode()

    def shim_roundtrip(self, batch: Batch) -> tuple[int, str]:
        """Pack every site again with the production host shim (host/crisp_shim.c) from the rebuilt reference
        structures and compare with `batch`; returns (number of differing arrays, first difference)."""
        fn = self.lib.crisp_ref_shim_roundtrip
        fn.restype = C.c_int
        fn.argtypes = [C.c_void_p, C.POINTER(CBatch), C.c_char_p, C.c_size_t]
        cb = batch.as_c()
        msg = C.create_string_buffer(256)
        rc = fn(self.h, C.byref(cb), msg, 256)
        return rc, msg.value.decode()

    def shim_pack_seconds(self, batch: Batch, compact: bool) -> float:
        """Seconds the production packer (crisp_shim_append_site, one host core) spends on the sites of `batch`."""
        fn = self.lib.crisp_ref_shim_pack_seconds
        fn.restype = C.c_double
        fn.argtypes = [C.c_void_p, C.POINTER(CBatch), C.c_int]
        cb = batch.as_c()
        t = fn(self.h, C.byref(cb), 1 if compact else 0)
        if t < 0:
            raise RuntimeError("shim pack timing failed")
        return t

    def sort_allelecounts(self, batch: Batch):
        """The reference's own sort_allelecounts (parsebam/allelecounts.c:134-183) on every site of the batch."""
        n = batch.n_sites
        al, ct = np.zeros((n, 8), np.uint8), np.zeros((n, 8), np.int32)
        fn = self.lib.crisp_ref_sort_allelecounts
        fn.restype = None
        fn.argtypes = [C.c_void_p, C.POINTER(CBatch), C.c_void_p, C.c_void_p]
        cb = batch.as_c()
        fn(self.h, C.byref(cb), al.ctypes.data, ct.ctypes.data)
        return al, ct

    def scalar(self, name: str, restype, argtypes):
        fn = getattr(self.lib, "crisp_ref_" + name)
        fn.restype = restype
        fn.argtypes = argtypes
        return fn


class OracleEngine(_CpuEngine):
    lib_path = ORACLE_LIB
    prefix = "crisp_oracle"

    def screen(self, batch: Batch):
        """Candidate screen restatement (variantcalls.c:62-92): (maxvec_al, maxvec_ct, potential)."""
        n = batch.n_sites
        al, ct, pot = np.zeros((n, 8), np.uint8), np.zeros((n, 8), np.int32), np.zeros(n, np.int32)
        fn = self.lib.crisp_oracle_screen
        fn.restype = None
        fn.argtypes = [C.c_void_p, C.c_int] + [C.c_void_p] * 6
        fn(self.h, n, batch.refbase.ctypes.data, batch.counts.ctypes.data, batch.indcounts.ctypes.data, al.ctypes.data, ct.ctypes.data, pot.ctypes.data)
        return al, ct, pot

    def scalar(self, name: str, restype, argtypes):
        fn = getattr(self.lib, "crisp_oracle_" + name)
        fn.restype = restype
        fn.argtypes = argtypes
        return fn


def have_ref() -> bool:
    return os.path.exists(REF_LIB)


def have_oracle() -> bool:
    return os.path.exists(ORACLE_LIB)
