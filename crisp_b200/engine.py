"""Host-side mirror of the C ABI: `Engine` drives crisp_b200/csrc/libcrispgpu.so through ctypes.

This is the replacement for the per-site call `newCRISPcaller(...)` (variantcalls.c:115): the front end packs
candidate sites into a `Batch`, `Engine.run` returns the fields print_pooledvariant (crisp/newcrispprint.c:347)
consumes.  There is no CPU fallback: a missing library or a missing GPU raises.
"""
from __future__ import annotations

import ctypes as C
import os

import numpy as np

from .abi import Batch, CBatch, Config, CResults, EngineConfig, Results, Timing

LIB_PATH = os.path.join(os.path.dirname(os.path.abspath(__file__)), "csrc", "libcrispgpu.so")

EXPORTS = [
    "crispgpu_config_defaults", "crispgpu_create", "crispgpu_destroy", "crispgpu_submit", "crispgpu_wait", "crispgpu_advance", "crispgpu_screen_positions",
    "crispgpu_upload", "crispgpu_run_resident", "crispgpu_get_timing", "crispgpu_last_error",
    "crispgpu_host_alloc", "crispgpu_host_free", "crispgpu_abi_version", "crispgpu_device_count",
    "crispgpu_measure_fp64_tflops", "crispgpu_measure_int32_tops",
    "crispgpu_pileup_run", "crispgpu_pileup_upload", "crispgpu_pileup_run_resident", "crispgpu_pileup_set_window",
    "crispgpu_batch_narrow_bytes", "crispgpu_batch_pack_narrow",
]

_lib = None


class EngineError(RuntimeError):
    pass


def load_library() -> C.CDLL:
    """Load libcrispgpu.so (built in-tree by `python -m crisp_b200.build`); raises if it is absent."""
    global _lib
    if _lib is not None:
        return _lib
    path = os.environ.get("CRISPGPU_LIB", LIB_PATH)   # the same override host/crisp_frontend.c honours (A/B measurements)
    if not os.path.exists(path):
        raise EngineError(f"{path} is missing: build it with `python -m crisp_b200.build` (there is no CPU fallback)")
    lib = C.CDLL(path)
    lib.crispgpu_config_defaults.argtypes = [C.POINTER(Config)]
    lib.crispgpu_config_defaults.restype = None
    lib.crispgpu_create.argtypes = [C.POINTER(Config), C.c_int, C.c_void_p, C.POINTER(C.c_void_p)]
    lib.crispgpu_create.restype = C.c_int
    lib.crispgpu_destroy.argtypes = [C.c_void_p]
    lib.crispgpu_destroy.restype = None
    lib.crispgpu_submit.argtypes = [C.c_void_p, C.POINTER(CBatch), C.POINTER(C.c_int64)]
    lib.crispgpu_submit.restype = C.c_int
    lib.crispgpu_wait.argtypes = [C.c_void_p, C.c_int64, C.POINTER(CResults)]
    lib.crispgpu_wait.restype = C.c_int
    lib.crispgpu_upload.argtypes = [C.c_void_p, C.POINTER(CBatch)]
    lib.crispgpu_upload.restype = C.c_int
    lib.crispgpu_run_resident.argtypes = [C.c_void_p, C.c_int, C.POINTER(CResults)]
    lib.crispgpu_run_resident.restype = C.c_int
    lib.crispgpu_get_timing.argtypes = [C.c_void_p, C.POINTER(Timing)]
    lib.crispgpu_get_timing.restype = C.c_int
    lib.crispgpu_last_error.argtypes = [C.c_void_p]
    lib.crispgpu_last_error.restype = C.c_char_p
    lib.crispgpu_host_alloc.argtypes = [C.c_size_t]
    lib.crispgpu_host_alloc.restype = C.c_void_p
    lib.crispgpu_host_free.argtypes = [C.c_void_p]
    lib.crispgpu_host_free.restype = None
    lib.crispgpu_abi_version.restype = C.c_int
    lib.crispgpu_device_count.restype = C.c_int
    lib.crispgpu_measure_fp64_tflops.argtypes = [C.c_int, C.POINTER(C.c_double)]
    lib.crispgpu_measure_fp64_tflops.restype = C.c_int
    _lib = lib
    return lib


class _PinnedBlock:
    """One crispgpu_host_alloc allocation; freed when the last array viewing it is collected."""

    def __init__(self, nbytes: int):
        self.lib = load_library()
        self.ptr = self.lib.crispgpu_host_alloc(max(nbytes, 1))
        if not self.ptr:
            raise EngineError("crispgpu_host_alloc failed")
        self.buf = (C.c_char * max(nbytes, 1)).from_address(self.ptr)

    def __del__(self):
        try:
            self.lib.crispgpu_host_free(self.ptr)
        except Exception:
            pass


def pinned_like(a: np.ndarray):
    """(copy of `a` in pinned host memory, owner object that must outlive it)."""
    block = _PinnedBlock(a.nbytes)
    out = np.frombuffer(block.buf, dtype=a.dtype, count=a.size).reshape(a.shape)
    out[...] = a
    return out, block


def pin_batch(batch: Batch, arena: bool = True) -> Batch:
    """A copy of the batch in pinned memory, so that crispgpu_submit's copies are asynchronous.  With `arena` (default) all
    arrays are packed 256-byte aligned into ONE pinned allocation that the batch declares (`arena_base`/`arena_bytes`):
    the engine then moves the batch with a single copy."""
    from .abi import _BATCH_ARRAYS

    out = Batch.__new__(Batch)
    out.__dict__.update({k: v for k, v in batch.__dict__.items() if not k.startswith("_")})
    out._pinned_blocks = []
    present = [(name, getattr(batch, name)) for name, _, _ in _BATCH_ARRAYS if getattr(batch, name) is not None]
    if arena:
        offs, total = [], 0
        for _, a in present:
            offs.append(total)
            total += (a.nbytes + 255) & ~255
        block = _PinnedBlock(total)
        out._pinned_blocks.append(block)
        for (name, a), off in zip(present, offs):
            view = np.frombuffer(block.buf, dtype=a.dtype, count=a.size, offset=off).reshape(a.shape)
            view[...] = a
            setattr(out, name, view)
        out._arena = (block.ptr, total)
    else:
        for name, a in present:
            arr, block = pinned_like(a)
            out._pinned_blocks.append(block)
            setattr(out, name, arr)
    return out


class NarrowBatch:
    """A compact batch packed into one pinned arena in the narrow transport form (16-bit bins and counts, one length byte per
    cell: crispgpu_batch_pack_narrow) — what a front end ships when the host link is the bound.  Takes the place of a `Batch`
    in Engine.submit / upload."""

    def __init__(self, batch: Batch, pinned: bool = True):
        lib = load_library()
        lib.crispgpu_batch_narrow_bytes.argtypes = [C.POINTER(CBatch), C.c_int32]
        lib.crispgpu_batch_narrow_bytes.restype = C.c_size_t
        lib.crispgpu_batch_pack_narrow.argtypes = [C.POINTER(CBatch), C.c_int32, C.c_void_p, C.c_size_t, C.POINTER(CBatch)]
        lib.crispgpu_batch_pack_narrow.restype = C.c_int
        src = batch.compact()
        csrc = src.as_c()
        csrc.arena_base, csrc.arena_bytes = None, 0
        need = lib.crispgpu_batch_narrow_bytes(C.byref(csrc), src.n_pools)
        if need == 0:
            raise EngineError("batch cannot be packed in the narrow form (a cell with more than 255 bins or count entries)")
        if pinned:
            self._block = _PinnedBlock(need)
            ptr = self._block.ptr
        else:   # pageable memory (CPU-side tests of the packer)
            self._block = np.empty(need + 256, np.uint8)
            ptr = (self._block.ctypes.data + 255) & ~255
        self._c = CBatch()
        rc = lib.crispgpu_batch_pack_narrow(C.byref(csrc), src.n_pools, ptr, need, C.byref(self._c))
        if rc != 0:
            raise EngineError(f"crispgpu_batch_pack_narrow failed: status {rc}")
        self.n_sites, self.n_pools = batch.n_sites, batch.n_pools

    def nbytes(self) -> int:
        return int(self._c.arena_bytes)

    def array(self, name: str, dtype, count: int) -> np.ndarray:
        """View of one arena array (tests)."""
        ptr = getattr(self._c, name)
        return np.ctypeslib.as_array(C.cast(ptr, C.POINTER(C.c_uint8)), shape=(count * np.dtype(dtype).itemsize,)).view(dtype) if ptr and count else np.zeros(0, dtype)

    def as_c(self) -> CBatch:
        return self._c


class Engine:
    """One engine context bound to one CUDA device (`crispgpu_ctx`)."""

    def __init__(self, cfg: EngineConfig, device: int = 0, stream: int | None = None):
        self.lib = load_library()
        self.cfg = cfg
        self._ccfg = cfg.to_c()
        h = C.c_void_p()
        rc = self.lib.crispgpu_create(C.byref(self._ccfg), device, C.c_void_p(stream) if stream else None, C.byref(h))
        if rc != 0:
            raise EngineError(f"crispgpu_create failed: status {rc} (no usable sm_100 device?)" if rc == -2 else f"crispgpu_create failed: status {rc}")
        self.h = h
        self.device = device

    def close(self):
        if getattr(self, "h", None):
            self.lib.crispgpu_destroy(self.h)
            self.h = None

    def __del__(self):
        try:
            self.close()
        except Exception:
            pass

    def _check(self, rc: int, what: str):
        if rc != 0:
            raise EngineError(f"{what} failed: status {rc}: {self.lib.crispgpu_last_error(self.h).decode()}")

    # ---- the call a front end makes: host batch in, host results out ----
    def submit(self, batch: Batch) -> int:
        cb = batch.as_c()
        t = C.c_int64(0)
        self._check(self.lib.crispgpu_submit(self.h, C.byref(cb), C.byref(t)), "crispgpu_submit")
        self._inflight = (cb, batch)
        return t.value

    def advance(self, ticket: int) -> None:
        """Enqueue the batch's caller kernels and result copies without waiting for them (crispgpu_advance)."""
        self._check(self.lib.crispgpu_advance(self.h, C.c_int64(ticket)), "crispgpu_advance")

    def wait(self, ticket: int, copy: bool = True):
        cres = CResults()
        self._check(self.lib.crispgpu_wait(self.h, ticket, C.byref(cres)), "crispgpu_wait")
        self._inflight = None
        return Results.from_c(cres, self.cfg.n_pools) if copy else cres

    def screen_positions(self, batch: Batch):
        """(maxvec_al [n,8], maxvec_ct [n,8], potential [n]) of the candidate screen (crispgpu_screen_positions) for the
        positions of `batch` (its counts, refbase and dense or sparse per-pool counts)."""
        n = batch.n_sites
        al, ct, pot = np.zeros((n, 8), np.uint8), np.zeros((n, 8), np.int32), np.zeros(n, np.int32)
        ptr = lambda a: a.ctypes.data if a is not None else None  # noqa: E731
        fn = self.lib.crispgpu_screen_positions
        fn.argtypes = [C.c_void_p, C.c_int32] + [C.c_void_p] * 8
        self._check(fn(self.h, n, ptr(batch.refbase), ptr(batch.counts), ptr(batch.indcounts), ptr(batch.ic_off), ptr(batch.ic_sparse),
                       ptr(al), ptr(ct), ptr(pot)), "crispgpu_screen_positions")
        return al, ct, pot

    def run(self, batch: Batch) -> Results:
        return self.wait(self.submit(batch))

    # ---- device-resident path (kernel timing) ----
    def upload(self, batch: Batch) -> None:
        cb = batch.as_c()
        self._check(self.lib.crispgpu_upload(self.h, C.byref(cb)), "crispgpu_upload")

    def run_resident(self, fetch: bool = False):
        cres = CResults()
        self._check(self.lib.crispgpu_run_resident(self.h, 1 if fetch else 0, C.byref(cres)), "crispgpu_run_resident")
        return Results.from_c(cres, self.cfg.n_pools) if fetch else None

    # ---- pileup on the device: alignments in, candidate sites + their results out ----
    def pileup_run(self, aln, window, want_sites: bool = True):
        """crispgpu_pileup_run: (PileupSites, Results) for the candidate sites of `window` (a crisp_b200.pileup.make_window
        pair) piled up from `aln` (crisp_b200.pileup.Alignments)."""
        from .pileup import CPileupSites, sites_from_c

        cw, keep = window
        ca = aln.as_c()
        cs, cres = CPileupSites(), CResults()
        fn = self.lib.crispgpu_pileup_run
        fn.argtypes = [C.c_void_p, C.c_void_p, C.c_void_p, C.c_void_p, C.c_void_p]
        fn.restype = C.c_int
        self._check(fn(self.h, C.byref(ca), C.byref(cw), C.byref(cs), C.byref(cres)), "crispgpu_pileup_run")
        del keep
        return (sites_from_c(cs, self.cfg.n_pools) if want_sites else cs.n_sites), Results.from_c(cres, self.cfg.n_pools)

    def pileup_upload(self, aln, window) -> None:
        cw, keep = window
        ca = aln.as_c()
        fn = self.lib.crispgpu_pileup_upload
        fn.argtypes = [C.c_void_p, C.c_void_p, C.c_void_p]
        fn.restype = C.c_int
        self._check(fn(self.h, C.byref(ca), C.byref(cw)), "crispgpu_pileup_upload")
        del keep

    def pileup_set_window(self, window) -> None:
        """Another window over the resident alignments (crispgpu_pileup_set_window); a window made with refseq=None keeps the
        reference stretch that is on the device."""
        cw, keep = window
        fn = self.lib.crispgpu_pileup_set_window
        fn.argtypes = [C.c_void_p, C.c_void_p]
        fn.restype = C.c_int
        self._check(fn(self.h, C.byref(cw)), "crispgpu_pileup_set_window")
        del keep

    def pileup_run_resident(self, fetch: bool = False):
        from .pileup import CPileupSites, sites_from_c

        cs, cres = CPileupSites(), CResults()
        fn = self.lib.crispgpu_pileup_run_resident
        fn.argtypes = [C.c_void_p, C.c_int, C.c_void_p, C.c_void_p]
        fn.restype = C.c_int
        self._check(fn(self.h, 1 if fetch else 0, C.byref(cs), C.byref(cres)), "crispgpu_pileup_run_resident")
        if not fetch:
            return cs.n_sites
        return sites_from_c(cs, self.cfg.n_pools), Results.from_c(cres, self.cfg.n_pools)

    def timing(self) -> dict:
        t = Timing()
        self._check(self.lib.crispgpu_get_timing(self.h, C.byref(t)), "crispgpu_get_timing")
        return {name: getattr(t, name) for name, _ in Timing._fields_}


def measure_fp64_tflops(device: int = 0) -> float:
    """FP64 FMA throughput of the device measured with a register-resident DFMA loop (TFLOP/s)."""
    v = C.c_double(0.0)
    rc = load_library().crispgpu_measure_fp64_tflops(device, C.byref(v))
    if rc != 0:
        raise EngineError(f"crispgpu_measure_fp64_tflops failed: status {rc}")
    return v.value


def measure_int32_tops(device: int = 0) -> float:
    """INT32 multiply-add throughput of the device (10^12 lane-instructions per second), measured with an IMAD loop."""
    v = C.c_double(0.0)
    lib = load_library()
    lib.crispgpu_measure_int32_tops.argtypes = [C.c_int, C.POINTER(C.c_double)]
    lib.crispgpu_measure_int32_tops.restype = C.c_int
    rc = lib.crispgpu_measure_int32_tops(device, C.byref(v))
    if rc != 0:
        raise EngineError(f"crispgpu_measure_int32_tops failed: status {rc}")
    return v.value


def device_count() -> int:
    return load_library().crispgpu_device_count()
