"""The C-ABI library: loads, exports every symbol include/crispgpu.h declares, and fails loudly without a GPU."""
import ctypes as C
import os
import re

import numpy as np
import pytest

from crisp_b200 import build as cbuild
from crisp_b200.abi import Batch, CBatch, Config, CResults, EngineConfig, Timing
from crisp_b200.engine import EXPORTS, LIB_PATH, Engine, EngineError, load_library
from crisp_b200.synth import survey_demo_batch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


@pytest.fixture(scope="module")
def lib():
    cbuild.build()
    return load_library()


def test_header_symbols_are_exported(lib):
    header = open(os.path.join(ROOT, "include", "crispgpu.h")).read()
    declared = set(re.findall(r"\b(crispgpu_[a-z0-9_]+)\s*\(", header))
    assert declared == set(EXPORTS)
    for name in declared:
        assert hasattr(lib, name), name


def test_abi_version_and_defaults(lib):
    assert lib.crispgpu_abi_version() == 3
    cfg = Config()
    lib.crispgpu_config_defaults(C.byref(cfg))
    # reference defaults: newcrispcaller.c:20-32, readmultiplebams.c:25-26, crispEM.c:14,26
    assert (cfg.em_flag, cfg.max_iter, cfg.chisq_permutation, cfg.fast_filter, cfg.use_base_qvs, cfg.min_reads) == (1, 20000, 0, 1, 1, 4)
    assert (cfg.qv_offset, cfg.min_q) == (33, 13)
    assert (cfg.thresh1, cfg.min_qpv, cfg.thresh3, cfg.ser, cfg.relax, cfg.qmin) == (-3.5, -5.0, -2.0, -1.0, 0.75, 23.0)
    assert (cfg.agilent_ref_bias, cfg.theta) == (0.5, 0.001)


def test_struct_sizes_match_header():
    # LP64 layout of include/crispgpu.h
    assert C.sizeof(Config) == 4 * 2 + 8 + 4 * 12 + 8 + 8 * 8 + 8 + 4 * 4
    assert C.sizeof(CBatch) == 8 + 8 * 16 + 8 + 8 * 9 + 8 * 7
    assert C.sizeof(CResults) == 8 + 8 * 29
    assert C.sizeof(Timing) == 4 * 8 + 4 * 2 + 8 * 4 + 4 * 4 + 8


def test_argument_errors(lib):
    h = C.c_void_p()
    assert lib.crispgpu_create(None, 0, None, C.byref(h)) == -1
    cfg = EngineConfig(ploidy=np.full(4, 10)).to_c()
    cfg.abi_version = 99
    assert lib.crispgpu_create(C.byref(cfg), 0, None, C.byref(h)) == -1


def test_no_cpu_fallback_without_gpu(lib):
    if lib.crispgpu_device_count() > 0:
        pytest.skip("a GPU is present")
    with pytest.raises(EngineError):
        Engine(EngineConfig(ploidy=np.full(6, 20)))


def test_product_does_not_reference_the_oracle():
    """Nothing under crisp_b200/, host/ or include/ may import, link or execute oracle/."""
    for top in ("crisp_b200", "host", "include"):
        for dirpath, _, files in os.walk(os.path.join(ROOT, top)):
            for f in files:
                if f.endswith((".py", ".cu", ".cuh", ".h", ".c")):
                    text = open(os.path.join(dirpath, f)).read()
                    for needle in ("oracle.loader", "from oracle", "import oracle", "liboracle", "libcrispref", "oracle_abi.h", "crisp_oracle"):
                        assert needle not in text, (dirpath, f, needle)


def test_batch_validation():
    b = survey_demo_batch()
    with pytest.raises(ValueError):
        Batch(6, tid=b.tid, pos=b.pos, refbase=b.refbase, maxvec_al=b.maxvec_al, maxvec_ct=b.maxvec_ct, counts=b.counts,
              indcounts=b.indcounts[:-1], read_off=b.read_off, reads=b.reads)
    with pytest.raises(KeyError):
        EngineConfig(ploidy=np.full(6, 20), options=dict(nonsense=1))


def test_header_is_plain_c_and_links(tmp_path):
    """include/crispgpu.h is the boundary a C front end compiles against: it must be valid C99 on its own, and a C
    program linked against libcrispgpu.so must get the reference's defaults without touching a GPU."""
    import subprocess

    from crisp_b200 import build as cbuild

    src = tmp_path / "front_end.c"
    src.write_text(r'''
#include <stdio.h>
#include "crispgpu.h"
int main(void)
{
	crispgpu_config c;
	crispgpu_batch b = {0};
	crispgpu_bin bin = CRISPGPU_BIN(CRISPGPU_READ('I', 2, 1, 0), 7);
	crispgpu_count cnt = CRISPGPU_COUNT(9, 33);
	crispgpu_config_defaults(&c);
	printf("%d %d %d %d %.1f %u %u %u %d\n", crispgpu_abi_version(), c.em_flag, c.max_iter, c.min_reads, c.thresh1,
	       (unsigned)(bin >> 16), (unsigned)(bin & 0xffff), (unsigned)(cnt >> 27), (int)(b.bins == NULL));
	printf("%u %u %u %u %u %u\n", (unsigned)sizeof(crispgpu_alnrec), (unsigned)sizeof(crispgpu_alignments), (unsigned)sizeof(crispgpu_window),
	       (unsigned)sizeof(crispgpu_pileup_sites), (unsigned)sizeof(crispgpu_timing), (unsigned)sizeof(crispgpu_batch));
	return 0;
}
''')
    exe = tmp_path / "front_end"
    libdir = os.path.dirname(cbuild.LIB)
    subprocess.run(["gcc", "-std=c99", "-Wall", "-Wextra", "-pedantic", "-Werror", "-I", os.path.join(ROOT, "include"), str(src),
                    "-L", libdir, "-lcrispgpu", "-Wl,-rpath," + libdir, "-o", str(exe)], check=True)
    out = subprocess.run([str(exe)], check=True, capture_output=True, text=True).stdout.split()
    assert out[:5] == ["3", "1", "20000", "4", "-3.5"]
    assert out[5:9] == ["7", str(ord("I") | (2 << 8) | (1 << 10)), "9", "1"]
    from crisp_b200.pileup import ALNREC_DTYPE, CAlignments, CPileupSites, CWindow
    from crisp_b200.abi import CBatch
    assert [int(x) for x in out[9:]] == [ALNREC_DTYPE.itemsize, C.sizeof(CAlignments), C.sizeof(CWindow), C.sizeof(CPileupSites), C.sizeof(Timing), C.sizeof(CBatch)]
