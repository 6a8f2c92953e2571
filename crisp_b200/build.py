"""Build crisp_b200/csrc/libcrispgpu.so in-tree with nvcc for sm_100a (cross-compiles without a GPU)."""
from __future__ import annotations

import os
import subprocess
import sys

HERE = os.path.dirname(os.path.abspath(__file__))
CSRC = os.path.join(HERE, "csrc")
LIB = os.path.join(CSRC, "libcrispgpu.so")
SOURCES = ["crispgpu.cu"]
HEADERS = ["dmath.cuh", "engine_types.cuh", "k_evaluate.cuh", "k_em.cuh", "k_permute.cuh", "k_legacy.cuh", "k_pileup.cuh",
           os.path.join("..", "..", "include", "crispgpu.h")]
NVCC_FLAGS = [
    "-O3", "-std=c++17", "-gencode", "arch=compute_100a,code=sm_100a", "-lineinfo",
    "-Xcompiler", "-fPIC", "-shared", "--fmad=true", "-Xptxas", "-v",
]


BAMLIB = os.path.join(CSRC, "libcrispbam.so")
HOST = os.path.join(HERE, "..", "host")


def build_bam_decoder(force: bool = False) -> str:
    """host/bam_decode.c (parallel BGZF inflate + BAM record split, plain C with zlib and pthreads) -> csrc/libcrispbam.so."""
    src = [os.path.join(HOST, "bam_decode.c"), os.path.join(HOST, "bam_decode.h"), os.path.join(HERE, "..", "include", "crispgpu.h")]
    if not force and os.path.exists(BAMLIB) and all(os.path.getmtime(f) <= os.path.getmtime(BAMLIB) for f in src):
        return BAMLIB
    cmd = [os.environ.get("CC", "gcc"), "-O2", "-std=gnu99", "-Wall", "-fPIC", "-shared", "-o", BAMLIB, src[0], "-lz", "-lpthread"]
    res = subprocess.run(cmd, capture_output=True, text=True)
    if res.returncode != 0:
        sys.stderr.write(res.stdout + res.stderr)
        raise RuntimeError("gcc failed building libcrispbam.so")
    return BAMLIB


VCFLIB = os.path.join(CSRC, "libcrispvcf.so")


def build_vcf_writer(force: bool = False) -> str:
    """host/vcf_writer.c (batch VCF record formatter) + host/vcf_post.c (pooled-count -> genotype conversion, BGZF output), plain C
    with pthreads and zlib -> csrc/libcrispvcf.so."""
    src = [os.path.join(HOST, "vcf_writer.c"), os.path.join(HOST, "vcf_post.c"), os.path.join(HOST, "vcf_writer.h"), os.path.join(HOST, "vcf_post.h"),
           os.path.join(HERE, "..", "include", "crispgpu.h")]
    if not force and os.path.exists(VCFLIB) and all(os.path.getmtime(f) <= os.path.getmtime(VCFLIB) for f in src):
        return VCFLIB
    cmd = [os.environ.get("CC", "gcc"), "-O2", "-std=gnu99", "-Wall", "-fPIC", "-shared", "-o", VCFLIB, src[0], src[1], "-lpthread", "-lz"]
    res = subprocess.run(cmd, capture_output=True, text=True)
    if res.returncode != 0:
        sys.stderr.write(res.stdout + res.stderr)
        raise RuntimeError("gcc failed building libcrispvcf.so")
    return VCFLIB


def needs_build() -> bool:
    if not os.path.exists(LIB):
        return True
    t = os.path.getmtime(LIB)
    return any(os.path.getmtime(os.path.join(CSRC, f)) > t for f in SOURCES + HEADERS)


def build(force: bool = False, verbose: bool = False) -> str:
    build_bam_decoder(force)
    build_vcf_writer(force)
    if not force and not needs_build():
        return LIB
    nvcc = os.environ.get("NVCC", "/usr/local/cuda/bin/nvcc")
    cmd = [nvcc] + NVCC_FLAGS + ["-o", LIB] + [os.path.join(CSRC, s) for s in SOURCES]
    res = subprocess.run(cmd, capture_output=True, text=True)
    if verbose or res.returncode != 0:
        sys.stderr.write(res.stdout + res.stderr)
    if res.returncode != 0:
        raise RuntimeError("nvcc failed building libcrispgpu.so")
    with open(os.path.join(CSRC, "ptxas_info.txt"), "w") as f:
        f.write(res.stderr)
    return LIB


if __name__ == "__main__":
    print(build(force="--force" in sys.argv, verbose=True))
