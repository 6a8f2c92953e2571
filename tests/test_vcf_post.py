"""host/vcf_post.c (the part after the records are formatted): the pooled-count -> genotype conversion of
scripts/convert_pooled_vcf.py and BGZF output.  The conversion is checked against the restatement (oracle/convert_pooled_oracle.py),
both against the outputs of the reference's own script committed under tests/golden/convert_pooled/ (scripts/make_convert_golden.py)
and, where /root/reference exists, against the script run live on fresh inputs."""
import glob
import gzip
import json
import os
import struct
import sys
import zlib

import numpy as np
import pytest

from crisp_b200.vcf import bgzf_compress, expand_genotypes
from oracle.convert_pooled_oracle import convert

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
GOLD = os.path.join(ROOT, "tests", "golden", "convert_pooled")
NAMES = sorted(os.path.basename(p)[:-len(".args.json")] for p in glob.glob(os.path.join(GOLD, "*.args.json")))


def _load(name):
    args = json.load(open(os.path.join(GOLD, name + ".args.json")))
    return open(os.path.join(GOLD, name + ".vcf")).read(), open(os.path.join(GOLD, name + ".expected.vcf")).read(), args


@pytest.mark.parametrize("name", NAMES)
def test_conversion_matches_the_reference_script_outputs(name):
    vcf, want, args = _load(name)
    got_oracle, st_oracle = convert(vcf, args["poolsize"] or -1, args["pool_sizes"])
    assert got_oracle == want, "the restatement differs from the reference script's output"
    got, st = expand_genotypes(vcf.encode(), poolsize=args["poolsize"], pool_sizes=args["pool_sizes"])
    assert got.decode() == want
    assert st == st_oracle
    assert st["n_records_in"] == st["n_records_out"] + st["n_multiallelic_indel"] + st["n_over_poolsize"]


def test_golden_set_covers_every_branch():
    """What the fixtures exercise: both pool-size modes, missing genotypes, a multi-allelic indel, counts above the pool size,
    a real CRISP.binary file."""
    _, _, a = _load("fixed4")
    st = expand_genotypes(_load("fixed4")[0].encode(), poolsize=a["poolsize"])[1]
    assert st["n_multiallelic_indel"] == 1 and st["n_over_poolsize"] == 1 and st["n_records_out"] == 5
    v = _load("variable")
    st = expand_genotypes(v[0].encode(), pool_sizes=v[2]["pool_sizes"])[1]
    assert st["n_over_poolsize"] == 1 and st["n_records_out"] == 3
    assert "./././." in _load("fixed4")[1] and "./././././." in v[1]
    assert "crisp_binary_pooled" in NAMES


def _random_vcf(rng, n_records, n_pools, poolsize, variable):
    lines = ["##fileformat=VCFv4.0", "#CHROM\tPOS\tID\tREF\tALT\tQUAL\tFILTER\tINFO\tFORMAT\t" + "\t".join(f"p{i}" for i in range(n_pools))]
    for r in range(n_records):
        nalt = int(rng.integers(1, 4))
        indel = rng.random() < 0.15
        ref = "AC" if indel else "A"
        alts = [("A" if indel and k == 0 else ("ACC" if indel else "CGT"[k])) for k in range(nalt)]
        cols = []
        for i in range(n_pools):
            if rng.random() < 0.1:
                cols.append(".:0:0")
                continue
            size = poolsize[i] if variable else poolsize
            k = nalt + (1 if variable else 0)
            cnt = rng.multinomial(int(rng.integers(0, size + (2 if rng.random() < 0.1 else 0) + 1)) if not variable else size + (1 if rng.random() < 0.05 else 0),
                                  np.ones(k) / k)
            cols.append(",".join(str(int(c)) for c in cnt) + f":{int(rng.integers(0, 99))}:{int(rng.integers(0, 300))}")
        lines.append(f"chr{1 + r % 3}\t{100 + 7 * r}\t.\t{ref}\t{','.join(alts)}\t{int(rng.integers(0, 999))}\tPASS\tNP={n_pools}\tAC:GQ:DP\t" + "\t".join(cols))
    return "\n".join(lines) + "\n"


@pytest.mark.parametrize("variable", [False, True], ids=["one_pool_size", "pool_sizes"])
def test_conversion_matches_restatement_on_random_files(variable):
    rng = np.random.default_rng(5 + variable)
    for trial in range(6):
        n_pools = int(rng.integers(1, 40))
        sizes = [int(x) for x in rng.integers(1, 60, n_pools)]
        poolsize = sizes if variable else int(rng.integers(1, 60))
        vcf = _random_vcf(rng, 200, n_pools, poolsize, variable)
        want, st_want = convert(vcf, -1 if variable else poolsize, sizes if variable else None)
        got, st = expand_genotypes(vcf.encode(), poolsize=None if variable else poolsize, pool_sizes=sizes if variable else None)
        assert got.decode() == want and st == st_want
        assert st["n_records_out"] > 0


def test_conversion_matches_the_reference_script_live(tmp_path):
    """The reference's script itself (syntax translated in memory, scripts/make_convert_golden.py) on fresh random files."""
    if not os.path.exists("/root/reference/scripts/convert_pooled_vcf.py"):
        pytest.skip("needs /root/reference")
    sys.path.insert(0, os.path.join(ROOT, "scripts"))
    from make_convert_golden import run_reference
    rng = np.random.default_rng(11)
    for variable in (False, True):
        n_pools = 7
        sizes = [int(x) for x in rng.integers(2, 30, n_pools)]
        poolsize = sizes if variable else 12
        vcf = _random_vcf(rng, 150, n_pools, poolsize, variable)
        bams = "".join(f"p{i}.bam" + (f" PS={sizes[i]}" if variable else "") + "\n" for i in range(n_pools))
        want, rc = run_reference(vcf, bams, None if variable else poolsize)
        assert rc == 0
        got, _ = expand_genotypes(vcf.encode(), poolsize=None if variable else poolsize, pool_sizes=sizes if variable else None)
        assert got.decode() == want
        assert convert(vcf, -1 if variable else poolsize, sizes if variable else None)[0] == want


def test_conversion_argument_and_format_errors():
    vcf = b"#h\nchr1\t1\t.\tA\tG\t9\tPASS\tX\tAC:GQ\t1:3\n"
    with pytest.raises(ValueError):
        expand_genotypes(vcf)                                  # no pool size from either source (script :23-24)
    with pytest.raises(ValueError):
        expand_genotypes(b"chr1\t1\t.\tA\tG\n", poolsize=4)   # a record without FORMAT column
    with pytest.raises(ValueError):
        expand_genotypes(vcf.replace(b"1:3", b"x:3"), poolsize=4)   # a count that is not a number (ValueError in the script)
    with pytest.raises(ValueError):
        expand_genotypes(vcf, pool_sizes=[])                   # more sample columns than pool sizes (IndexError in the script)
    got, st = expand_genotypes(b"", poolsize=4)
    assert got == b"" and st["n_records_in"] == 0
    got, _ = expand_genotypes(vcf[:-1], poolsize=2)            # last line without newline
    assert got == b"#h\nchr1\t1\t.\tA\tG\t9\tPASS\tX\tGT:GQ\t0/1:3\n"


def _bgzf_blocks(data: bytes):
    """(offset, block size, input size) of every BGZF member, checking the framing of samtools/bgzf.c."""
    out, o = [], 0
    while o < len(data):
        assert data[o:o + 4] == b"\x1f\x8b\x08\x04" and data[o + 10:o + 16] == b"\x06\x00BC\x02\x00", f"bad BGZF header at {o}"
        bsize = struct.unpack_from("<H", data, o + 16)[0] + 1
        crc, isize = struct.unpack_from("<II", data, o + bsize - 8)
        raw = zlib.decompress(data[o + 18:o + bsize - 8], -15)
        assert len(raw) == isize and zlib.crc32(raw) == crc
        out.append((o, bsize, isize))
        o += bsize
    assert o == len(data)
    return out


def test_bgzf_output_is_what_bgzip_writes():
    rng = np.random.default_rng(3)
    text = open(os.path.join(GOLD, "crisp_binary_pooled.vcf"), "rb").read() * 40          # 450 KB of VCF text: 7 blocks
    noise = rng.integers(0, 256, 200000, dtype=np.uint8).tobytes()                      # incompressible: the stored fallback
    for data in (b"", b"x", text, noise, text[:0xff00], text[:0xff00 + 1]):
        z = bgzf_compress(data, threads=8)
        assert gzip.decompress(z) == data
        blocks = _bgzf_blocks(z)
        assert blocks[-1][1:] == (28, 0) and z[-28:] == bytes.fromhex("1f8b08040000000000ff0600424302001b0003000000000000000000")
        assert all(b[1] <= 0x10000 and b[2] <= 0xff00 for b in blocks)
        assert len(blocks) == 1 + (len(data) + 0xff00 - 1) // 0xff00
        assert z == bgzf_compress(data, threads=1)                                       # independent of the thread count
    assert len(bgzf_compress(text, level=9)) <= len(bgzf_compress(text, level=1)) < len(text) // 3
    assert gzip.decompress(bgzf_compress(text, level=0)) == text


def test_bgzf_output_is_read_by_the_bam_decoder_indexer():
    """The writer's blocks through the reader's framing code (host/bam_decode.c index_blocks): a BAM written with it decodes."""
    from crisp_b200.bamsynth import make_dataset
    from crisp_b200.pileup import decode_bams
    import tempfile
    with tempfile.TemporaryDirectory() as d:
        ds = make_dataset(d, pools=2, poolsize=4, length=3000, depth=8, seed=2)
        raw = [gzip.decompress(open(b, "rb").read()) for b in ds["bams"]]
        paths = []
        for k, r in enumerate(raw):
            p = os.path.join(d, f"re{k}.bam")
            open(p, "wb").write(bgzf_compress(r, threads=4))
            paths.append(p)
        a, b = decode_bams(ds["bams"], tid=0, threads=2, pinned=False), decode_bams(paths, tid=0, threads=2, pinned=False)
        assert a.n_reads == b.n_reads and a.n_reads > 0
        assert np.array_equal(a.rec, b.rec) and np.array_equal(a.seq4[:a.n_bases // 2], b.seq4[:b.n_bases // 2])


def test_call_bams_output_forms_with_the_device_stubbed(tmp_path, monkeypatch):
    """The file-writing end of crisp_b200.caller.call_bams (plain, BGZF, genotype lists) on the CPU: decode is the real one into
    pageable memory, the engine and the record formatter are stand-ins (the real chain is tests/test_pileup.py, GPU)."""
    import types
    import crisp_b200.caller as caller
    from crisp_b200 import pileup
    from crisp_b200.bamsynth import make_dataset

    class StubEngine:
        def __init__(self, cfg, device=0): pass
        def pileup_upload(self, aln, win): pass
        def pileup_set_window(self, win): pass
        def pileup_run_resident(self, fetch=True): return types.SimpleNamespace(n_sites=2), None
        def close(self): pass

    rec = b"chr1\t100\t.\tA\tG\t312\tPASS\tNP=2\tAC:GQ:DP\t1:33:60\t0:50:58\nchr1\t200\t.\tC\tT\t90\tPASS\tNP=2\tAC:GQ:DP\t4:20:60\t.:0:0\n"
    real_decode = pileup.decode_bams
    decoded = []

    def decode(bams, tid=0, threads=None, reuse=None, **kw):
        decoded.append(kw)
        return real_decode(bams, tid=tid, threads=threads, pinned=False, **kw)

    monkeypatch.setattr(caller, "decode_bams", decode)
    monkeypatch.setattr(caller, "Engine", StubEngine)
    monkeypatch.setattr(caller, "format_records", lambda *a, **k: rec)
    monkeypatch.setattr(caller, "_ARENA", {})
    ds = make_dataset(str(tmp_path), pools=2, poolsize=4, length=3000, depth=8, seed=2)
    texts = {}
    for name, kw in (("a.vcf", {}), ("b.vcf.gz", {}), ("c.vcf.gz", dict(expand_gt=True)), ("d.vcf", dict(expand_gt=True))):
        p = str(tmp_path / name)
        r = caller.call_bams(ds["bams"], ds["fasta"], 4, out_vcf=p, threads=2, **kw)
        raw = open(p, "rb").read()
        if name.endswith(".gz"):
            _bgzf_blocks(raw)
            raw = gzip.decompress(raw)
        texts[name] = raw
        assert r["records"] == 2 and ("expand_stats" in r) == bool(kw)
    assert texts["a.vcf"] == texts["b.vcf.gz"] and texts["c.vcf.gz"] == texts["d.vcf"]
    assert texts["a.vcf"].endswith(rec) and texts["a.vcf"].startswith(b"##fileformat=VCF")
    assert texts["d.vcf"] == expand_genotypes(texts["a.vcf"], poolsize=4)[0]
    assert b"\tGT:GQ:DP\t0/0/0/1:33:60\t0/0/0/0:50:58\n" in texts["d.vcf"] and b"\t1/1/1/1:20:60\t./././.:0:0\n" in texts["d.vcf"]
    assert caller.call_bams(ds["bams"], ds["fasta"], 4, threads=2)["vcf_body"] == rec
    # a region: the decoder is asked for it, and the windows cover it and nothing else
    windows = []
    monkeypatch.setattr(caller, "make_window", lambda tid, start, end, ref, **kw: windows.append((start, end)) or (start, end))
    whole = caller.call_bams(ds["bams"], ds["fasta"], 4, threads=2, window=1000)
    assert windows == [(0, 1000), (1000, 2000), (2000, 3000)] and decoded[-1] == {}
    del windows[:]
    part = caller.call_bams(ds["bams"], ds["fasta"], 4, threads=2, window=1000, region=(500, 2200))
    assert windows == [(500, 1500), (1500, 2200)] and decoded[-1] == dict(start=500, end=2200)
    assert 0 < part["alignments"] < whole["alignments"]
