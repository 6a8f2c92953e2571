"""Golden vectors of the pooled-count -> genotype conversion from the reference's own scripts/convert_pooled_vcf.py.

The script is Python 2; it is run here through a syntactic translation applied in memory (print statements -> print(),
xrange -> range, backticks -> repr(), the unused `compiler` import dropped) — its logic is untouched and nothing of it is
stored in this repository.  Inputs: small pooled VCFs written below (every branch of the script: one pool size for all,
one per pool, missing genotypes, multi-allelic records, counts above the pool size) and, where oracle/_ref exists, the
VCF CRISP.binary writes for the synthetic BAM set of tests/frontend_io.py.

    python scripts/make_convert_golden.py        # -> tests/golden/convert_pooled/*.vcf, *.expected.vcf, *.args.json
"""
import json
import os
import re
import subprocess
import sys
import tempfile

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
REF_SCRIPT = "/root/reference/scripts/convert_pooled_vcf.py"
OUT = os.path.join(ROOT, "tests", "golden", "convert_pooled")


def translate_py2(src: str) -> str:
    """The reference script as Python 3 text: syntax only."""
    out = []
    for line in src.split("\n"):
        body = line.lstrip()
        if body.startswith("#") and not body.startswith("#!"):
            out.append(line)
            continue
        line = line.replace(", re, compiler", ", re").replace("xrange(", "range(")
        line = re.sub(r"`([^`]*)`", r"repr(\1)", line)
        m = re.search(r"print >>sys\.stderr,\s*", line)
        if m:
            rest = line[m.end():]
            cut, quote = len(rest), None   # the statement ends at the first ';' outside a string
            for k, ch in enumerate(rest):
                if quote:
                    if ch == quote and rest[k - 1] != "\\":
                        quote = None
                elif ch in "\"'":
                    quote = ch
                elif ch == ";":
                    cut = k
                    break
            line = line[:m.start()] + "print(" + rest[:cut].rstrip() + ", file=sys.stderr)" + rest[cut:]
        elif re.match(r"\s*print line,\s*$", line):
            line = line.replace("print line,", "print(line, end='')")
        else:
            m = re.match(r"(\s*)print (.*?)\s*$", line)
            if m:
                line = m.group(1) + "print(" + m.group(2) + ")"
        out.append(line)
    return "\n".join(out)


def run_reference(vcf_text: str, bam_list_text: str, poolsize):
    """(stdout, return code) of the reference script on the given inputs."""
    with tempfile.TemporaryDirectory() as d:
        open(os.path.join(d, "conv.py"), "w").write(translate_py2(open(REF_SCRIPT).read()))
        open(os.path.join(d, "in.vcf"), "w").write(vcf_text)
        open(os.path.join(d, "bams.txt"), "w").write(bam_list_text)
        cmd = [sys.executable, os.path.join(d, "conv.py"), os.path.join(d, "in.vcf"), os.path.join(d, "bams.txt")] + ([str(poolsize)] if poolsize else [])
        r = subprocess.run(cmd, capture_output=True, text=True)
        return r.stdout, r.returncode


HEADER = "##fileformat=VCFv4.0\n##source=CRISP\n#CHROM\tPOS\tID\tREF\tALT\tQUAL\tFILTER\tINFO\tFORMAT\tp0\tp1\tp2\n"


def cases():
    fmt = "AC:GQ:DP:ADf:ADr:ADb"
    fixed = HEADER + "".join([
        f"chr1\t100\t.\tA\tG\t312\tPASS\tNP=3;DP=10,12,0\t{fmt}\t1:33:60:28,2:29,1:0,0\t0:50:58:30,0:28,0:0,0\t3:12:70:30,4:33,3:0,0\n",
        f"chr1\t200\t.\tC\tT,G\t90\tPASS\tNP=3\t{fmt}\t1,1:20:60:1:2:3\t0,0:50:58:1:2:3\t.:0:0:.:.:.\n",
        f"chr1\t300\t.\tCA\tC,CAA\t80\tPASS\tNP=3\t{fmt}\t1,0:20:60:1:2:3\t0,1:50:58:1:2:3\t0,0:9:9:1:2:3\n",       # multi-allelic indel: left out
        f"chr1\t400\t.\tG\tA\t70\tLowDepth\tNP=3\t{fmt}\t5:20:60:1:2:3\t0:50:58:1:2:3\t0:9:9:1:2:3\n",                   # 5 > poolsize 4: left out
        f"chr1\t500\t.\tT\tC,A,G\t60\tPASS\tNP=3\t{fmt}\t1,1,2:20:60:1:2:3\t0,0,0:50:58:1:2:3\t4,0,0:9:9:1:2:3\n",
        f"chr1\t600\t.\tAT\tA\t55\tPASS\tNP=3\t{fmt}\t2:20:60:1:2:3\t.:0:0:.:.:.\t0:9:9:1:2:3\n",                       # bi-allelic indel: kept
        f"chr1\t700\t.\tA\tG\t50\tPASS\tNP=3\tAC\t1\t0\t2\n",                                                          # one FORMAT key only
    ])
    variable = HEADER + "".join([
        f"chr2\t100\t.\tA\tG\t312\tPASS\tNP=3\t{fmt}\t3,1:33:60:1:2:3\t6,0:50:58:1:2:3\t2,0:12:70:1:2:3\n",
        f"chr2\t200\t.\tC\tT,G\t90\tPASS\tNP=3\t{fmt}\t2,1,1:20:60:1:2:3\t.:0:0:.:.:.\t1,0,1:9:9:1:2:3\n",
        f"chr2\t300\t.\tG\tA\t70\tPASS\tNP=3\t{fmt}\t4,1:20:60:1:2:3\t6,0:50:58:1:2:3\t2,0:9:9:1:2:3\n",                 # 5 alleles in a pool of 4: left out
        f"chr2\t400\t.\tG\tA\t70\tPASS\tNP=3\t{fmt}\t1,1:20:60:1:2:3\t0,0:50:58:1:2:3\t2,0:9:9:1:2:3\n",                 # fewer alleles than the pool size: kept as they are
    ])
    out = {"fixed4": (fixed, "p0.bam\np1.bam\np2.bam\n", 4, None),
           "variable": (variable, "p0.bam PS=4\np1.bam PS=6\np2.bam PS=2\n", None, [4, 6, 2])}
    refbin = os.path.join(ROOT, "oracle", "_ref", "CRISP.binary")
    if os.path.exists(refbin):
        # a real pooled VCF: CRISP.binary on a small synthetic BAM set (6 pools x 20 haplotypes, 20 kb at 60x, some indels)
        sys.path.insert(0, ROOT)
        from crisp_b200 import frontend as fe
        from crisp_b200.bamsynth import make_dataset
        with tempfile.TemporaryDirectory() as d:
            ds = make_dataset(d, pools=6, poolsize=20, length=20000, depth=60, seed=7, indel_frac=0.2, pe_frac=0.3)
            vcf = os.path.join(d, "ref.vcf")
            fe.run_program(refbin, ds, vcf, 20, [])
            text = open(vcf).read()
        # the run's time stamp and paths are not part of what is checked
        text = "".join(l + "\n" for l in text.split("\n") if l and not (l.startswith("##") and ("Date" in l or "options" in l or "/" in l)))
        out["crisp_binary_pooled"] = (text, "".join(f"pool{i}.bam\n" for i in range(6)), 20, None)
    return out


def main():
    os.makedirs(OUT, exist_ok=True)
    for name, (vcf, bams, poolsize, sizes) in cases().items():
        want, rc = run_reference(vcf, bams, poolsize)
        assert rc == 0, (name, rc)
        open(os.path.join(OUT, name + ".vcf"), "w").write(vcf)
        open(os.path.join(OUT, name + ".expected.vcf"), "w").write(want)
        json.dump({"poolsize": poolsize, "pool_sizes": sizes}, open(os.path.join(OUT, name + ".args.json"), "w"))
        print(name, len(vcf.splitlines()), "lines ->", len(want.splitlines()))


if __name__ == "__main__":
    main()
