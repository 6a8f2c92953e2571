"""TEST INFRASTRUCTURE — restatement of the reference's scripts/convert_pooled_vcf.py (Python 2; lines cited), the checker of
host/vcf_post.c::crisp_vcf_expand_genotypes.  Pinned: tests/test_vcf_post.py runs the reference's own script (translated to
Python 3 by lib2to3 on the fly, where /root/reference exists) on the same inputs, and tests/golden/convert_pooled/ holds its
outputs (scripts/make_convert_golden.py) for where it does not.  Only tests may import this module."""
from __future__ import annotations


def convert(vcf_text: str, poolsize: int = -1, pool_sizes=None) -> tuple[str, dict]:
    pool_sizes = list(pool_sizes or [])
    out = []
    stats = dict(n_header_lines=0, n_records_in=0, n_records_out=0, n_multiallelic_indel=0, n_over_poolsize=0)
    for line in vcf_text.split("\n"):
        if line == "":
            continue                                    # (iteration over a file never yields an empty string)
        if line[0] == "#":                              # :31-33
            out.append(line)
            stats["n_header_lines"] += 1
            continue
        stats["n_records_in"] += 1
        var = line.strip().split("\t")                  # :35
        alleles = var[4].split(",")
        if len(alleles) >= 2 and (len(var[3]) != len(alleles[0]) or len(var[3]) != len(alleles[1])):   # :37-39
            stats["n_multiallelic_indel"] += 1
            continue
        valid = True
        genotypes = []
        for i in range(9, len(var)):                    # :43
            genotype = var[i].split(":")
            counts = genotype[0].split(",")
            gv = []
            if poolsize > 0:                            # :46-62
                refsum = poolsize
                if genotype[0] == ".":
                    gv = ["."] * poolsize
                else:
                    for c in counts:
                        refsum -= int(c)
                    if refsum < 0:
                        valid = False
                        break
                    gv = ["0"] * refsum
                    for a, c in enumerate(counts):
                        gv += [str(a + 1)] * int(c)
            else:                                       # :64-78
                if genotype[0] == ".":
                    gv = ["."] * pool_sizes[i - 9]
                else:
                    for a, c in enumerate(counts):
                        gv += [str(a)] * int(c)
                    if len(gv) > pool_sizes[i - 9]:
                        valid = False
                        break
            genotypes.append("/".join(gv) + ":" + ":".join(genotype[1:]))   # :82
        if not valid:                                   # :80
            stats["n_over_poolsize"] += 1
            continue
        new_col8 = "GT:" + ":".join(var[8].split(":")[1:])                 # :81
        out.append("\t".join(var[0:8] + [new_col8] + genotypes).strip())    # :83
        stats["n_records_out"] += 1
    return "".join(x + "\n" for x in out), stats
