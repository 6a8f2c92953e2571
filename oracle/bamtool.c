/*
 * bamtool.c — TEST INFRASTRUCTURE.  `bamtool index x.bam` writes x.bam.bai with the vendored samtools 0.1.18 library
 * (bam_index_build, /root/reference/samtools/bam_index.c), so that the reference's `--regions` (bamsreader.c:147-173)
 * can be exercised on the synthetic BAMs written by scripts/make_bams.py.
 */
#include <stdio.h>
#include <string.h>
#include "bam.h"

int main(int argc, char** argv)
{
	if (argc == 3 && strcmp(argv[1], "index") == 0) return bam_index_build(argv[2]) < 0;
	fprintf(stderr, "usage: bamtool index file.bam\n");
	return 2;
}
