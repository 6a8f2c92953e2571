/*
 * bam_decode.h — BGZF/BAM decode for the device pileup.
 *
 * Replaces, for the reads the device pileup consumes, what the reference does per alignment with the vendored samtools
 * reader and get_read_bamfile (samtools/bgzf.c; parsebam/bamread.c:80-152; the flag filter of readmultiplebams.c:121):
 * the BGZF blocks of every pool's BAM are inflated in parallel by host threads (zlib, one block per work item), the
 * records are split and written straight into the struct-of-arrays `crispgpu_alignments` the pileup kernels read — bases
 * stay in BAM's 4-bit code and qualities in BAM's bytes, so a record is moved with two memcpy calls and never expanded
 * to text.  Plain C (pthreads + zlib); no CUDA dependency: the caller passes the allocator (crispgpu_host_alloc for
 * pinned memory).
 */
#ifndef CRISP_BAM_DECODE_H
#define CRISP_BAM_DECODE_H

#include <stddef.h>
#include <stdint.h>

#include "../include/crispgpu.h"

#ifdef __cplusplus
extern "C" {
#endif

typedef struct crisp_bam_options {
	int32_t tid;            /* keep alignments on this reference id; < 0: every reference (then start/end are ignored) */
	int32_t start, end;     /* keep alignments whose match block overlaps [start, end) */
	uint32_t filter_mask;   /* BAM_FILTER_MASK: reads with (flag & mask) != 0 are dropped (0x704 = unmapped|secondary|qcfail|dup) */
	int32_t threads;        /* host threads (<= 0: one) */
	int32_t paired;         /* 1: also fill mate_pos / isize */
	int32_t reuse;          /* 1: `out` holds the arrays of an earlier call (same allocator): they are used again when large
	                           enough — pinning memory costs more than decoding into it — and whatever viewed them is overwritten */
	int32_t use_index;      /* 1: with a region (tid >= 0 and start/end narrower than everything) and a BAI index next to a file
	                           (x.bam.bai or x.bai), only the blocks the index names for the region are inflated; without an
	                           index, or with one that does not fit the file, the whole file is decoded and filtered as before */
	int32_t reserved;
	void* (*alloc)(size_t);     /* allocator of the output arrays (NULL: malloc) */
	void (*release)(void*);     /* its counterpart (NULL: free) */
} crisp_bam_options;

typedef struct crisp_bam_reads {
	crispgpu_alignments aln;    /* what crispgpu_pileup_run takes; arrays owned by this struct */
	int64_t n_records;          /* alignment records seen in the files */
	int64_t n_flag_filtered;    /* dropped by filter_mask */
	int64_t n_outside;          /* dropped: other reference / no overlap with [start, end) */
	int64_t n_complex;          /* dropped: CIGAR is not [H][S] M [S][H] (insertions, deletions, skips, pads) — the caller must
	                               route such data through the host pileup; crispgpu_pileup_run would otherwise miss them */
	int64_t compressed_bytes, inflated_bytes;
	double read_s, inflate_s, parse_s;  /* wall seconds: file read, parallel inflate, record split + packing */
	int32_t n_refs;
	int32_t first_ref_len;
	char first_ref_name[64];
	uint64_t cap_reads, cap_bases;  /* capacity of the arrays (alignments, bases): what a later call with opt->reuse can take over */
	int32_t cap_pools, cap_paired;
} crisp_bam_reads;

/* Decode n_pools coordinate-sorted BAM files (one per pool) into one alignment set.  Returns 0, or a negative number:
 * -1 bad argument, -2 I/O error, -3 not a BGZF/BAM file or corrupt block, -4 out of memory. */
int crisp_bam_decode(const char* const* paths, int32_t n_pools, const crisp_bam_options* opt, crisp_bam_reads* out);
void crisp_bam_release(crisp_bam_reads* r, void (*release)(void*));

#ifdef __cplusplus
}
#endif
#endif
