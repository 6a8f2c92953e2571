/*
 * vcf_post.h — what happens to the VCF text after the records are formatted:
 *
 *  - crisp_vcf_expand_genotypes: the conversion scripts/convert_pooled_vcf.py (lines 29-87) applies to a pooled CRISP VCF so
 *    that standard tools can read it — the per-pool allele counts in the first FORMAT field become a genotype of pool-size
 *    alleles ("0/0/0/1"), the key becomes GT, multi-allelic indels and records whose counts exceed the pool size are left
 *    out.  Text in, text out, as the script does it; the script's messages on stderr become counters.
 *  - crisp_bgzf_compress: the text as BGZF (samtools/bgzf.c: gzip members of at most 0xff00 input bytes with the BC extra
 *    field, closed by the empty end-of-file block), blocks deflated on several host threads — the bytes do not depend on
 *    the number of threads.
 *
 * Plain C (pthreads + zlib), no CUDA dependency.
 */
#ifndef CRISP_VCF_POST_H
#define CRISP_VCF_POST_H

#include <stddef.h>
#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

typedef struct crisp_vcf_expand_stats {
	int64_t n_header_lines;
	int64_t n_records_in, n_records_out;
	int64_t n_multiallelic_indel;   /* left out: two or more ALT alleles and REF differs in length from the first or second (script :37-39) */
	int64_t n_over_poolsize;        /* left out: the counts of a pool exceed its size (script :55-59, :75-78) */
} crisp_vcf_expand_stats;

/* poolsize > 0: every pool has this many alleles and the counts are those of the ALT alleles (script :48-63); otherwise
 * pool_sizes[n_pools] gives each pool's size and the counts include the reference allele first (script :64-79).
 * *out is malloc'ed (crisp_vcf_post_free).  Returns 0; -1 bad argument or a count that is not a number (the script stops with
 * a ValueError there), a record with more sample columns than n_pools in the variable-size mode; -4 out of memory. */
int crisp_vcf_expand_genotypes(const char* vcf, size_t len, int32_t poolsize, const int32_t* pool_sizes, int32_t n_pools, char** out, size_t* out_len,
                               crisp_vcf_expand_stats* stats);

/* level: zlib's 0-9 (-1: default).  *out is malloc'ed (crisp_vcf_post_free).  Returns 0; -1 bad argument, -4 out of memory,
 * -6 zlib failure. */
int crisp_bgzf_compress(const void* data, size_t len, int level, int threads, unsigned char** out, size_t* out_len);

void crisp_vcf_post_free(void* p);

#ifdef __cplusplus
}
#endif
#endif
