/*
 * vcf_writer.h — batch VCF record writer: the text print_pooledvariant
 * (crisp/newcrispprint.c:347-476, with print_VCF_first_5cols :25-96, print_VCF_allelefreqs :99-131,
 * print_VCF_genotypes_diploid :134-180, print_VCF_genotypes_pooled :248-298, print_flanking_sequence :301-345) writes for
 * the called sites of a batch, formatted from the engine's flat result arrays on several host threads into one buffer in
 * coordinate order — byte for byte the reference's records.  Plain C (pthreads), no CUDA dependency.
 */
#ifndef CRISP_VCF_WRITER_H
#define CRISP_VCF_WRITER_H

#include <stddef.h>
#include <stdint.h>

#include "../include/crispgpu.h"

#ifdef __cplusplus
extern "C" {
#endif

/* The candidate sites of the batch as the front end (or crispgpu_pileup_run: crispgpu_pileup_sites) knows them. */
typedef struct crisp_vcf_sites {
	int32_t n_sites, n_pools;
	const char* chrom;           /* variant->chrom */
	const int32_t* pos;          /* [n] 0-based */
	const uint8_t* refbase;      /* [n] variant->refbase */
	const int32_t* counts;       /* [n][24] */
	const int32_t* indcounts;    /* [n][P][24] */
	const int32_t* readdepths;   /* [n][P] */
	const int32_t* mqcounts;     /* [n][4] */
	const char* refseq;          /* reference characters from position ref_start on (flanking sequence, REF column) */
	int32_t ref_start, ref_len;
	int64_t chrom_len;           /* reflist->lengths[current] */
} crisp_vcf_sites;

typedef struct crisp_vcf_options {
	const int32_t* ploidy;       /* [P] */
	int32_t em_flag;             /* EMflag */
	int32_t varpoolsize;         /* variant->varpoolsize */
	int32_t threads;
	int32_t reserved;
} crisp_vcf_options;

/* Records of every site with results->varalleles >= 1, in site order.  *out is malloc'ed (crisp_vcf_free).  Returns 0;
 * -1 bad argument / inconsistent results, -4 out of memory, -5 a called insertion/deletion allele (not in scope). */
int crisp_vcf_format(const crisp_vcf_sites* sites, const crispgpu_results* results, const crisp_vcf_options* opt, char** out, size_t* out_len,
                     int32_t* n_records);
void crisp_vcf_free(char* p);

#ifdef __cplusplus
}
#endif
#endif
