/*
 * oracle_abi.h — TEST INFRASTRUCTURE ONLY.
 *
 * Caller-allocated, writable mirror of `crispgpu_results` (include/crispgpu.h) used by the two
 * CPU checkers:
 *   oracle/crisp_oracle.c   the plain-C restatement of the reference algorithm, and
 *   oracle/ref_harness.c    the UNMODIFIED reference compiled from /root/reference (oracle/_ref/).
 * Only tests/, __graft_entry__.smoke() and bench.py's cpu_baseline / --impl reference leg may
 * load these; the product (crisp_b200/) never does.
 */
#ifndef CRISP_ORACLE_ABI_H
#define CRISP_ORACLE_ABI_H
#include "../include/crispgpu.h"

typedef struct crisp_host_results {
	int32_t n_sites;
	int32_t n_calls;      /* out */
	int32_t max_calls;    /* in: capacity (rows) of ac/qv */
	int32_t reserved;
	int32_t* varalleles;  /* [n] */
	int32_t* strandpass;  /* [n] */
	uint8_t* status;      /* [n][5] */
	uint8_t* allele;
	int8_t* call_rank;
	int32_t* call_row;
	double* paf;
	double* ctpval;
	double* chi2pval;
	double* af_ml;
	double* delta;
	double* deltaf;
	double* hwe;
	double* qvpf;
	double* qvpr;
	int32_t* varpools;
	int32_t* allelecount;
	int32_t* variantpools;
	int32_t* em_iters;
	int32_t* perm_k;
	int32_t* perm_hits;
	double* assoc_p;
	double* assoc_llr;
	double* assoc_or;
	double* assoc_p1;
	double* assoc_llr1;
	double* assoc_or1;
	int32_t* ac;          /* [max_calls][P] */
	double* qv;           /* [max_calls][P] */
	/* optional per-(site,slot) dumps for kernel-level parity (may be NULL):
	 * GENLL, GENLLf, GENLLr, GENLLb of the slot, each [n][5][P][maxploidy+1] */
	double* genll;        /* [n][5][4][P][maxploidy+1] or NULL */
	int32_t genll_stride; /* maxploidy+1 */
	int32_t reserved2;
} crisp_host_results;

/* RNG used for the permutation tests */
enum {
	CRISP_RNG_DRAND48 = 0, /* the reference's generator (FET/contables.c:314), srand48(cfg->rng_seed) once per batch */
	CRISP_RNG_PHILOX = 1   /* the GPU engine's counter-based sampler, restated serially */
};

#endif
