// engine_types.cuh — device-side views of the batch, the configuration and the results (plain pointers, SoA).
#pragma once
#include <cstdint>

#include "../../include/crispgpu.h"

namespace crisp {

constexpr int kNA = 8;          // maxalleles
constexpr int kNC = 24;         // 3 strand classes x 8 alleles
constexpr int kSlots = 5;
constexpr int kICPad = 25;      // padded row of the staged per-pool counts (bank-conflict-free with lanes over pools)
constexpr int kNQ = 128;        // quality characters 0..127 index the likelihood table
constexpr int kDenseNQ = 16;     // most distinct qualities the dense likelihood path handles
constexpr int kNcrTabN = 256;   // rows of the host-computed log10 C(n,r) table
constexpr double kExpCut = 20.0;  // terms below 10^-20 of the running maximum cannot change a double sum

struct DevBatch {
	int n_sites;
	int P;
	const int32_t* tid;
	const int32_t* pos;
	const uint8_t* refbase;
	const uint8_t* maxvec_al;
	const int32_t* maxvec_ct;
	const int32_t* counts;
	const int32_t* indcounts;
	const int16_t* indel_len;
	const int16_t* hplen;
	const uint32_t* read_off;
	const uint16_t* reads;
	const uint32_t* alt_off;
	const crispgpu_altread* alt_reads;
	const int32_t* amb_idx;
	const int32_t* amb_dec;
	const double* amb_ep;
	const double* qhigh;
	const double* stats;
	const double* tstats;
	const uint32_t* bin_off;
	const uint32_t* bins;
	const uint32_t* ic_off;
	const uint32_t* ic_sparse;
};

struct DevCfg {
	int P, K, nmax, n_pclass, ploidy0, varpoolsize;
	int em_flag, max_iter, chisq_permutation, fast_filter, use_base_qvs, min_reads, qv_offset, min_q, pivot_sample;
	int use_qv, exact_mode, association;
	double thresh1, min_qpv, thresh3, ser, relax, qmin, ref_bias, theta, alpha_prior;  // alpha_prior = theta * H_K
	uint64_t rng_seed;
	const int32_t* ploidy;   // [P]
	const int32_t* pclass;   // [P] index of the pool's ploidy among the distinct ploidies
	const int32_t* phenotypes;  // [P] bit mask 1 control / 2 case / 4 early onset (options->phenotypes), NULL without --phenotypes
	const double* NC;        // [n_pclass][nmax+1]  log10 C(n,j)  (parsebam/init_variant.c:19-25, computed on the host)
	const double* T;         // [n_pclass][2][kNQ][nmax+1] log10 e / log10(1-e) per (quality, allele count)
	// distinct quality characters of the batch (k_qualmap): when there are at most kDenseNQ of them and the pools share one
	// size, k_em uses the dense likelihood path with the table rows staged in shared memory
	int nq, dense_lik;
	uint8_t qlist[16];
	const double* ncr_tab;   // [kNcrTabN][kNcrTabN/2+1] log10 C(n,r), r <= n/2, host-computed in the reference's summation
	                         // order (FET/contables.c:80-87) so that the epsilon-free comparisons of lowcovFET.c:132-133
	                         // see bit-identical terms; NULL unless the low-coverage permutation test is enabled
};

struct DevOut {
	int32_t* varalleles;
	int32_t* strandpass;
	uint8_t* status;
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
	int32_t* ac;   // [rows][P]
	double* qv;    // [rows][P]
};

// work queues filled on the device
struct DevQueues {
	int32_t* perm_tasks;   // site*5+slot
	int32_t* call_tasks;   // site*5+slot that reached the caller: quality-value (SNP) likelihood tasks, or every task with --EM 0
	int32_t* call_tasks2;  // EM tasks whose likelihood is count-based (indel alleles, --usebq 0)
	uint8_t* need_reads;   // [n_sites] or null: site has a quality-value likelihood task (its reads are fetched from host memory on demand)
	int32_t* counters;     // [14..15]=read elements fetched on demand (64-bit); [0]=n_perm [1]=n_call [2]=perm head [3]=call head [4..7]=quality-presence bitmap (128 bits) [8]=n_call2 [9]=call2 head [10]=some read is bidirectional
};

struct EngineParams {
	DevBatch b;
	DevCfg c;
	DevOut o;
	DevQueues q;
	double* gws;          // global workspace for genotype likelihoods when they do not fit in shared memory
	size_t gws_stride;    // doubles per CTA
	long long* em_trace;  // measurement switch (CRISPGPU_EM_TRACE): per EM task eight words: ns at task start / counts staged / histogram built / planes built / first iteration done / EM done / task end, then SM | task << 16 | (2 iterations + called) << 48; null in normal runs
};

}  // namespace crisp
