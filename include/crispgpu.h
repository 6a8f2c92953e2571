/*
 * crispgpu.h — C ABI of the B200 statistics engine for CRISP candidate sites.
 *
 * This is the drop-in boundary.  The reference (vibansal/crisp) has no plugin /
 * FFI interface: its per-site statistics engine is entered by one direct call,
 *     v = newCRISPcaller(reflist,current,position,bq,bamfiles_data,variant,maxvec,vfile)
 *                                                   (variantcalls.c:115, prototype crisp/crispcaller.h:21)
 * which reads `struct VARIANT` (parsebam/variant.h:50-95) plus the live read queue and
 * writes its outputs back into `struct VARIANT` before print_pooledvariant()
 * (crisp/newcrispprint.c:347) formats the VCF record.  The functions below replace that
 * call for a *batch* of candidate sites: the host front end (parsebam/, unchanged C)
 * packs the fields of `struct VARIANT` and the three read-queue snapshots the engine
 * needs into the struct-of-arrays `crispgpu_batch`, the GPU computes everything
 * newCRISPcaller would have computed, and `crispgpu_results` returns exactly the fields
 * print_pooledvariant() consumes (newcrispprint.c:357-473, 248-298).
 *
 * Plain C, plain pointers and sizes; no C++ or torch types.  All functions return 0 on
 * success or a negative crispgpu_status; no exceptions, nothing is written to stdout.
 * There is no CPU fallback: if no CUDA device is usable crispgpu_create() fails.
 */
#ifndef CRISPGPU_H
#define CRISPGPU_H

#include <stdint.h>
#include <stddef.h>

#ifdef __cplusplus
extern "C" {
#endif

#define CRISPGPU_ABI_VERSION 3  /* 2: compact batch arrays (bins, sparse counts), crispgpu_advance, fetched_bytes
                                   3: crispgpu_pileup_run, config.minimal_results, timing.pileup_ms, crispgpu_measure_int32_tops, narrow transport arrays */

#define CRISPGPU_MAXALLELES 8  /* `maxalleles`, readmultiplebams.c:27: A,C,G,T + indel slots 4..6 */
#define CRISPGPU_NCOUNT 24     /* 3 strand classes x 8 alleles: index = allele + s*8, s=0 fwd,1 rev,2 bidirectional
                                  (parsebam/variant.c:301-312) */
#define CRISPGPU_MAXSLOTS 5    /* alleles maxvec[1..5] are tested, newcrispcaller.c:205 */
#define CRISPGPU_NCHISTAT 5

typedef enum crispgpu_status {
	CRISPGPU_OK = 0,
	CRISPGPU_ERR_ARG = -1,       /* NULL pointer, bad size, inconsistent offsets */
	CRISPGPU_ERR_NO_DEVICE = -2, /* no usable CUDA device (there is no CPU fallback) */
	CRISPGPU_ERR_CUDA = -3,      /* CUDA runtime error; see crispgpu_last_error() */
	CRISPGPU_ERR_NOMEM = -4,
	CRISPGPU_ERR_UNSUPPORTED = -5, /* configuration needs an input array the batch does not carry */
	CRISPGPU_ERR_TICKET = -6
} crispgpu_status;

/* Snapshot of the reference's global option variables that reach the engine
 * (optionparser.c:17-136).  Field comments name the reference global. */
typedef struct crispgpu_config {
	int32_t abi_version;       /* must be CRISPGPU_ABI_VERSION */
	int32_t n_pools;           /* variant->samples */
	const int32_t* ploidy;     /* variant->ploidy[n_pools]  (--poolsize / PS=)  init_variant.c:135 */
	int32_t varpoolsize;       /* variant->varpoolsize: 1 if pool sizes differ (HWE skipped, crispEM.c:187) */
	int32_t em_flag;           /* EMflag (--EM), newcrispcaller.c:32 : 1 = EM caller, 0 = legacy caller */
	int32_t max_iter;          /* MAXITER (--perms), newcrispcaller.c:21 */
	int32_t chisq_permutation; /* CHISQ_PERMUTATION (--chisquare 0 sets 1), FET/contables.c:11 */
	int32_t fast_filter;       /* FAST_FILTER, newcrispcaller.c:24 */
	int32_t use_base_qvs;      /* USE_BASE_QVS (--usebq), newcrispcaller.c:20 */
	int32_t min_reads;         /* MIN_READS (--minc), newcrispcaller.c:23 */
	int32_t qv_offset;         /* QVoffset, readmultiplebams.c:25 */
	int32_t min_q;             /* MINQ (--mbq), readmultiplebams.c:26 */
	int32_t pivot_sample;      /* PIVOTSAMPLE (--pivotsample), readmultiplebams.c:46 */
	int32_t association;       /* options->association (--phenotypes given), crispEM.c:383 */
	const int32_t* phenotypes; /* options->phenotypes[n_pools] bit mask 1 control/2 case/4 early; may be NULL */
	double thresh1;            /* --ctpval  (-3.5), newcrispcaller.c:22 */
	double min_qpv;            /* MINQpv, --qvpval (-5) */
	double thresh3;            /* thresh3 (-2) */
	double ser;                /* SER (-e) (-1) */
	double relax;              /* RELAX (0.75) */
	double qmin;               /* Qmin (23) */
	double agilent_ref_bias;   /* AGILENT_REF_BIAS (--refbias) 0.5, crispEM.c:26 */
	double theta;              /* THETA 0.001, crispEM.c:14 */
	/* engine-only switches (no reference counterpart) */
	uint64_t rng_seed;         /* key of the counter-based RNG; results are a pure function of
	                              (rng_seed, tid, pos, slot), independent of batching and GPU count */
	int32_t exact_mode;        /* 1 (with chisq_permutation=1, pooled): tables with 2 <= pools < 8, < 70 alt reads and
	                              columns < 2001 reads get compute_pvalue_exact (FET/contables_exact.c:62-117) on the
	                              unstranded table instead of the Monte-Carlo estimate; reported with perm_k = 0, perm_hits = -1 */
	int32_t use_qv;            /* USE_QV (--weighted), readmultiplebams.c:24: how remove_ambiguous_bases adjusts Qhighcounts */
	int32_t blocking_wait;     /* 1: crispgpu_wait sleeps on the result event instead of spinning a core — for front ends
	                              that keep more contexts waiting than the host has idle cores (adds wake-up latency) */
	int32_t minimal_results;   /* 1: crispgpu_results carries only what print_pooledvariant reads; the diagnostic fields (paf, chi2pval,
	                              em_iters) are not copied back and read as 0.  Fields no configuration of this context can
	                              produce (assoc_* without --phenotypes, qvpf/qvpr/strandpass with --EM 1, perm_k/perm_hits
	                              without --chisquare 0) are never copied and always read as 0. */
} crispgpu_config;

/* One packed read of the per-site pileup snapshot: what calculate_pooled_likelihoods_snp
 * (crisp/crispEM.c:104-136) reads from `struct alignedread`.  Only reads with
 * filter=='0' && type==0 are packed (others contribute nothing there).
 *   bits 0-7   quality char  bcall->quality[l1+delta]   (post-OPE value, still +33, bamread.c:124)
 *   bits 8-9   BTI[bcall->sequence[l1+delta]]  0..3
 *   bit  10    bcall->strand != 0
 *   bit  11    bcall->bidir == 1
 * Reads are grouped by pool: reads of (site s, pool i) are read_off[s*P+i] .. read_off[s*P+i+1]. */
typedef uint16_t crispgpu_read;
#define CRISPGPU_READ(qualchar, base, rev, bidir) \
	((uint16_t)(((qualchar) & 0xff) | (((base) & 3) << 8) | (((rev) ? 1 : 0) << 10) | (((bidir) ? 1 : 0) << 11)))

/* (value, count) pair of the compact read format: low 16 bits = a crispgpu_read value, high 16 bits = how many reads
 * of the cell carry exactly that value (1..65535; a value may appear in several bins when it occurs more often). */
typedef uint32_t crispgpu_bin;
#define CRISPGPU_BIN(read_value, count) ((crispgpu_bin)(((uint32_t)(count) << 16) | (uint32_t)(uint16_t)(read_value)))

/* One non-zero entry of a pool's indcounts row: bits 27..31 = index (allele + 8*strand class, 0..23), bits 0..26 = count. */
typedef uint32_t crispgpu_count;
#define CRISPGPU_COUNT(index, count) ((crispgpu_count)(((uint32_t)(index) << 27) | ((uint32_t)(count) & 0x7ffffffu)))

/* One alternate-allele read as calculate_uniquereads (parsebam/allelecounts.c:50-77) sees it. */
typedef struct crispgpu_altread {
	uint16_t pool;    /* sample */
	uint16_t offset;  /* bcall->l1 + bcall->delta */
	uint16_t aligned; /* bcall->alignedbases */
	uint8_t strand;   /* bcall->strand */
	uint8_t pad;
} crispgpu_altread;

/* Struct-of-arrays batch of n_sites candidate sites (sites that passed variantcalls.c:76-92).
 * P = config.n_pools.  All arrays are caller-owned host memory (pinned memory from
 * crispgpu_host_alloc makes the copies asynchronous) and must stay valid until crispgpu_wait(). */
typedef struct crispgpu_batch {
	int32_t n_sites;
	int32_t reserved;
	const int32_t* tid;         /* [n]     reference id  (pass-through key, RNG key) */
	const int32_t* pos;         /* [n]     variant->position (0-based) */
	const uint8_t* refbase;     /* [n]     variant->refbase 0..3 */
	const uint8_t* maxvec_al;   /* [n][8]  maxvec[k].al after sort_allelecounts (allelecounts.c:155); NULL (with maxvec_ct): sorted on the device */
	const int32_t* maxvec_ct;   /* [n][8]  maxvec[k].ct */
	const int32_t* counts;      /* [n][24] variant->counts */
	const int32_t* indcounts;   /* [n][P][24] variant->indcounts[i][0..3A) */
	const int16_t* indel_len;   /* [n][8]  alleles>=4: +len for "+XX", -len for "-XX" (strlen(itb[a])-1), else 0; may be NULL if no indels */
	const int16_t* hplen;       /* [n][8]  variant->HPlength[a] (calculate_indel_ambiguity); may be NULL */
	const uint32_t* read_off;   /* [n*P+1] */
	const crispgpu_read* reads; /* [read_off[n*P]] */
	/* calculate_uniquereads inputs: for (site s, slot k=al-1): alt_off[s*5+k] .. alt_off[s*5+k+1],
	 * entries sorted by pool, at most (alt read count of the pool) entries per pool. */
	const uint32_t* alt_off;    /* [n*5+1] */
	const crispgpu_altread* alt_reads;
	/* what remove_ambiguous_bases(bq,variant,HPlength[allele2],+1) (process_indel_variant.c:20-66) subtracts
	 * from the reference-base counts before an indel allele is evaluated: amb_idx[s*5+k] = -1 (none) or row r,
	 * amb_dec[r][i][0..3] = number of removed reference reads of pool i that are {forward, reverse,
	 * bidirectional with strand 0, bidirectional with strand 1} (:39-48: the first two come off
	 * indcounts[i][refbase], [refbase+8], the last two off [refbase+16]; Qhighcounts/stats come off by strand, :57-60).
	 * amb_ep[r][i][0..1] = sum of 10^(-q/10) over the removed reads of strand 0 / 1 (only read when
	 * stats or weighted Qhighcounts are in use).  May be NULL. */
	const int32_t* amb_idx;     /* [n][5] */
	const int32_t* amb_dec;     /* [n_amb][P][4] */
	const double* amb_ep;       /* [n_amb][P][2] or NULL */
	int32_t n_amb;
	int32_t reserved2;
	/* optional real-valued pileup sums (NULL unless needed):
	 *   qhigh  variant->Qhighcounts[i][0..2A)  read when chisq_permutation=1 (chi2pvalue, FET/chisquare.c:16);
	 *          NULL means Qhighcounts[i][a+8s] = indcounts[i][a+8s], s=0,1 (exact when there are no
	 *          bidirectional reads and --weighted is off, parsebam/variant.c:318-320)
	 *   stats  variant->stats[i][0..2A), tstats variant->tstats[0..2A)  needed when em_flag=0 (chernoff_bound);
	 *          NULL means zeros */
	const double* qhigh;        /* [n][P][16] */
	const double* stats;        /* [n][P][16] */
	const double* tstats;       /* [n][16] */
	/* Compact alternative to `reads` (the likelihood depends on a pool's reads only through how many carry each
	 * (quality, base, strand, bidirectional) value): cell c = s*P+i owns bins[bin_off[c] .. bin_off[c+1]), any order.
	 * When `bins` is non-NULL `reads` and `read_off` are ignored and may be NULL.  The bins are the likelihood's sufficient statistic: k_em adds them into its
	 * histogram (dense path) or multiplies them with the table rows directly (sparse path); no read stream is rebuilt.
	 * Results are those of the equivalent `reads` batch. */
	const uint32_t* bin_off;    /* [n*P+1] or NULL */
	const crispgpu_bin* bins;   /* [bin_off[n*P]] or NULL */
	/* Compact alternative to `indcounts` (a pool's 24-entry row is mostly zeros): cell c owns
	 * ic_sparse[ic_off[c] .. ic_off[c+1]), the non-zero entries of variant->indcounts[i] in any order.  When
	 * `ic_sparse` (or an all-empty `ic_off`) is given `indcounts` is ignored and may be NULL; the engine scatters the
	 * entries into a zeroed dense table on the device (k_expand_counts). */
	const uint32_t* ic_off;          /* [n*P+1] or NULL */
	const crispgpu_count* ic_sparse; /* [ic_off[n*P]] or NULL */
	/* Optional: the front end packed every array above into ONE host allocation (pinned, from crispgpu_host_alloc):
	 * [arena_base, arena_base + arena_bytes), arena_base 256-byte aligned, every array 16-byte aligned inside it.  The
	 * engine then moves the batch with a single copy and addresses the arrays by their offsets (a dozen copies
	 * otherwise — what matters once several front-end threads share the device's driver lock).  Every non-NULL array
	 * must lie inside the arena (CRISPGPU_ERR_ARG otherwise).  NULL / 0: arrays are copied one by one. */
	const void* arena_base;
	uint64_t arena_bytes;
	/* Narrow transport form of the two compact arrays (what crosses PCIe when the host link is the bound — eight ranks of
	 * a box share it): 16-bit entries and one length byte per cell instead of 32-bit entries and 32-bit offsets.  Built
	 * from a compact batch by crispgpu_batch_pack_narrow; when `bins16` / `ic16` are non-NULL the corresponding wide
	 * arrays above are ignored and may be NULL.  The engine widens them on the device (k_unpack_bins16,
	 * k_expand_counts16) into exactly the arrays the compact batch would have produced, so results are identical.
	 *   bins16[k]  bits 0-6 quality char, 7-8 base, 9 strand, 10 bidirectional (crispgpu_read without its unused bit 7),
	 *              bits 11-15 count - 1 (a bin of more than 32 reads takes several entries)
	 *   bin_len[c] entries of cell c = s*P+i (a cell with more than 255 entries cannot be packed: pack_narrow refuses)
	 *   bin_site[s] first entry of site s, bin_site[n] = total
	 *   ic16[k]    bits 11-15 index (allele + 8*strand class), bits 0-10 count (a count above 2047 takes several ADJACENT entries of one index, which add up)
	 *   ic_len[c], ic_site[s] as above */
	const uint16_t* bins16;
	const uint8_t* bin_len;      /* [n*P] */
	const uint32_t* bin_site;    /* [n+1] */
	const uint16_t* ic16;
	const uint8_t* ic_len;       /* [n*P] */
	const uint32_t* ic_site;     /* [n+1] */
	/* In a narrow batch `counts` may be NULL (crispgpu_batch_pack_narrow leaves it out): the site totals are the column sums
	 * of the per-pool counts (k_expand_counts16); `maxvec_al`/`maxvec_ct` may be NULL as in any batch.  `alt32`, when non-NULL,
	 * replaces `alt_reads` (same alt_off): pool in bits 0-10, read offset in bits 11-20, aligned bases in bits 21-30, strand
	 * in bit 31 — packed by crispgpu_batch_pack_narrow when every record fits (reads shorter than 1024 bases). */
	const uint32_t* alt32;
} crispgpu_batch;

#define CRISPGPU_BIN16(read_value, count) ((uint16_t)(((read_value) & 0x7f) | ((((read_value) >> 8) & 0xf) << 7) | ((((count) - 1) & 0x1f) << 11)))
#define CRISPGPU_COUNT16(index, count) ((uint16_t)((((index) & 0x1f) << 11) | ((count) & 0x7ff)))
#define CRISPGPU_ALT32(pool, offset, aligned, strand) \
	((uint32_t)(((pool) & 0x7ffu) | (((offset) & 0x3ffu) << 11) | (((aligned) & 0x3ffu) << 21) | (((strand) ? 1u : 0u) << 31)))

/* Per tested allele slot (al = slot+1 of maxvec), state of the reference's per-allele loop
 * (newcrispcaller.c:205-246). */
enum {
	CRISPGPU_SLOT_UNTESTED = 0,  /* maxvec[al].ct < 3 */
	CRISPGPU_SLOT_REJECT_AF = 1, /* evaluate_variant returned 0 at the p[0] filter (newcrispcaller.c:72) */
	CRISPGPU_SLOT_REJECT_CT = 2, /* evaluate_variant returned 0 at the homopolymer / FAST_FILTER tests (:108-127) */
	CRISPGPU_SLOT_NOCALL = 3,    /* reached the caller (EM or legacy) but was not called a variant */
	CRISPGPU_SLOT_CALLED = 4     /* variant allele: occupies index `call_rank` of the site's varalleles list */
};

/* Results for the batch.  Arrays are owned by the context (pinned host memory) and stay valid
 * until the next crispgpu_submit()/crispgpu_destroy() on that context. */
typedef struct crispgpu_results {
	int32_t n_sites;
	int32_t n_calls;            /* total called alleles in the batch = rows of ac/qv */
	const int32_t* varalleles;  /* [n]    variant->varalleles : >=1  <=> a VCF record is printed */
	const int32_t* strandpass;  /* [n]    variant->strandpass (EM 0 only) */
	/* per (site, slot) — [n][5] */
	const uint8_t* status;      /* CRISPGPU_SLOT_* */
	const uint8_t* allele;      /* allele2 = maxvec[al].al */
	const int8_t* call_rank;    /* index into variant->alleles[]/crispvar[] or -1 */
	const int32_t* call_row;    /* row of ac/qv or -1 */
	const double* paf;          /* p[0] initial AF estimate of evaluate_variant (x ploidy sum) */
	const double* ctpval;       /* variant->ctpval[]   (VCF CT=) */
	const double* chi2pval;     /* variant->chi2pval[] */
	const double* af_ml;        /* crispvar[].AF_ML   (VCF AF=) */
	const double* delta;        /* crispvar[].delta   (VCF QUAL, EMstats) */
	const double* deltaf;       /* crispvar[].deltaf  (EMstats, VF=) */
	const double* hwe;          /* crispvar[].deltar = log10 HWE p (HWEstats) */
	const double* qvpf;         /* variant->qvpvaluef[] (EM 0) */
	const double* qvpr;         /* variant->qvpvaluer[] (EM 0) */
	const int32_t* varpools;    /* variant->varpools[] = nvariants[] (VCF VP=) */
	const int32_t* allelecount; /* crispvar[].allelecount (VCF AC=) */
	const int32_t* variantpools;/* crispvar[].variantpools */
	const int32_t* em_iters;    /* EM iterations executed (diagnostic) */
	const int32_t* perm_k;      /* permutations executed (stop index) (diagnostic) */
	const int32_t* perm_hits;   /* hits at stop (diagnostic) */
	/* case/control likelihood-ratio association statistic (crisp/LRstatistic.c:48-69; the reference prints these on its
	 * `ASSOC` stdout line only).  Filled when config.association = 1 and the slot satisfies crispEM.c:383
	 * (AF_ML >= 1e-4 and delta >= 3); 0 otherwise.  *1 = the early-onset variant (phenotype masks 4 / 5). */
	const double* assoc_p;      /* log10 p = kf_gammaq(0.5, LLR ln 10)/ln 10 */
	const double* assoc_llr;    /* LL(cases) + LL(controls) - LL(all) */
	const double* assoc_or;     /* odds ratio of the two allele-frequency estimates */
	const double* assoc_p1;
	const double* assoc_llr1;
	const double* assoc_or1;
	/* per called allele — [n_calls][P] */
	const int32_t* ac;          /* crispvar[].AC[i]  (bit-exact) */
	const double* qv;           /* crispvar[].QV[i] */
} crispgpu_results;

/* Device timings of the last completed batch, CUDA events on the engine's stream (ms).  The events are read when
 * crispgpu_get_timing is called: call it after crispgpu_wait / crispgpu_run_resident and before the next submit. */
typedef struct crispgpu_timing {
	float h2d_ms;
	float evaluate_ms;   /* evaluate_variant + contingency-table p-values (asymptotic) */
	float permute_ms;    /* permutation / exact p-value kernel */
	float filter_ms;     /* FAST_FILTER + task compaction */
	float em_ms;         /* likelihood + EM + genotypes + pool statistics */
	float legacy_ms;     /* --EM 0 caller */
	float d2h_ms;
	float total_ms;      /* first H2D byte to last D2H byte (kernels only for resident runs) */
	int32_t n_tasks_em;  /* (site,slot) tasks that reached the EM kernel */
	int32_t n_tasks_perm;
	int64_t launches;    /* kernels launched for the batch */
	int64_t h2d_bytes;   /* bytes moved host -> device for the batch, on-demand fetches included */
	int64_t d2h_bytes;
	int64_t fetched_bytes; /* the part of h2d_bytes pulled on demand from pinned host memory (k_fetch_reads; inside filter_ms) */
	float pileup_ms;       /* crispgpu_pileup_run: alignments -> counts -> candidate screen -> packed sites (after h2d_ms, before evaluate_ms) */
	float pile_tile_ms;    /* the part of pileup_ms spent in k_pile_tile (the kernel that streams the alignments) */
	int32_t n_positions;   /* crispgpu_pileup_run: reference positions piled up */
	int32_t n_alignments;  /* crispgpu_pileup_run: alignments handed over */
	int64_t pile_bytes;    /* algorithmic bytes of k_pile_tile: 16 per alignment + 1.5 per base + 32 per (position, pool) cell written */
} crispgpu_timing;

typedef struct crispgpu_ctx crispgpu_ctx;

/* Fill *cfg with the reference's defaults (readmultiplebams.c:21-47, newcrispcaller.c:20-32). */
void crispgpu_config_defaults(crispgpu_config* cfg);

/* Create an engine bound to CUDA device `device_id`.  `stream` may be a cudaStream_t the caller
 * owns (e.g. torch's current stream) or NULL for a private stream. */
int crispgpu_create(const crispgpu_config* cfg, int device_id, void* stream, crispgpu_ctx** out);
void crispgpu_destroy(crispgpu_ctx* ctx);

/* crispgpu_submit is asynchronous: it enqueues the host-to-device copies and the screening kernels and returns.
 * crispgpu_wait enqueues the caller kernels (they are sized by the number of surviving (site, slot) tasks, which has
 * arrived by then) and the device-to-host copies, and blocks until the results are in pinned host memory.
 * One batch may be in flight per context; keep several contexts busy (one front-end thread each; a context is not
 * thread-safe, different contexts are independent) so that the copies and kernels of consecutive batches overlap
 * (bench.py measures `e2e` that way, with six).  A second submit before wait returns
 * CRISPGPU_ERR_TICKET; so does a wait with a stale or unknown ticket. */
int crispgpu_submit(crispgpu_ctx* ctx, const crispgpu_batch* batch, int64_t* ticket);
int crispgpu_wait(crispgpu_ctx* ctx, int64_t ticket, crispgpu_results* out);
/* Candidate screen for n positions — what jointvariantcaller does between calculate_allelecounts and the call into the
 * engine (variantcalls.c:62-92): sort_allelecounts (parsebam/allelecounts.c:134-183, the
 * PICALL==3 ordering: reference allele first, then the exchange sort by read count) and the potential-variant score
 * (variantcalls.c:76-92: +2 per (allele, pool) with >= 3 reads of a non-reference allele, + the read count for diploid
 * pools with >= 1).  Inputs are the batch arrays `counts`, `refbase`, and `indcounts` or `ic_off`/`ic_sparse`; outputs
 * into caller memory: maxvec_al [n][8], maxvec_ct [n][8], potential [n].  A position is a candidate when
 * potential >= 2 (and its reference base is not N, which the front end knows).  Synchronous.
 * crispgpu_submit runs the same kernel for a batch whose maxvec_al/maxvec_ct are NULL. */
int crispgpu_screen_positions(crispgpu_ctx* ctx, int32_t n_sites, const uint8_t* refbase, const int32_t* counts,
                              const int32_t* indcounts, const uint32_t* ic_off, const crispgpu_count* ic_sparse,
                              uint8_t* maxvec_al, int32_t* maxvec_ct, int32_t* potential);

/* ---- pileup on the device -----------------------------------------------------------------
 * Replaces, for a window of reference positions, what the reference's front end does between reading the alignments and
 * the call into the engine: calculate_allelecounts + advance_read (parsebam/variant.c:225-342, 56-131), the mismatch
 * count of parse_cigar (parsebam/bamread.c:10-76), sort_allelecounts and the potential-variant test
 * (variantcalls.c:62-92), and the three read-list walks of host/crisp_shim.c.  The host inflates BGZF and splits the BAM
 * records (host/bam_decode.c); the alignments are the only input that crosses PCIe, in the BAM's own encoding (4-bit
 * bases, one quality byte per base): the candidate sites are found and packed in device memory and go straight
 * through the engine.
 *
 * This version piles up substitutions only: every alignment has the CIGAR [S|H] M [S|H] (one match block) and mates
 * must not overlap (the reference's overlapping-mate merge, variant.c:137-220, and its indel allele discovery stay on
 * the host).  An alignment outside that scope makes the call fail with CRISPGPU_ERR_UNSUPPORTED — it is never silently
 * mis-piled.  The front end filters on the BAM flag (BAM_FILTER_MASK, readmultiplebams.c:121) before handing reads over. */
typedef struct crispgpu_alnrec {  /* one alignment, 16 bytes */
	int32_t pos;        /* bam1_core_t.pos: 0-based reference position of the first base of the match block */
	uint32_t base_off;  /* index of the read's first base (soft clip included) in `qual`; a multiple of 8.  Its 4-bit
	                       bases start at seq4 + base_off/2 */
	uint16_t clip;      /* leading soft clip: bases before the match block */
	uint16_t mlen;      /* length of the match block */
	uint8_t mapq;       /* bam1_core_t.qual */
	uint8_t strand;     /* (flag & 16) != 0 */
	uint16_t reserved;
} crispgpu_alnrec;

typedef struct crispgpu_alignments {
	int32_t n_pools;
	int32_t reserved;
	const uint32_t* pool_off;      /* [n_pools+1] first alignment of each pool; a pool's alignments are in BAM (coordinate) order */
	const crispgpu_alnrec* rec;    /* [R], R = pool_off[n_pools] */
	const int32_t* mate_pos;       /* [R] bam1_core_t.mpos, or NULL for single-end data: used only to refuse overlapping mates */
	const int32_t* isize;          /* [R] bam1_core_t.isize, or NULL (with mate_pos) */
	const uint8_t* seq4;           /* [n_bases/2] bases as BAM stores them: two per byte, high nibble first, codes "=ACMGRSVTWYHKDBN" */
	const uint8_t* qual;           /* [n_bases] phred scores as BAM stores them (no +33) */
	uint64_t n_bases;              /* size of `qual`; every read occupies a multiple of 8 bases (padding is never read as data) */
} crispgpu_alignments;

typedef struct crispgpu_window {
	int32_t tid;               /* reference id (pass-through: results and RNG key) */
	int32_t start, end;        /* reference positions [start, end) to call */
	int32_t ref_start;         /* position of refseq[0]; the sequence must cover every alignment that overlaps the window */
	int32_t ref_len;
	int32_t want_streams;      /* 1: also return the packed read stream and alt-read lists of the candidate sites (tests) */
	int32_t min_mapq;          /* MIN_M (--mmq, default 20), readmultiplebams.c:26; the base quality threshold is config.min_q */
	int32_t reserved;
	const char* refseq;        /* reference bases (either case) */
} crispgpu_window;

/* The candidate sites of the window in coordinate order: what variantcalls.c:76-92 lets through, with the struct VARIANT
 * fields the printer needs.  Context-owned pinned memory, valid until the next call on the context. */
typedef struct crispgpu_pileup_sites {
	int32_t n_sites;
	int32_t n_positions;
	const int32_t* pos;        /* [n] */
	const uint8_t* refbase;    /* [n] */
	const uint8_t* maxvec_al;  /* [n][8] */
	const int32_t* maxvec_ct;  /* [n][8] */
	const int32_t* counts;     /* [n][24] */
	const int32_t* indcounts;  /* [n][P][24] */
	const int32_t* readdepths; /* [n][P]  variant->readdepths */
	const int32_t* mqcounts;   /* [n][4]  variant->MQcounts */
	const int32_t* filtered;   /* [n][2]  variant->filteredreads[0..1] */
	const uint32_t* read_off;  /* want_streams: [n*P+1] */
	const crispgpu_read* reads;
	const uint32_t* alt_off;   /* want_streams: [n*5+1] */
	const crispgpu_altread* alt_reads;
} crispgpu_pileup_sites;

/* Synchronous: pile up the window, screen and pack its candidate sites on the device, run the engine on them.
 * `results` is what crispgpu_wait would return for the same sites.  `aln` arrays in pinned memory (crispgpu_host_alloc)
 * make the copies DMA transfers. */
int crispgpu_pileup_run(crispgpu_ctx* ctx, const crispgpu_alignments* aln, const crispgpu_window* win, crispgpu_pileup_sites* sites,
                        crispgpu_results* results);
/* Device-resident variant used to time the pileup kernels alone: crispgpu_pileup_upload copies the alignments and the
 * reference window once, crispgpu_pileup_run_resident piles them up (fetch_results = 0: nothing is copied back) — the way to
 * time the kernels alone, and, with crispgpu_pileup_set_window, to call a decoded region window by window. */
int crispgpu_pileup_upload(crispgpu_ctx* ctx, const crispgpu_alignments* aln, const crispgpu_window* win);
/* Another window over the alignments that are resident: a region decoded once is called window by window without moving
 * its alignments again.  win->refseq NULL keeps the reference stretch already on the device. */
int crispgpu_pileup_set_window(crispgpu_ctx* ctx, const crispgpu_window* win);
int crispgpu_pileup_run_resident(crispgpu_ctx* ctx, int fetch_results, crispgpu_pileup_sites* sites, crispgpu_results* results);

/* Optional, between submit and wait: waits only for the screening counters of the batch and enqueues its caller
 * kernels and device-to-host copies without waiting for them.  A front end juggling two contexts calls it on the other
 * context before blocking in crispgpu_wait, so that this batch's EM runs while that batch's results travel back. */
int crispgpu_advance(crispgpu_ctx* ctx, int64_t ticket);

/* Device-resident path used to time the kernels alone: upload once, run many times. */
int crispgpu_upload(crispgpu_ctx* ctx, const crispgpu_batch* batch);
int crispgpu_run_resident(crispgpu_ctx* ctx, int fetch_results, crispgpu_results* out);

int crispgpu_get_timing(crispgpu_ctx* ctx, crispgpu_timing* out);
const char* crispgpu_last_error(crispgpu_ctx* ctx);

/* Narrow transport: bytes crispgpu_batch_pack_narrow needs for `compact` (a batch with bins/bin_off and ic_sparse/ic_off), or 0
 * when the batch cannot be packed (not compact, or a cell with more than 255 entries). */
size_t crispgpu_batch_narrow_bytes(const crispgpu_batch* compact, int32_t n_pools);
/* Packs every array of `compact` into `arena` (256-byte aligned, from crispgpu_host_alloc) with the two compact arrays in their narrow
 * form; *out views the arena (arena_base / arena_bytes set: one host-to-device copy).  Host code only.  Returns 0, or
 * CRISPGPU_ERR_ARG when the batch cannot be packed or the arena is too small. */
int crispgpu_batch_pack_narrow(const crispgpu_batch* compact, int32_t n_pools, void* arena, size_t arena_cap, crispgpu_batch* out);

/* Pinned host memory for batch arrays. */
void* crispgpu_host_alloc(size_t bytes);
void crispgpu_host_free(void* p);

int crispgpu_abi_version(void);
int crispgpu_device_count(void);

/* Measured FP64 FMA throughput of the device (TFLOP/s, 2 flop per DFMA) from a register-resident DFMA loop:
 * the roofline denominator of the EM kernel (MEASURED_PEAKS.json carries no FP64 figure). */
int crispgpu_measure_fp64_tflops(int device_id, double* tflops);
/* Measured 32-bit integer multiply-add throughput (10^12 IMAD lane-instructions per second): the roofline denominator of
 * the Monte-Carlo permutation kernels (counter-based RNG rounds and index arithmetic on the INT32 pipe). */
int crispgpu_measure_int32_tops(int device_id, double* tops);

#ifdef __cplusplus
}
#endif
#endif /* CRISPGPU_H */
