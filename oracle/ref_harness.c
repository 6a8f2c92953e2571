/*
 * ref_harness.c — TEST INFRASTRUCTURE ONLY (never linked into, loaded by or shipped with crisp_b200/).
 *
 * Drives the UNMODIFIED reference engine (vibansal/crisp, compiled where it lies under
 * /root/reference by oracle/Makefile into oracle/_ref/) on a `crispgpu_batch`, so that the CUDA
 * engine and the C restatement (oracle/crisp_oracle.c) can be checked against what CRISP itself
 * computes on identical inputs.  No reference source is copied: the reference's translation units
 *     crisp/newcrispcaller.c  parsebam/variant.c  parsebam/allelecounts.c  FET/chisquare.c  FET/contables.c
 * are compiled as they are, and this file only (a) defines the host globals they expect from
 * readmultiplebams.c:17-54, (b) rebuilds `struct VARIANT` (parsebam/variant.h:50-95) and the read queue
 * (`struct alignedread`, parsebam/bamread.h:17-43) from the batch, and (c) calls the reference entry points.
 *
 * Two drive modes:
 *   CRISP_REF_MODE_SLOTS   the per-allele loop of newCRISPcaller (crisp/newcrispcaller.c:205-246) is walked here,
 *                          calling the reference's evaluate_variant / compute_GLL_pooled / calculate_pool_statistics /
 *                          pooled_variantcaller / remove_ambiguous_bases, so that every (site, slot) state is visible
 *                          and HPlength can be taken from the batch;
 *   CRISP_REF_MODE_CALLER  newCRISPcaller itself is called (SNP-only sites; HPlength is recomputed by the reference from
 *                          a synthetic flanking sequence) — used to cross-check the slot walk and to obtain VCF text.
 * In both modes the VCF record is produced by the reference's own print_pooledvariant (crisp/newcrispprint.c:347).
 */
#define _GNU_SOURCE
#include <stdio.h>
#include <stdlib.h>
#include <string.h>
#include <unistd.h>
#include <fcntl.h>
#include <math.h>

#include "bamread.h"
#include "bamsreader.h"
#include "variant.h"
#include "allelecounts.h"
#include "contables.h"
#include "crispcaller.h"

#include "oracle_abi.h"
#define CRISP_SHIM_WITH_REFERENCE_TYPES
#include "../host/crisp_shim.h"

/* ---- globals the engine objects leave undefined (values: readmultiplebams.c:21-47, bamread.c:3) ---- */
int SOLID = 0;
int QVset = 0;
int USE_QV = 0;
int QVoffset = 33;
int MIN_M = 20, MAX_MM = 4;
int MINQ = 13;
int maxalleles = 8;
int READLENGTH = 100;
int MINFLANK = 5;
int OVERLAPPING_PE_READS = 1;
int INDEL_REALIGNMENT = 0;
int PIVOTSAMPLE = 0;
int PFLAG = 0;
char INT_CIGAROP[] = {'M', 'I', 'D', 'N', 'S', 'H', 'P', '=', 'X'};
int realign_read(struct BAMFILE_data* bd, int sample, struct alignedread* bcall, REFLIST* reflist) { return 0; }

/* ---- engine globals and functions defined inside the reference objects ---- */
extern int EMflag, FAST_FILTER, USE_BASE_QVS, MAXITER, MIN_READS, CHISQ_PERMUTATION;
extern double thresh1, MINQpv, thresh3, SER, RELAX, Qmin, AGILENT_REF_BIAS, THETA;
int evaluate_variant(READQUEUE* bq, struct VARIANT* variant, int allele1, int allele2, int allele3, double paf[]);
int compute_GLL_pooled(READQUEUE* bq, struct VARIANT* variant, int allele1, int allele2, int allele3, double p[]);
int calculate_pool_statistics(READQUEUE* bq, struct BAMFILE_data* bd, struct VARIANT* variant, int allele1, int allele2);
int remove_ambiguous_bases(READQUEUE* bq, struct VARIANT* variant, int HPlength, int delta);
int print_pooledvariant(struct VARIANT* variant, FILE* vfile, REFLIST* reflist, int current, ALLELE* maxvec, int EMflag);
void calculate_pooled_likelihoods_snp(READQUEUE* bq, struct VARIANT* variant, int a1, int a2, int a3, double E01[], double E10[]);
double ncr(int n, int r);
double fet(int c1, int r1, int c2, int r2);
double minreads_pvalue(int R, int A, double e);
double binomial_pvalue(int R, int A, double e);
double kf_lgamma(double z);
double kf_gammaq(double s, double z);
void chernoff_bound(int* counts, double* stats, int refbase, int altbase, double SER, double* P0, double* P1, int maxalleles);
int pvalue_contable_iter_stranded(int** ctable, double** ctable_weighted, int* strata, int size, int maxiter, int** newtable, double* permp, double* chisqp, double chi2stats[]);
int pvalue_lowcoverage(int** ctable, int size, int iter, struct acounts* atable, double* permp);

enum { CRISP_REF_MODE_SLOTS = 0, CRISP_REF_MODE_CALLER = 1 };

#define QBUFLEN 1024
#define FLANK 64

typedef struct ref_ctx {
	crispgpu_config cfg;
	int P, maxploidy;
	int* ploidy;
	int* phenotypes;
	struct OPTIONS options;
	struct VARIANT variant;
	struct BAMFILE_data* bd;
	READQUEUE rq;
	struct alignedread* reads;
	int reads_cap;
	int* cigars;
	char* basebuf[4];
	char* qualbuf[256];
	REFLIST reflist;
	char* refname;
	unsigned char* refseq;
	int reflen;
	char err[256];
	/* optional printer state of the batch being replayed (crisp_ref_set_printer_state): what the front end recorded */
	const int32_t* pr_depths;   /* [n][P] variant->readdepths */
	const int32_t* pr_mq;       /* [n][4] variant->MQcounts */
} ref_ctx;

static const char BASES[4] = {'A', 'C', 'G', 'T'};

static void apply_config(const ref_ctx* c)
{
	const crispgpu_config* g = &c->cfg;
	EMflag = g->em_flag;
	MAXITER = g->max_iter;
	CHISQ_PERMUTATION = g->chisq_permutation;
	FAST_FILTER = g->fast_filter;
	USE_BASE_QVS = g->use_base_qvs;
	MIN_READS = g->min_reads;
	QVoffset = g->qv_offset;
	MINQ = g->min_q;
	PIVOTSAMPLE = g->pivot_sample;
	thresh1 = g->thresh1;
	MINQpv = g->min_qpv;
	thresh3 = g->thresh3;
	SER = g->ser;
	RELAX = g->relax;
	Qmin = g->qmin;
	AGILENT_REF_BIAS = g->agilent_ref_bias;
	THETA = g->theta;
	USE_QV = g->use_qv;
	PFLAG = 0;
}

void* crisp_ref_create(const crispgpu_config* cfg)
{
	if (!cfg || cfg->n_pools <= 0 || !cfg->ploidy) return NULL;
	ref_ctx* c = (ref_ctx*)calloc(1, sizeof(ref_ctx));
	c->cfg = *cfg;
	c->P = cfg->n_pools;
	c->ploidy = (int*)malloc(sizeof(int) * c->P);
	for (int i = 0; i < c->P; i++) {
		c->ploidy[i] = cfg->ploidy[i];
		if (cfg->ploidy[i] > c->maxploidy) c->maxploidy = cfg->ploidy[i];
	}
	c->cfg.ploidy = c->ploidy;
	c->phenotypes = (int*)calloc(c->P, sizeof(int));
	if (cfg->phenotypes) for (int i = 0; i < c->P; i++) c->phenotypes[i] = cfg->phenotypes[i];
	c->cfg.phenotypes = c->phenotypes;
	memset(&c->options, 0, sizeof(c->options));
	c->options.bamfiles = c->P;
	c->options.samples = c->P;
	c->options.association = cfg->association;
	c->options.phenotypes = c->phenotypes;
	c->options.ploidy = c->ploidy;
	/* as multisampleVC does (readmultiplebams.c:84-87): ploidy first, then init_variant */
	memset(&c->variant, 0, sizeof(c->variant));
	c->variant.ploidy = (int*)calloc(c->P, sizeof(int));
	for (int i = 0; i < c->P; i++) c->variant.ploidy[i] = c->ploidy[i];
	c->variant.varpoolsize = cfg->varpoolsize;
	apply_config(c);
	init_variant(&c->variant, c->P, c->P);
	c->variant.options = &c->options;
	c->bd = (struct BAMFILE_data*)calloc(c->P, sizeof(struct BAMFILE_data));
	for (int b = 0; b < 4; b++) {
		c->basebuf[b] = (char*)malloc(QBUFLEN);
		memset(c->basebuf[b], BASES[b], QBUFLEN);
	}
	/* synthetic flanking sequence for print_flanking_sequence / calculate_indel_ambiguity */
	c->reflen = 2 * FLANK + 1;
	c->refseq = (unsigned char*)malloc(c->reflen + 1);
	c->refname = (char*)malloc(64);
	c->reflist.ns = 1;
	c->reflist.names = &c->refname;
	c->reflist.lengths = &c->reflen;
	c->reflist.sequences = &c->refseq;
	return c;
}

void crisp_ref_destroy(void* h)
{
	/* the reference never frees struct VARIANT; the harness leaks it the same way (test process only) */
	ref_ctx* c = (ref_ctx*)h;
	if (!c) return;
	free(c->reads);
	free(c->cigars);
	free(c->bd);
	free(c);
}

static char* qual_buffer(ref_ctx* c, int q)
{
	q &= 0xff;
	if (!c->qualbuf[q]) {
		c->qualbuf[q] = (char*)malloc(QBUFLEN);
		memset(c->qualbuf[q], q, QBUFLEN);
	}
	return c->qualbuf[q];
}

/* deterministic flanking sequence: position-hashed bases, site base = refbase, base after the site differs from it
 * (so that the CALLER mode never finds a homopolymer); previousbase etc. come from here in both modes. */
static void make_flank(ref_ctx* c, int tid, int pos, int refbase)
{
	for (int k = 0; k < c->reflen; k++) {
		unsigned x = (unsigned)(pos + k - FLANK) * 2654435761u + (unsigned)tid * 40503u;
		x ^= x >> 15;
		c->refseq[k] = BASES[(x >> 7) & 3];
	}
	c->refseq[c->reflen] = 0;
	c->refseq[FLANK] = BASES[refbase & 3];
	snprintf(c->refname, 64, "chr%d", tid);
}

static void fill_site(ref_ctx* c, const crispgpu_batch* b, int s)
{
	struct VARIANT* v = &c->variant;
	const int P = c->P;
	v->position = FLANK; /* position inside the synthetic flank; the VCF column is patched by the caller */
	v->refbase = b->refbase[s];
	v->refb = BASES[v->refbase & 3];
	v->IUPAC = 0;
	snprintf(v->chrom, 64, "chr%d", b->tid[s]);
	for (int a = 0; a < 4; a++) {
		v->itb[a][0] = BASES[a];
		v->itb[a][1] = 0;
	}
	for (int a = 4; a < 8; a++) {
		int len = b->indel_len ? b->indel_len[s * 8 + a] : 0;
		int L = len < 0 ? -len : len;
		if (L > 1000) L = 1000;
		v->itb[a][0] = len < 0 ? '-' : '+';
		for (int k = 0; k < L; k++) v->itb[a][1 + k] = BASES[(a + k) & 3];
		v->itb[a][1 + L] = 0;
		if (len == 0) {
			v->itb[a][0] = 'N';
			v->itb[a][1] = 0;
		}
	}
	for (int k = 0; k < 10; k++) v->HPlength[k] = 0;
	if (b->hplen) for (int a = 0; a < 8; a++) v->HPlength[a] = b->hplen[s * 8 + a];
	v->varalleles = 0;
	v->strandpass = 0;
	for (int q = 0; q < 32; q++) v->counts[q] = q < 24 ? b->counts[s * 24 + q] : 0;
	for (int q = 0; q < 16; q++) v->tstats[q] = b->tstats ? b->tstats[s * 16 + q] : 0.0;
	int total = 0;
	for (int i = 0; i < P; i++) {
		const int32_t* ic = b->indcounts + ((size_t)s * P + i) * 24;
		int depth = 0;
		for (int q = 0; q < 32; q++) v->indcounts[i][q] = q < 24 ? ic[q] : 0;
		for (int q = 0; q < 24; q++) depth += ic[q];
		v->readdepths[i] = depth;
		total += depth;
		for (int q = 0; q < 48; q++) v->Qhighcounts[i][q] = 0.0;
		for (int q = 0; q < 16; q++) {
			v->Qhighcounts[i][q] = b->qhigh ? b->qhigh[((size_t)s * P + i) * 16 + q] : (double)ic[q];
			v->stats[i][q] = b->stats ? b->stats[((size_t)s * P + i) * 16 + q] : 0.0;
		}
	}
	for (int k = 0; k < 8; k++) v->filteredreads[k] = 0;
	v->MQcounts[0] = v->MQcounts[1] = v->MQcounts[2] = 0;
	v->MQcounts[3] = total;
	make_flank(c, b->tid[s], b->pos[s], v->refbase);
}

/* Build the read queue of site s.  Returns 0, or -1 when the alt-read side list does not match the stream. */
static int build_reads(ref_ctx* c, const crispgpu_batch* b, int s)
{
	struct VARIANT* v = &c->variant;
	const int P = c->P;
	const uint32_t r0 = b->read_off[(size_t)s * P], r1 = b->read_off[(size_t)(s + 1) * P];
	uint32_t nalt = 0;
	if (b->alt_off) nalt = b->alt_off[s * 5 + 5] - b->alt_off[s * 5];
	int need = (int)(r1 - r0) + (int)nalt + 1;
	if (need > c->reads_cap) {
		c->reads_cap = need * 2;
		c->reads = (struct alignedread*)realloc(c->reads, sizeof(struct alignedread) * c->reads_cap);
		c->cigars = (int*)realloc(c->cigars, sizeof(int) * c->reads_cap);
	}
	int n = 0;
	struct alignedread* prev = NULL;
	c->rq.first = c->rq.last = NULL;
	c->rq.reads = 0;
	for (int i = 0; i < P; i++) c->bd[i].first = c->bd[i].last = NULL;
	/* per (slot) cursors into the alt list, per pool */
	for (int i = 0; i < P; i++) {
		const uint32_t a0 = b->read_off[(size_t)s * P + i], a1 = b->read_off[(size_t)s * P + i + 1];
		struct alignedread* pprev = NULL;
		uint32_t cursor[5], cend[5];
		int taken[5] = {0, 0, 0, 0, 0};
		for (int k = 0; k < 5; k++) {
			cursor[k] = cend[k] = 0;
			if (!b->alt_off) continue;
			uint32_t lo = b->alt_off[s * 5 + k], hi = b->alt_off[s * 5 + k + 1];
			while (lo < hi && b->alt_reads[lo].pool < i) lo++;
			uint32_t e = lo;
			while (e < hi && b->alt_reads[e].pool == i) e++;
			cursor[k] = lo;
			cend[k] = e;
		}
		/* ambiguous reference reads: the m-th reference read of class cls gets remaining match length
		 * rem = min HPlength[slot allele] over slots whose decrement row covers m (nested by construction) */
		int refseen[4] = {0, 0, 0, 0};
		for (uint32_t r = a0; r < a1; r++) {
			crispgpu_read w = b->reads[r];
			struct alignedread* rd = &c->reads[n];
			memset(rd, 0, sizeof(*rd));
			int q = w & 0xff, base = (w >> 8) & 3, rev = (w >> 10) & 1, bidir = (w >> 11) & 1;
			rd->sampleid = i;
			rd->strand = rev;
			rd->bidir = bidir;
			rd->position = v->position - 10;
			rd->lastpos = v->position + 50;
			rd->filter = '0';
			rd->type = 0;
			rd->sequence = c->basebuf[base];
			rd->quality = qual_buffer(c, q);
			rd->l1 = 50;
			rd->delta = 0;
			rd->alignedbases = 100;
			rd->allele = base + (rev ? 8 : 0);
			rd->mquality = 60;
			int rem = 1000000;
			if (base == v->refbase && b->amb_idx) {
				int cls = bidir ? 2 + rev : rev;
				int m = refseen[cls]++;
				for (int k = 0; k < 5; k++) {
					int row = b->amb_idx[s * 5 + k];
					if (row < 0) continue;
					int dec = b->amb_dec[((size_t)row * P + i) * 4 + cls];
					int hp = v->HPlength[b->maxvec_al[s * 8 + k + 1]];
					if (m < dec && hp < rem) rem = hp;
				}
			}
			c->cigars[n] = (rem << 4) | BAM_CMATCH;
			rd->fcigarlist = &c->cigars[n];
			rd->fcigs = 1;
			rd->cigoffset = 0;
			/* alternate-allele reads that calculate_uniquereads (allelecounts.c:53-57) would collect */
			if (q >= MINQ + QVoffset) {
				for (int k = 0; k < 5; k++) {
					int al = b->maxvec_al[s * 8 + k + 1];
					if (al >= 4 || al != base) continue;
					if (cursor[k] < cend[k]) {
						const crispgpu_altread* e = &b->alt_reads[cursor[k]++];
						if (e->strand != rev) {
							snprintf(c->err, sizeof c->err, "site %d slot %d pool %d: alt-read strand mismatch", s, k, i);
							return -1;
						}
						rd->l1 = e->offset;
						rd->alignedbases = e->aligned;
						taken[k]++;
					}
				}
			}
			if (prev) prev->next = rd; else c->rq.first = rd;
			if (pprev) pprev->nextread = rd; else c->bd[i].first = rd;
			prev = rd;
			pprev = rd;
			n++;
		}
		/* indel alleles: one synthetic indel-carrying read per alt-list entry */
		for (int k = 0; k < 5; k++) {
			int al = b->maxvec_al[s * 8 + k + 1];
			if (al < 4) {
				if (cursor[k] != cend[k]) {
					snprintf(c->err, sizeof c->err, "site %d slot %d pool %d: %u alt-list entries without a matching read", s, k, i, cend[k] - cursor[k]);
					return -1;
				}
				continue;
			}
			while (cursor[k] < cend[k]) {
				const crispgpu_altread* e = &b->alt_reads[cursor[k]++];
				struct alignedread* rd = &c->reads[n];
				memset(rd, 0, sizeof(*rd));
				rd->sampleid = i;
				rd->strand = e->strand;
				rd->position = v->position - 10;
				rd->lastpos = v->position + 50;
				rd->filter = '0';
				rd->type = b->indel_len ? b->indel_len[s * 8 + al] : 1;
				rd->sequence = c->basebuf[0];
				rd->quality = qual_buffer(c, 30 + 33);
				rd->l1 = e->offset;
				rd->alignedbases = e->aligned;
				rd->allele = -1;
				rd->mquality = 60;
				c->cigars[n] = (1000000 << 4) | BAM_CMATCH;
				rd->fcigarlist = &c->cigars[n];
				rd->fcigs = 1;
				if (prev) prev->next = rd; else c->rq.first = rd;
				if (pprev) pprev->nextread = rd; else c->bd[i].first = rd;
				prev = rd;
				pprev = rd;
				n++;
			}
		}
		c->bd[i].last = pprev;
	}
	c->rq.last = prev;
	c->rq.reads = n;
	return 0;
}

static int stdout_saved = -1;
static FILE* stdout_capture = NULL;
/* the engine logs to stdout (and the association statistic exists nowhere else: LRstatistic.c:62-64): capture it */
static void mute_stdout(void)
{
	fflush(stdout);
	stdout_saved = dup(1);
	stdout_capture = tmpfile();
	if (stdout_capture) dup2(fileno(stdout_capture), 1);
	else {
		int nul = open("/dev/null", O_WRONLY);
		dup2(nul, 1);
		close(nul);
	}
}
static void unmute_stdout(void)
{
	fflush(stdout);
	if (stdout_saved >= 0) {
		dup2(stdout_saved, 1);
		close(stdout_saved);
		stdout_saved = -1;
	}
}

static void dump_genll(ref_ctx* c, crisp_host_results* out, int s, int k)
{
	if (!out->genll) return;
	const int P = c->P, st = out->genll_stride;
	struct VARIANT* v = &c->variant;
	double** src[4] = {v->GENLL, v->GENLLf, v->GENLLr, v->GENLLb};
	for (int t = 0; t < 4; t++)
		for (int i = 0; i < P; i++) {
			double* dst = out->genll + ((((size_t)s * 5 + k) * 4 + t) * P + i) * st;
			for (int j = 0; j <= c->ploidy[i] && j < st; j++) dst[j] = src[t][i][j];
		}
}

/* Parse the captured `ASSOC` lines (one per slot that met crispEM.c:383, in call order) into the result arrays.
 * Printed precision only: %0.2f / %0.3f. */
static void collect_assoc(crisp_host_results* out, const int* slots, int nslots)
{
	if (!stdout_capture) return;
	char* line = NULL;
	size_t cap = 0;
	int k = 0;
	rewind(stdout_capture);
	while (getline(&line, &cap, stdout_capture) > 0) {
		char* a = strstr(line, "ASSOC ");
		if (!a) continue;
		char* bar = strchr(a, '|');
		if (!bar || k >= nslots) continue;
		double pc, pctl, pca, pca1, l0, l1, l2, llr, orr, pv, m0, m1, m2, llr1, or1, pv1;
		int got = sscanf(bar + 1, " %lf %lf %lf %lf LL %lf %lf %lf LLR %lf:%lf:%lf LL %lf %lf %lf LLR %lf:%lf:%lf", &pc, &pctl, &pca, &pca1, &l0, &l1, &l2, &llr, &orr, &pv, &m0, &m1, &m2, &llr1, &or1, &pv1);
		if (got == 16 && out->assoc_p) {
			int o = slots[k];
			out->assoc_p[o] = pv; out->assoc_llr[o] = llr; out->assoc_or[o] = orr;
			out->assoc_p1[o] = pv1; out->assoc_llr1[o] = llr1; out->assoc_or1[o] = or1;
		}
		k++;
	}
	free(line);
	fclose(stdout_capture);
	stdout_capture = NULL;
}

const char* crisp_ref_last_error(void* h) { return ((ref_ctx*)h)->err; }

/*
 * Run the reference engine over the batch.  rng_seed: srand48(seed) once at the start (the reference seeds with
 * time(NULL), readmultiplebams.c:252).  vcf: optional buffer receiving the VCF records of called sites,
 * each line prefixed by "<site index>\t" so it can be matched to the batch.
 */
int crisp_ref_run(void* h, const crispgpu_batch* b, crisp_host_results* out, int mode, long rng_seed, char* vcf, size_t vcf_cap, size_t* vcf_len)
{
	ref_ctx* c = (ref_ctx*)h;
	if (!c || !b || !out) return CRISPGPU_ERR_ARG;
	struct VARIANT* v = &c->variant;
	const int P = c->P, n = b->n_sites;
	apply_config(c);
	srand48(rng_seed);
	char* mem = NULL;
	size_t memlen = 0;
	FILE* vf = open_memstream(&mem, &memlen);
	char* tmem = NULL;
	size_t tlen = 0;
	FILE* tf = open_memstream(&tmem, &tlen);
	ALLELE maxvec[8];
	int rc = 0;
	out->n_sites = n;
	out->n_calls = 0;
	int* assoc_slots = (int*)malloc(sizeof(int) * (size_t)(n > 0 ? n : 1) * 5);
	int n_assoc = 0;
	mute_stdout();
	for (int s = 0; s < n && rc == 0; s++) {
		fill_site(c, b, s);
		for (int k = 0; k < 8; k++) {
			maxvec[k].al = b->maxvec_al[s * 8 + k];
			maxvec[k].ct = b->maxvec_ct[s * 8 + k];
			maxvec[k].vars = 0;
		}
		for (int k = 0; k < 5; k++) {
			int o = s * 5 + k;
			out->status[o] = CRISPGPU_SLOT_UNTESTED;
			out->allele[o] = maxvec[k + 1].al;
			out->call_rank[o] = -1;
			out->call_row[o] = -1;
			out->paf[o] = out->ctpval[o] = out->chi2pval[o] = out->af_ml[o] = out->delta[o] = out->deltaf[o] = out->hwe[o] = 0;
			out->qvpf[o] = out->qvpr[o] = 0;
			if (out->assoc_p) out->assoc_p[o] = out->assoc_llr[o] = out->assoc_or[o] = out->assoc_p1[o] = out->assoc_llr1[o] = out->assoc_or1[o] = 0;
			out->varpools[o] = out->allelecount[o] = out->variantpools[o] = 0;
			if (out->em_iters) out->em_iters[o] = 0;
			if (out->perm_k) out->perm_k[o] = 0;
			if (out->perm_hits) out->perm_hits[o] = 0;
		}
		int reads_built = 0;
		if (mode == CRISP_REF_MODE_CALLER) {
			if (build_reads(c, b, s) != 0) { rc = CRISPGPU_ERR_ARG; break; }
			rewind(tf);
			int called = newCRISPcaller(&c->reflist, 0, v->position, &c->rq, c->bd, v, maxvec, tf);
			fflush(tf);
			if (called) {
				fprintf(vf, "%d\t", s);
				fwrite(tmem, 1, tlen, vf);
			}
			out->varalleles[s] = v->varalleles;
			out->strandpass[s] = v->strandpass;
			for (int r = 0; r < v->varalleles; r++) {
				int k;
				for (k = 0; k < 5; k++)
					if (maxvec[k + 1].ct >= 3 && maxvec[k + 1].al == v->alleles[r]) break;
				if (k == 5) continue;
				int o = s * 5 + k;
				out->status[o] = CRISPGPU_SLOT_CALLED;
				out->call_rank[o] = r;
				out->ctpval[o] = v->ctpval[r];
				out->chi2pval[o] = v->chi2pval[r];
				out->af_ml[o] = v->crispvar[r].AF_ML;
				out->delta[o] = v->crispvar[r].delta;
				out->deltaf[o] = v->crispvar[r].deltaf;
				out->hwe[o] = v->crispvar[r].deltar;
				out->qvpf[o] = v->qvpvaluef[r];
				out->qvpr[o] = v->qvpvaluer[r];
				out->varpools[o] = v->nvariants[r];
				out->allelecount[o] = v->crispvar[r].allelecount;
				out->variantpools[o] = v->crispvar[r].variantpools;
				if (EMflag >= 1 && out->n_calls < out->max_calls) {
					int row = out->n_calls++;
					out->call_row[o] = row;
					for (int i = 0; i < P; i++) {
						out->ac[(size_t)row * P + i] = v->crispvar[r].AC[i];
						out->qv[(size_t)row * P + i] = v->crispvar[r].QV[i];
					}
				}
			}
			continue;
		}
		/* ---- CRISP_REF_MODE_SLOTS: the loop of newcrispcaller.c:203-246, with reference callees ---- */
		v->strandpass = 0;
		v->varalleles = 0;
		double paf[2] = {0.0, 0.0};
		for (int al = 1; al <= 5 && rc == 0; al++) {
			if (maxvec[al].ct < 3) continue;
			int o = s * 5 + al - 1;
			int allele1 = maxvec[0].al, allele2 = maxvec[al].al, allele3 = maxvec[al == 1 ? al + 1 : 1].al;
			v->ctpval[v->varalleles] = 0;
			int indelflag = 0, is_variant = 0;
			if ((allele2 >= 4 && v->HPlength[allele2] > 0) || (allele3 >= 4 && v->HPlength[allele3] >= 100)) {
				indelflag = v->HPlength[allele2];
				if (!reads_built) {
					if (build_reads(c, b, s) != 0) { rc = CRISPGPU_ERR_ARG; break; }
					reads_built = 1;
				}
				remove_ambiguous_bases(&c->rq, v, indelflag, 1);
			}
			int potentialvar = evaluate_variant(&c->rq, v, allele1, allele2, allele3, paf);
			out->ctpval[o] = v->ctpval[v->varalleles];
			out->chi2pval[o] = v->chi2pval[v->varalleles];
			if (potentialvar == 1) {
				out->paf[o] = paf[0];
				if (!reads_built) {
					if (build_reads(c, b, s) != 0) { rc = CRISPGPU_ERR_ARG; break; }
					reads_built = 1;
				}
				out->status[o] = CRISPGPU_SLOT_NOCALL;
				int slotrank = v->varalleles;
				if (EMflag >= 1) {
					is_variant = compute_GLL_pooled(&c->rq, v, allele1, allele2, allele3, paf);
					dump_genll(c, out, s, al - 1);
					out->af_ml[o] = v->crispvar[slotrank].AF_ML;
					out->delta[o] = v->crispvar[slotrank].delta;
					out->deltaf[o] = v->crispvar[slotrank].deltaf;
					out->hwe[o] = v->crispvar[slotrank].deltar;
					if (c->options.association == 1 && paf[0] >= 0.0001 && v->crispvar[slotrank].delta >= 3) assoc_slots[n_assoc++] = o;  /* crispEM.c:383 */
					if (is_variant == 1) {
						int variantpools = calculate_pool_statistics(&c->rq, c->bd, v, allele1, allele2);
						v->varpools[v->varalleles] = variantpools;
						v->alleles[v->varalleles] = allele2;
						v->nvariants[v->varalleles] = variantpools;
						v->varalleles++;
					}
				} else {
					is_variant = pooled_variantcaller(&c->rq, c->bd, v, allele1, allele2);
					out->qvpf[o] = v->qvpvaluef[slotrank];
					out->qvpr[o] = v->qvpvaluer[slotrank];
				}
				if (is_variant == 1) {
					out->status[o] = CRISPGPU_SLOT_CALLED;
					out->call_rank[o] = slotrank;
					out->varpools[o] = v->nvariants[slotrank];
					out->allelecount[o] = v->crispvar[slotrank].allelecount;
					out->variantpools[o] = v->crispvar[slotrank].variantpools;
					if (EMflag >= 1 && out->n_calls < out->max_calls) {
						int row = out->n_calls++;
						out->call_row[o] = row;
						for (int i = 0; i < P; i++) {
							out->ac[(size_t)row * P + i] = v->crispvar[slotrank].AC[i];
							out->qv[(size_t)row * P + i] = v->crispvar[slotrank].QV[i];
						}
					}
				}
			} else {
				/* distinguish the two reject points of evaluate_variant: the AF filter returns before any
				 * contingency-table work (newcrispcaller.c:72), so ctpval is still the 0 written at :209 */
				out->status[o] = CRISPGPU_SLOT_REJECT_CT;
				{
					/* recompute the AF filter decision exactly as :48-72 does */
					double p0 = 0;
					for (int i = 0; i < P; i++) {
						double C1 = 0, C2 = 0, C3 = 0;
						for (int t = 0; t < 3; t++) {
							C1 += v->indcounts[i][allele1 + t * 8];
							C2 += v->indcounts[i][allele2 + t * 8];
							C3 += v->indcounts[i][allele3 + t * 8];
						}
						double p1 = C2 / (C1 + C2 + C3 + 0.001);
						p1 *= v->ploidy[i];
						if (p1 >= 0.2 && C2 >= 4) p0 += p1;
						else if (p1 >= 0.4 && C2 >= 3) p0 += p1;
						else if (v->ploidy[i] == 2 && p1 >= 0.25 && C2 >= 1) p0 += p1;
					}
					out->paf[o] = p0;
					if (p0 < 0.25 || (v->ploidy[0] == 2 && p0 <= 0.4)) {
						out->status[o] = CRISPGPU_SLOT_REJECT_AF;
						out->chi2pval[o] = 0; /* not written on this path: variant->chi2pval[] still holds a stale value */
					}
				}
			}
			if (indelflag > 0) remove_ambiguous_bases(&c->rq, v, indelflag, -1);
		}
		out->varalleles[s] = v->varalleles;
		out->strandpass[s] = v->strandpass;
		if (v->varalleles >= 1 && rc == 0) {
			int indelflag = 0;
			for (int i = 0; i < v->varalleles; i++)
				if (v->alleles[i] >= 4 && v->HPlength[v->alleles[i]] > indelflag) indelflag = v->HPlength[v->alleles[i]];
			if (indelflag > 0) remove_ambiguous_bases(&c->rq, v, indelflag, 1);
			fprintf(vf, "%d\t", s);
			print_pooledvariant(v, vf, &c->reflist, 0, maxvec, EMflag);
		}
	}
	unmute_stdout();
	collect_assoc(out, assoc_slots, mode == CRISP_REF_MODE_SLOTS ? n_assoc : 0);
	free(assoc_slots);
	fclose(vf);
	fclose(tf);
	free(tmem);
	if (vcf && vcf_cap > 0) {
		size_t m = memlen < vcf_cap - 1 ? memlen : vcf_cap - 1;
		memcpy(vcf, mem, m);
		vcf[m] = 0;
	}
	if (vcf_len) *vcf_len = memlen;
	free(mem);
	return rc;
}

/*
 * Replay: fill struct VARIANT from a results block (as the host shim of INTEGRATION.md does) and let the
 * reference's print_pooledvariant write the VCF record.  Used to prove "GPU results -> unchanged printer"
 * gives the same text as the reference end to end.
 */
/* Read depths and mapping-quality classes the next crisp_ref_replay_vcf prints (NULL: rebuilt from the counts, as before).
 * The front end owns these counters (parsebam/variant.c:280-283); the batch does not carry them. */
void crisp_ref_set_printer_state(void* h, const int32_t* readdepths, const int32_t* mqcounts)
{
	ref_ctx* c = (ref_ctx*)h;
	c->pr_depths = readdepths;
	c->pr_mq = mqcounts;
}

int crisp_ref_replay_vcf(void* h, const crispgpu_batch* b, const crisp_host_results* res, char* vcf, size_t vcf_cap, size_t* vcf_len)
{
	ref_ctx* c = (ref_ctx*)h;
	struct VARIANT* v = &c->variant;
	const int P = c->P;
	apply_config(c);
	char* mem = NULL;
	size_t memlen = 0;
	FILE* vf = open_memstream(&mem, &memlen);
	ALLELE maxvec[8];
	crispgpu_results view;
	memset(&view, 0, sizeof view);
	view.n_sites = res->n_sites; view.n_calls = res->n_calls;
	view.varalleles = res->varalleles; view.strandpass = res->strandpass; view.status = res->status; view.allele = res->allele;
	view.call_rank = res->call_rank; view.call_row = res->call_row; view.ctpval = res->ctpval; view.chi2pval = res->chi2pval;
	view.af_ml = res->af_ml; view.delta = res->delta; view.deltaf = res->deltaf; view.hwe = res->hwe; view.qvpf = res->qvpf; view.qvpr = res->qvpr;
	view.varpools = res->varpools; view.allelecount = res->allelecount; view.variantpools = res->variantpools; view.ac = res->ac; view.qv = res->qv;
	for (int s = 0; s < b->n_sites; s++) {
		if (res->varalleles[s] < 1) continue;
		fill_site(c, b, s);
		if (c->pr_depths) for (int i = 0; i < c->P; i++) v->readdepths[i] = c->pr_depths[(size_t)s * c->P + i];
		if (c->pr_mq) for (int k = 0; k < 4; k++) v->MQcounts[k] = c->pr_mq[s * 4 + k];
		for (int k = 0; k < 8; k++) {
			maxvec[k].al = b->maxvec_al[s * 8 + k];
			maxvec[k].ct = b->maxvec_ct[s * 8 + k];
		}
		crisp_shim_fill_variant(&view, s, v);  /* the production copy-back (host/crisp_shim.c) */
		int hpmax = 0, hpslot = -1;
		for (int k = 0; k < 5; k++) {
			int o = s * 5 + k;
			if (res->status[o] != CRISPGPU_SLOT_CALLED || res->call_rank[o] < 0) continue;
			if (res->allele[o] >= 4 && v->HPlength[res->allele[o]] > hpmax) {
				hpmax = v->HPlength[res->allele[o]];
				hpslot = k;
			}
		}
		/* newcrispcaller.c:253-263: counts are reduced once more with the largest called-indel HPlength */
		if (hpmax > 0 && hpslot >= 0 && b->amb_idx && b->amb_idx[s * 5 + hpslot] >= 0) {
			int row = b->amb_idx[s * 5 + hpslot];
			for (int i = 0; i < P; i++) {
				const int32_t* d = b->amb_dec + ((size_t)row * P + i) * 4;
				v->indcounts[i][v->refbase] -= d[0];
				v->indcounts[i][v->refbase + 8] -= d[1];
				v->indcounts[i][v->refbase + 16] -= d[2] + d[3];
			}
		}
		fprintf(vf, "%d\t", s);
		print_pooledvariant(v, vf, &c->reflist, 0, maxvec, EMflag);
	}
	fclose(vf);
	if (vcf && vcf_cap > 0) {
		size_t m = memlen < vcf_cap - 1 ? memlen : vcf_cap - 1;
		memcpy(vcf, mem, m);
		vcf[m] = 0;
	}
	if (vcf_len) *vcf_len = memlen;
	free(mem);
	return 0;
}

/* ---- scalar known-answer entry points (SURVEY.md §8c level 2) ---- */
double crisp_ref_ncr(int n, int r) { return ncr(n, r); }
double crisp_ref_fet(int c1, int r1, int c2, int r2) { return fet(c1, r1, c2, r2); }
double crisp_ref_minreads_pvalue(int R, int A, double e) { return minreads_pvalue(R, A, e); }
double crisp_ref_binomial_pvalue(int R, int A, double e) { return binomial_pvalue(R, A, e); }
double crisp_ref_kf_lgamma(double z) { return kf_lgamma(z); }
double crisp_ref_kf_gammaq(double s, double z) { return kf_gammaq(s, z); }
void crisp_ref_chernoff(int* counts, double* stats, int refbase, int altbase, double ser, double* p0, double* p1)
{
	chernoff_bound(counts, stats, refbase, altbase, ser, p0, p1, 8);
}


/*
 * Round trip of the production packer (host/crisp_shim.c): rebuild struct VARIANT and the read lists of every site of
 * batch `b`, let crisp_shim_append_site pack them again, and compare the two batches.  Returns the number of arrays
 * that differ (0 = identical) and describes the first difference in msg.
 */
int crisp_ref_shim_roundtrip(void* h, const crispgpu_batch* b, char* msg, size_t msgcap)
{
	ref_ctx* c = (ref_ctx*)h;
	struct VARIANT* v = &c->variant;
	const int P = c->P, n = b->n_sites;
	apply_config(c);
	crisp_shim_batch* sb = crisp_shim_batch_new(P, b->qhigh != NULL, b->stats != NULL);
	/* a batch that also carries bins asks for the compact packer: its bins must come back, the reads are not emitted */
	const int compact = b->bins != NULL && b->bin_off != NULL;
	crisp_shim_batch_set_compact(sb, compact);
	ALLELE maxvec[8];
	int bad = 0;
	if (msg && msgcap) msg[0] = 0;
	for (int s = 0; s < n; s++) {
		fill_site(c, b, s);
		v->position = b->pos[s];  /* the packer records the real coordinate */
		if (build_reads(c, b, s) != 0) { crisp_shim_batch_free(sb); return -1; }
		for (int k = 0; k < 8; k++) { maxvec[k].al = b->maxvec_al[s * 8 + k]; maxvec[k].ct = b->maxvec_ct[s * 8 + k]; }
		crisp_shim_append_site(sb, &c->reflist, -1, &c->rq, c->bd, v, maxvec, b->tid[s]);
	}
	/* compare the PACKED batch (crisp_shim_batch_pack: every array copied into one arena), which also covers the view */
	crispgpu_batch g;
	const size_t arena_cap = crisp_shim_batch_packed_bytes(sb);
	void* arena = aligned_alloc(256, arena_cap ? arena_cap : 256);
	if (crisp_shim_batch_pack(sb, arena, arena_cap, &g) != 0 || g.arena_base != arena || g.arena_bytes != arena_cap) {
		if (msg) snprintf(msg, msgcap, "crisp_shim_batch_pack failed");
		free(arena); crisp_shim_batch_free(sb);
		return -2;
	}
#define CMP(name, bytes)                                                                                         \
	do {                                                                                                         \
		if ((bytes) && (b->name) && memcmp(g.name, b->name, (bytes)) != 0) {                                     \
			if (!bad && msg) snprintf(msg, msgcap, "array %s differs", #name);                                   \
			bad++;                                                                                               \
		}                                                                                                        \
	} while (0)
	if (g.n_sites != n) { bad++; if (msg) snprintf(msg, msgcap, "site count %d vs %d", g.n_sites, n); }
	CMP(tid, (size_t)n * 4); CMP(pos, (size_t)n * 4); CMP(refbase, (size_t)n); CMP(maxvec_al, (size_t)n * 8); CMP(maxvec_ct, (size_t)n * 32);
	CMP(counts, (size_t)n * 96);
	if (!compact) CMP(indcounts, (size_t)n * P * 96);
	else if (b->ic_off) {
		if (g.indcounts != NULL) { bad++; if (msg) snprintf(msg, msgcap, "compact view still exposes indcounts"); }
		CMP(ic_off, ((size_t)n * P + 1) * 4);
		if (n && g.ic_off[(size_t)n * P] == b->ic_off[(size_t)n * P]) CMP(ic_sparse, (size_t)b->ic_off[(size_t)n * P] * 4);
	} CMP(indel_len, (size_t)n * 16); CMP(hplen, (size_t)n * 16);
	CMP(read_off, ((size_t)n * P + 1) * 4);
	if (n && !compact) CMP(reads, (size_t)b->read_off[(size_t)n * P] * 2);
	if (compact) {
		if (g.reads != NULL) { bad++; if (msg) snprintf(msg, msgcap, "compact view still exposes reads"); }
		CMP(bin_off, ((size_t)n * P + 1) * 4);
		if (n && g.bin_off[(size_t)n * P] == b->bin_off[(size_t)n * P]) CMP(bins, (size_t)b->bin_off[(size_t)n * P] * 4);
	}
	if (b->alt_off) {
		CMP(alt_off, ((size_t)n * 5 + 1) * 4);
		if (n && g.alt_off[n * 5] == b->alt_off[n * 5]) CMP(alt_reads, (size_t)b->alt_off[n * 5] * sizeof(crispgpu_altread));
	}
	if (b->amb_idx) {
		for (int k = 0; k < n * 5 && bad == 0; k++) {
			const int ra = b->amb_idx[k], rb = g.amb_idx[k];
			if ((ra < 0) != (rb < 0)) { bad++; if (msg) snprintf(msg, msgcap, "amb_idx[%d]: %d vs %d", k, rb, ra); break; }
			if (ra < 0) continue;
			if (memcmp(b->amb_dec + (size_t)ra * P * 4, g.amb_dec + (size_t)rb * P * 4, (size_t)P * 16) != 0) { bad++; if (msg) snprintf(msg, msgcap, "amb_dec row of slot %d differs", k); break; }
			if (b->amb_ep)
				for (int q = 0; q < P * 2; q++)
					if (fabs(b->amb_ep[(size_t)ra * P * 2 + q] - g.amb_ep[(size_t)rb * P * 2 + q]) > 1e-12) { bad++; if (msg) snprintf(msg, msgcap, "amb_ep row of slot %d differs", k); break; }
		}
	}
	CMP(qhigh, (size_t)n * P * 128); CMP(stats, (size_t)n * P * 128); CMP(tstats, (size_t)n * 128);
#undef CMP
	free(arena);
	crisp_shim_batch_free(sb);
	return bad;
}


/*
 * Host cost of the production packer: rebuild the reference's structures for every site of `b` and time ONLY the
 * crisp_shim_append_site calls (one host core), in the per-read format (compact = 0) or the binned one (compact = 1).
 * Returns seconds spent in the packer for the batch's n_sites, or a negative value on failure.
 */
#include <time.h>
double crisp_ref_shim_pack_seconds(void* h, const crispgpu_batch* b, int compact)
{
	ref_ctx* c = (ref_ctx*)h;
	struct VARIANT* v = &c->variant;
	const int P = c->P, n = b->n_sites;
	apply_config(c);
	crisp_shim_batch* sb = crisp_shim_batch_new(P, b->qhigh != NULL, b->stats != NULL);
	if (!sb) return -1.0;
	crisp_shim_batch_set_compact(sb, compact);
	ALLELE maxvec[8];
	double total = 0.0;
	for (int s = 0; s < n; s++) {
		fill_site(c, b, s);
		v->position = b->pos[s];
		if (build_reads(c, b, s) != 0) { crisp_shim_batch_free(sb); return -1.0; }
		for (int k = 0; k < 8; k++) { maxvec[k].al = b->maxvec_al[s * 8 + k]; maxvec[k].ct = b->maxvec_ct[s * 8 + k]; }
		struct timespec t0, t1;
		clock_gettime(CLOCK_MONOTONIC, &t0);
		const int rc = crisp_shim_append_site(sb, &c->reflist, -1, &c->rq, c->bd, v, maxvec, b->tid[s]);
		clock_gettime(CLOCK_MONOTONIC, &t1);
		if (rc != 0) { crisp_shim_batch_free(sb); return -2.0; }
		total += (t1.tv_sec - t0.tv_sec) + 1e-9 * (t1.tv_nsec - t0.tv_nsec);
	}
	crisp_shim_batch_free(sb);
	return total;
}

/* ---- exact small-table mode: FET/contables_exact.c (dead in CRISP.binary, the specification of the GPU exact mode) ---- */
double compute_pvalue_exact(int** ctable, int size, int offset, int R, int A, double prob, double** NCRtable, int st);
void compute_NCRtable(double** NCRtable);
extern int MAXN;
static double** g_ncrtable = NULL;

/* log10 p of the exact 2 x k test exactly as pvalue_contable_iter forms it (contables_exact.c:96-109), or 1.0 when the
 * table fails its gate (ones < 70, max column < MAXN, 2 <= size < 8) */
double crisp_ref_exact_2xk(const int* N, const int* A, int size)
{
	if (!g_ncrtable) {
		g_ncrtable = (double**)calloc(MAXN, sizeof(double*));
		for (int i = 0; i < MAXN; i++) g_ncrtable[i] = (double*)calloc(MAXN, sizeof(double));
		compute_NCRtable(g_ncrtable);
	}
	int* rows = (int*)calloc((size_t)size * 2, sizeof(int));
	int** ctable = (int**)calloc(size, sizeof(int*));
	int ones = 0, reads = 0, maxcol = 0;
	double cl = 0;
	for (int i = 0; i < size; i++) {
		ctable[i] = rows + 2 * i;
		ctable[i][0] = N[i];
		ctable[i][1] = A[i];
		cl += ncr(N[i], A[i]);
		ones += A[i];
		reads += N[i];
		if (N[i] > maxcol) maxcol = N[i];
	}
	double out = 1.0;
	if (ones < 70 && maxcol < MAXN && size >= 2 && size < 8) {
		double p = compute_pvalue_exact(ctable, size, 0, reads, ones, cl, g_ncrtable, 0);
		out = p - ncr(reads, ones);
	}
	free(ctable);
	free(rows);
	return out;
}


/* ---- candidate screen: the reference's own sort_allelecounts (parsebam/allelecounts.c:134-183) on a batch's counts ---- */
int sort_allelecounts(struct VARIANT* variant, ALLELE* maxvec, int refbase);

void crisp_ref_sort_allelecounts(void* h, const crispgpu_batch* b, uint8_t* maxvec_al, int32_t* maxvec_ct)
{
	ref_ctx* c = (ref_ctx*)h;
	struct VARIANT* v = &c->variant;
	apply_config(c);
	ALLELE maxvec[8];
	for (int s = 0; s < b->n_sites; s++) {
		fill_site(c, b, s);
		for (int k = 0; k < 8; k++) { maxvec[k].ct = 0; maxvec[k].al = k; }  /* variantcalls.c:60 */
		sort_allelecounts(v, maxvec, v->refbase);
		for (int k = 0; k < 8; k++) { maxvec_al[(size_t)s * 8 + k] = (uint8_t)maxvec[k].al; maxvec_ct[(size_t)s * 8 + k] = maxvec[k].ct; }
	}
}
