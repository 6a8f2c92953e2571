/*
 * crisp_frontend.c — the binding a CRISP maintainer adds to run the reference's own program on the B200 engine.
 *
 * The reference enters its statistics engine through one call,
 *     v = newCRISPcaller(reflist,current,position,bq,bamfiles_data,variant,maxvec,options->vfile);   variantcalls.c:115
 * and that callee prints the VCF record itself (newcrispcaller.c:266).  This file provides the replacement for that
 * call: the reference's objects are linked unchanged with `ld --wrap=newCRISPcaller --wrap=fclose` (oracle/Makefile,
 * target _ref/CRISP.gpu), so jointvariantcaller's call lands in __wrap_newCRISPcaller below — no source line of the
 * reference is edited or copied.  Each candidate site is snapshotted into the struct-of-arrays batch by
 * host/crisp_shim.c (the three read-queue walks of the engine) together with the handful of struct VARIANT fields
 * only the printer reads; when the batch is full — or when the VCF stream is about to be closed, or at exit — the batch
 * goes through the C ABI (crispgpu_submit / crispgpu_wait, include/crispgpu.h), the results are copied back into a
 * struct VARIANT and the reference's own print_pooledvariant (crisp/newcrispprint.c:347) writes the records, in the
 * order the sites arrived (coordinate order).
 *
 * Environment:
 *   CRISPGPU_LIB        path of libcrispgpu.so (default: <dir of this executable>/../../crisp_b200/csrc/libcrispgpu.so)
 *   CRISPGPU_DEVICE     CUDA device (default 0)
 *   CRISPGPU_BATCH      candidate sites per batch (default 4096)
 *   CRISPGPU_COMPACT    1: ship (value,count) bins and sparse counts instead of one element per read
 *                       2: ... in their narrow transport form (crispgpu_batch_pack_narrow: 16-bit entries, one pinned arena)
 *   CRISPGPU_SEED       key of the engine's counter-based RNG (permutation tests); default 1
 *   CRISPGPU_EXACT      1: exact p-value for small tables (config.exact_mode)
 *   CRISPGPU_DUMP       file: every flushed batch is appended to it in the layout tests/frontend_io.py reads
 *   CRISPGPU_DUMP_ONLY  1: do not touch the GPU (no records are printed) — used by the CPU tests that check the
 *                       batch the front end builds against the reference's own results on the same BAMs
 *   CRISPGPU_STATS      file: one line of counters/timings at exit (sites, batches, seconds in pack / engine / print)
 * There is no CPU compute path here: without CRISPGPU_DUMP_ONLY a missing library or device is a fatal error.
 */
#include <stdio.h>
#include <stdlib.h>
#include <string.h>
#include <ctype.h>
#include <dlfcn.h>
#include <time.h>
#include <unistd.h>

#include "bamread.h"
#include "bamsreader.h"
#include "variant.h"
#define CRISP_SHIM_WITH_REFERENCE_TYPES
#include "crisp_shim.h"

/* the reference's option globals that reach the engine */
extern int EMflag, MAXITER, CHISQ_PERMUTATION, FAST_FILTER, USE_BASE_QVS, MIN_READS, QVoffset, MINQ, PIVOTSAMPLE, USE_QV;
extern int VARIANTS_CALLED, maxalleles;
extern double thresh1, MINQpv, thresh3, SER, RELAX, Qmin, AGILENT_REF_BIAS, THETA;
int print_pooledvariant(struct VARIANT* variant, FILE* vfile, REFLIST* reflist, int current, ALLELE* maxvec, int EMflag);
void init_variant(struct VARIANT* variant, int bamfiles, int samples);
int __real_fclose(FILE* f);

typedef struct {
	void* dl;
	void (*config_defaults)(crispgpu_config*);
	int (*create)(const crispgpu_config*, int, void*, crispgpu_ctx**);
	void (*destroy)(crispgpu_ctx*);
	int (*submit)(crispgpu_ctx*, const crispgpu_batch*, int64_t*);
	int (*wait)(crispgpu_ctx*, int64_t, crispgpu_results*);
	const char* (*last_error)(crispgpu_ctx*);
	int (*get_timing)(crispgpu_ctx*, crispgpu_timing*);
	size_t (*narrow_bytes)(const crispgpu_batch*, int32_t);
	int (*pack_narrow)(const crispgpu_batch*, int32_t, void*, size_t, crispgpu_batch*);
	void* (*host_alloc)(size_t);
	void (*host_free)(void*);
} engine_api;

/* what print_pooledvariant reads from struct VARIANT / REFLIST beyond the batch arrays (newcrispprint.c:347-476) */
typedef struct {
	int current;       /* chromosome index */
	char refb;
	short iupac;
	int mq[4];         /* MQcounts */
	int filt[2];       /* filteredreads[0..1] (variant.c:288-289; the log prints them, the device pileup reproduces them) */
	size_t itb_off[4]; /* indel allele strings of alleles 4..7 in `text` */
	size_t win_off;    /* reference window in `text` */
	int win_start, win_len;
} site_state;

static struct {
	int ready, dump_only, batch_sites, finished, narrow;
	void* arena;          /* pinned arena of the narrow transport form */
	size_t arena_cap;
	engine_api api;
	crispgpu_ctx* ctx;
	crispgpu_config cfg;
	int* ploidy;
	int* phenotypes;
	crisp_shim_batch* sb;
	site_state* st;
	int cap_st;
	int* depths;       /* [n][P] readdepths */
	size_t cap_depths;
	char* text;
	size_t n_text, cap_text;
	struct VARIANT pv; /* scratch for the printer */
	REFLIST* reflist;
	FILE* vfile;
	FILE* dump;
	long sites, batches, printed;
	double t_pack, t_engine, t_print, t_device;
	long h2d_bytes, d2h_bytes;
} F;

static double now_s(void)
{
	struct timespec t;
	clock_gettime(CLOCK_MONOTONIC, &t);
	return t.tv_sec + 1e-9 * t.tv_nsec;
}

static void die(const char* what, const char* detail)
{
	fprintf(stderr, "crisp_frontend: %s%s%s\n", what, detail ? ": " : "", detail ? detail : "");
	exit(3);
}

static void load_engine(void)
{
	char path[4096];
	const char* env = getenv("CRISPGPU_LIB");
	if (env) snprintf(path, sizeof path, "%s", env);
	else {
		char exe[3072];
		ssize_t k = readlink("/proc/self/exe", exe, sizeof exe - 1);
		if (k <= 0) die("cannot locate the executable", NULL);
		exe[k] = 0;
		char* slash = strrchr(exe, '/');
		if (slash) *slash = 0;
		snprintf(path, sizeof path, "%s/../../crisp_b200/csrc/libcrispgpu.so", exe);
	}
	F.api.dl = dlopen(path, RTLD_NOW | RTLD_LOCAL);
	if (!F.api.dl) die("cannot load the engine library (there is no CPU path)", dlerror());
#define SYM(field, name) do { *(void**)&F.api.field = dlsym(F.api.dl, name); if (!F.api.field) die("missing symbol", name); } while (0)
	SYM(config_defaults, "crispgpu_config_defaults");
	SYM(create, "crispgpu_create");
	SYM(destroy, "crispgpu_destroy");
	SYM(submit, "crispgpu_submit");
	SYM(wait, "crispgpu_wait");
	SYM(last_error, "crispgpu_last_error");
	SYM(get_timing, "crispgpu_get_timing");
	SYM(narrow_bytes, "crispgpu_batch_narrow_bytes");
	SYM(pack_narrow, "crispgpu_batch_pack_narrow");
	SYM(host_alloc, "crispgpu_host_alloc");
	SYM(host_free, "crispgpu_host_free");
#undef SYM
}

/* snapshot of the reference's option globals (optionparser.c:17-136) as the engine's configuration */
static void fill_config(struct VARIANT* variant)
{
	crispgpu_config* c = &F.cfg;
	const int P = variant->samples;
	memset(c, 0, sizeof *c);
	F.ploidy = (int*)malloc(sizeof(int) * P);
	F.phenotypes = (int*)calloc(P, sizeof(int));
	for (int i = 0; i < P; i++) F.ploidy[i] = variant->ploidy[i];
	c->abi_version = CRISPGPU_ABI_VERSION;
	c->n_pools = P;
	c->ploidy = F.ploidy;
	c->varpoolsize = variant->varpoolsize;
	c->em_flag = EMflag;
	c->max_iter = MAXITER;
	c->chisq_permutation = CHISQ_PERMUTATION;
	c->fast_filter = FAST_FILTER;
	c->use_base_qvs = USE_BASE_QVS;
	c->min_reads = MIN_READS;
	c->qv_offset = QVoffset;
	c->min_q = MINQ;
	c->pivot_sample = PIVOTSAMPLE;
	c->association = variant->options ? variant->options->association : 0;
	if (c->association && variant->options->phenotypes)
		for (int i = 0; i < P; i++) F.phenotypes[i] = variant->options->phenotypes[i];
	c->phenotypes = F.phenotypes;
	c->thresh1 = thresh1;
	c->min_qpv = MINQpv;
	c->thresh3 = thresh3;
	c->ser = SER;
	c->relax = RELAX;
	c->qmin = Qmin;
	c->agilent_ref_bias = AGILENT_REF_BIAS;
	c->theta = THETA;
	c->use_qv = USE_QV;
	c->minimal_results = 1;  /* the printer reads no diagnostic field */
	c->rng_seed = getenv("CRISPGPU_SEED") ? strtoull(getenv("CRISPGPU_SEED"), NULL, 10) : 1;
	c->exact_mode = getenv("CRISPGPU_EXACT") ? atoi(getenv("CRISPGPU_EXACT")) : 0;
}

static void frontend_finish(void);

static void frontend_init(struct VARIANT* variant, REFLIST* reflist, FILE* vfile)
{
	const int P = variant->samples;
	F.dump_only = getenv("CRISPGPU_DUMP_ONLY") && atoi(getenv("CRISPGPU_DUMP_ONLY"));
	F.batch_sites = getenv("CRISPGPU_BATCH") ? atoi(getenv("CRISPGPU_BATCH")) : 4096;
	if (F.batch_sites < 1) F.batch_sites = 1;
	fill_config(variant);
	if (getenv("CRISPGPU_DUMP")) {
		F.dump = fopen(getenv("CRISPGPU_DUMP"), "wb");
		if (!F.dump) die("cannot open CRISPGPU_DUMP", getenv("CRISPGPU_DUMP"));
	}
	if (!F.dump_only) {
		load_engine();
		const int dev = getenv("CRISPGPU_DEVICE") ? atoi(getenv("CRISPGPU_DEVICE")) : 0;
		const int rc = F.api.create(&F.cfg, dev, NULL, &F.ctx);
		if (rc != 0) die("crispgpu_create failed (no usable B200; there is no CPU path)", NULL);
	}
	/* em_flag 0 reads stats/tstats (chernoff_bound); the permutation mode reads Qhighcounts (chi2pvalue) */
	F.sb = crisp_shim_batch_new(P, CHISQ_PERMUTATION != 0, EMflag == 0);
	if (!F.sb) die("out of memory", NULL);
	crisp_shim_batch_set_compact(F.sb, getenv("CRISPGPU_COMPACT") && atoi(getenv("CRISPGPU_COMPACT")));
	F.narrow = getenv("CRISPGPU_COMPACT") && atoi(getenv("CRISPGPU_COMPACT")) == 2;
	memset(&F.pv, 0, sizeof F.pv);
	F.pv.ploidy = (int*)calloc(P, sizeof(int));
	for (int i = 0; i < P; i++) F.pv.ploidy[i] = variant->ploidy[i];
	init_variant(&F.pv, variant->bamfiles, P);
	F.pv.varpoolsize = variant->varpoolsize;
	F.pv.options = variant->options;
	F.reflist = reflist;
	F.vfile = vfile;
	F.ready = 1;
	atexit(frontend_finish);
}

static void put_bytes(FILE* f, const void* p, size_t n)
{
	static const char zero[8] = {0};
	uint64_t len = n;
	fwrite(&len, 8, 1, f);
	if (n) fwrite(p, 1, n, f);
	if (n & 7) fwrite(zero, 1, 8 - (n & 7), f);
}

/* one record per flushed batch: magic, config scalars, then every array as (u64 length, bytes, pad to 8) in the order
 * of crispgpu_batch; tests/frontend_io.py is the reader */
static void dump_batch(const crispgpu_batch* b)
{
	FILE* f = F.dump;
	const size_t n = (size_t)b->n_sites, P = (size_t)F.cfg.n_pools;
	const char magic[8] = {'C', 'R', 'S', 'P', 'B', 'T', 'C', '1'};
	fwrite(magic, 8, 1, f);
	int32_t head[16] = {b->n_sites, F.cfg.n_pools, b->n_amb, F.cfg.em_flag, F.cfg.max_iter, F.cfg.chisq_permutation, F.cfg.fast_filter,
	                    F.cfg.use_base_qvs, F.cfg.min_reads, F.cfg.qv_offset, F.cfg.min_q, F.cfg.pivot_sample, F.cfg.varpoolsize,
	                    F.cfg.use_qv, F.cfg.association, 0};
	fwrite(head, sizeof head, 1, f);
	double dhead[8] = {F.cfg.thresh1, F.cfg.min_qpv, F.cfg.thresh3, F.cfg.ser, F.cfg.relax, F.cfg.qmin, F.cfg.agilent_ref_bias, F.cfg.theta};
	fwrite(dhead, sizeof dhead, 1, f);
	put_bytes(f, F.ploidy, P * 4);
	put_bytes(f, b->tid, n * 4);
	put_bytes(f, b->pos, n * 4);
	put_bytes(f, b->refbase, n);
	put_bytes(f, b->maxvec_al, n * 8);
	put_bytes(f, b->maxvec_ct, n * 32);
	put_bytes(f, b->counts, n * 96);
	put_bytes(f, F.sb->indcounts, n * P * 96);
	put_bytes(f, b->indel_len, n * 16);
	put_bytes(f, b->hplen, n * 16);
	put_bytes(f, b->read_off, (n * P + 1) * 4);
	put_bytes(f, b->reads, b->reads ? (size_t)b->read_off[n * P] * 2 : 0);
	put_bytes(f, b->alt_off, (n * 5 + 1) * 4);
	put_bytes(f, b->alt_reads, (size_t)b->alt_off[n * 5] * sizeof(crispgpu_altread));
	put_bytes(f, b->amb_idx, n * 20);
	put_bytes(f, b->amb_dec, (size_t)b->n_amb * P * 16);
	put_bytes(f, b->amb_ep, (size_t)b->n_amb * P * 16);
	put_bytes(f, b->qhigh, b->qhigh ? n * P * 128 : 0);
	put_bytes(f, b->stats, b->stats ? n * P * 128 : 0);
	put_bytes(f, b->tstats, b->tstats ? n * 128 : 0);
	put_bytes(f, b->bin_off, b->bins ? (n * P + 1) * 4 : 0);
	put_bytes(f, b->bins, b->bins ? (size_t)b->bin_off[n * P] * 4 : 0);
	put_bytes(f, b->ic_off, b->ic_sparse ? (n * P + 1) * 4 : 0);
	put_bytes(f, b->ic_sparse, b->ic_sparse ? (size_t)b->ic_off[n * P] * 4 : 0);
	/* printer state, so that a test can replay the records: readdepths, MQcounts, refb; filteredreads[0..1] */
	put_bytes(f, F.depths, n * P * 4);
	int32_t* mq = (int32_t*)malloc(n * 16 + 16);
	char* rb = (char*)malloc(n + 1);
	for (size_t s = 0; s < n; s++) { memcpy(mq + s * 4, F.st[s].mq, 16); rb[s] = F.st[s].refb; }
	put_bytes(f, mq, n * 16);
	put_bytes(f, rb, n);
	for (size_t s = 0; s < n; s++) { mq[s * 2] = F.st[s].filt[0]; mq[s * 2 + 1] = F.st[s].filt[1]; }
	put_bytes(f, mq, n * 8);
	free(mq);
	free(rb);
	fflush(f);
}

static void print_site(const crispgpu_batch* b, const crispgpu_results* r, int s)
{
	struct VARIANT* v = &F.pv;
	const site_state* st = &F.st[s];
	const int P = F.cfg.n_pools;
	ALLELE maxvec[8];
	v->position = b->pos[s];
	v->refbase = b->refbase[s];
	v->refb = st->refb;
	v->IUPAC = st->iupac;
	strcpy(v->chrom, F.reflist->names[st->current]);
	for (int a = 4; a < 8; a++) strcpy(v->itb[a], F.text + st->itb_off[a - 4]);
	for (int k = 0; k < 10; k++) v->HPlength[k] = k < 8 ? b->hplen[s * 8 + k] : 0;
	for (int q = 0; q < 4 * 8; q++) v->counts[q] = q < 24 ? b->counts[s * 24 + q] : 0;
	for (int i = 0; i < P; i++) {
		const int32_t* ic = F.sb->indcounts + ((size_t)s * P + i) * 24;
		for (int q = 0; q < 4 * 8; q++) v->indcounts[i][q] = q < 24 ? ic[q] : 0;
		v->readdepths[i] = F.depths[(size_t)s * P + i];
	}
	for (int k = 0; k < 4; k++) v->MQcounts[k] = st->mq[k];
	for (int k = 0; k < 8; k++) { maxvec[k].al = b->maxvec_al[s * 8 + k]; maxvec[k].ct = b->maxvec_ct[s * 8 + k]; }
	crisp_shim_fill_variant(r, s, v);
	/* newcrispcaller.c:253-263: before printing, the reference-allele counts are reduced once more by the reads that are
	 * ambiguous within the longest homopolymer among the called indel alleles — the row packed for that slot */
	int hpmax = 0, hpslot = -1;
	for (int k = 0; k < 5; k++) {
		const int o = s * 5 + k;
		if (r->status[o] != CRISPGPU_SLOT_CALLED || r->call_rank[o] < 0) continue;
		if (r->allele[o] >= 4 && v->HPlength[r->allele[o]] > hpmax) { hpmax = v->HPlength[r->allele[o]]; hpslot = k; }
	}
	if (hpslot >= 0 && b->amb_idx[s * 5 + hpslot] >= 0) {
		const int row = b->amb_idx[s * 5 + hpslot];
		for (int i = 0; i < P; i++) {
			const int32_t* d = b->amb_dec + ((size_t)row * P + i) * 4;
			v->indcounts[i][v->refbase] -= d[0];
			v->indcounts[i][v->refbase + 8] -= d[1];
			v->indcounts[i][v->refbase + 16] -= d[2] + d[3];
		}
	}
	/* the printer indexes reflist->sequences[current][position + j]; the chromosome may already have been released
	 * (readmultiplebams.c:134), so it gets a window snapshotted when the site was packed, addressed the same way */
	REFLIST rl = *F.reflist;
	unsigned char* seqs[1];
	rl.sequences = seqs - st->current;
	seqs[0] = (unsigned char*)(F.text + st->win_off) - st->win_start;
	print_pooledvariant(v, F.vfile, &rl, st->current, maxvec, EMflag);
	F.printed++;
	VARIANTS_CALLED++;
}

static void frontend_flush(void)
{
	if (!F.ready || F.sb->n_sites == 0) return;
	crispgpu_batch b;
	crisp_shim_batch_view(F.sb, &b);
	if (F.dump) dump_batch(&b);
	if (!F.dump_only) {
		crispgpu_results r;
		int64_t ticket = 0;
		double t0 = now_s();
		/* narrow transport: the compact arrays as 16-bit entries in one pinned arena (falls back to the compact batch when a cell
		 * is too full for a length byte) */
		crispgpu_batch nb;
		const crispgpu_batch* ship = &b;
		if (F.narrow) {
			const size_t need = F.api.narrow_bytes(&b, F.cfg.n_pools);
			if (need > F.arena_cap) {
				if (F.arena) F.api.host_free(F.arena);
				F.arena_cap = need + need / 2;
				F.arena = F.api.host_alloc(F.arena_cap);
				if (!F.arena) die("pinned allocation failed", NULL);
			}
			if (need && F.api.pack_narrow(&b, F.cfg.n_pools, F.arena, F.arena_cap, &nb) == 0) ship = &nb;
		}
		int rc = F.api.submit(F.ctx, ship, &ticket);
		if (rc == 0) rc = F.api.wait(F.ctx, ticket, &r);
		if (rc != 0) die("engine call failed", F.api.last_error(F.ctx));
		crispgpu_timing tm;
		if (F.api.get_timing(F.ctx, &tm) == 0) { F.t_device += tm.total_ms * 1e-3; F.h2d_bytes += tm.h2d_bytes; F.d2h_bytes += tm.d2h_bytes; }
		double t1 = now_s();
		F.t_engine += t1 - t0;
		for (int s = 0; s < b.n_sites; s++)
			if (r.varalleles[s] >= 1) print_site(&b, &r, s);
		F.t_print += now_s() - t1;
	}
	F.batches++;
	crisp_shim_batch_clear(F.sb);
	F.n_text = 0;
}

static void frontend_finish(void)
{
	if (!F.ready || F.finished) return;
	frontend_flush();
	F.finished = 1;
	if (F.vfile) fflush(F.vfile);
	if (F.dump) __real_fclose(F.dump);
	if (getenv("CRISPGPU_STATS")) {
		FILE* f = fopen(getenv("CRISPGPU_STATS"), "w");
		if (f) {
			fprintf(f, "{\"sites\": %ld, \"batches\": %ld, \"records\": %ld, \"pack_s\": %.6f, \"engine_s\": %.6f, \"device_s\": %.6f, "
			           "\"print_s\": %.6f, \"h2d_bytes\": %ld, \"d2h_bytes\": %ld}\n",
			        F.sites, F.batches, F.printed, F.t_pack, F.t_engine, F.t_device, F.t_print, F.h2d_bytes, F.d2h_bytes);
			__real_fclose(f);
		}
	}
	if (F.ctx) { F.api.destroy(F.ctx); F.ctx = NULL; }
}

static size_t text_put(const char* p, size_t n)
{
	if (F.n_text + n + 1 > F.cap_text) {
		F.cap_text = (F.n_text + n + 1) * 2 + 4096;
		F.text = (char*)realloc(F.text, F.cap_text);
		if (!F.text) die("out of memory", NULL);
	}
	const size_t off = F.n_text;
	memcpy(F.text + off, p, n);
	F.text[off + n] = 0;
	F.n_text += n + 1;
	return off;
}

/* replaces newCRISPcaller at variantcalls.c:115 (ld --wrap).  Returns 0: the record, if any, is printed at the next
 * flush and VARIANTS_CALLED is brought up to date there. */
int __wrap_newCRISPcaller(REFLIST* reflist, int current, int position, READQUEUE* bq, struct BAMFILE_data* bd, struct VARIANT* variant,
                          ALLELE* maxvec, FILE* vfile)
{
	if (!F.ready) frontend_init(variant, reflist, vfile);
	const double t0 = now_s();
	const int P = variant->samples, s = F.sb->n_sites;
	if (crisp_shim_append_site(F.sb, reflist, current, bq, bd, variant, maxvec, current) != 0) die("cannot pack site (out of memory or value out of range)", NULL);
	if (s + 1 > F.cap_st) {
		F.cap_st = (s + 1) * 2 + 64;
		F.st = (site_state*)realloc(F.st, sizeof(site_state) * F.cap_st);
		if (!F.st) die("out of memory", NULL);
	}
	if ((size_t)(s + 1) * P > F.cap_depths) {
		F.cap_depths = (size_t)(s + 1) * P * 2 + 1024;
		F.depths = (int*)realloc(F.depths, sizeof(int) * F.cap_depths);
		if (!F.depths) die("out of memory", NULL);
	}
	site_state* st = &F.st[s];
	st->current = current;
	st->refb = variant->refb;
	st->iupac = variant->IUPAC;
	for (int k = 0; k < 4; k++) st->mq[k] = variant->MQcounts[k];
	st->filt[0] = variant->filteredreads[0]; st->filt[1] = variant->filteredreads[1];
	for (int i = 0; i < P; i++) F.depths[(size_t)s * P + i] = variant->readdepths[i];
	int reach = 12;
	for (int a = 4; a < 8; a++) {
		const size_t len = strlen(variant->itb[a]);
		st->itb_off[a - 4] = text_put(variant->itb[a], len < 1023 ? len : 1023);
		if (variant->HPlength[a] + (int)len + 12 > reach) reach = variant->HPlength[a] + (int)len + 12;
	}
	st->win_start = position - 12 < 0 ? 0 : position - 12;
	int end = position + reach;
	if (end > reflist->lengths[current]) end = reflist->lengths[current];
	st->win_len = end - st->win_start;
	st->win_off = text_put((const char*)reflist->sequences[current] + st->win_start, (size_t)st->win_len);
	F.sites++;
	F.t_pack += now_s() - t0;
	if (F.sb->n_sites >= F.batch_sites) frontend_flush();
	return 0;
}

/* the reference closes the VCF stream at the end of main (readmultiplebams.c:276): print what is pending first */
int __wrap_fclose(FILE* f)
{
	if (F.ready && !F.finished && f == F.vfile) frontend_finish();
	return __real_fclose(f);
}
