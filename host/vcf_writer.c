/*
 * vcf_writer.c — batch VCF record writer.  See vcf_writer.h.
 *
 * The text of a record is specified by the reference's printer (crisp/newcrispprint.c): the column layout and every
 * printf conversion below are the reference's — "%d", "%.1f", "%0.5f", "%0.2f:%0.2f", (int) truncation of 10 * delta,
 * (int)(QV + 0.5) — so that the same libc renders the same bytes.  What differs is the shape of the work: the reference
 * formats one site at a time into a FILE* from a struct VARIANT; this writer formats the called sites of a whole batch
 * from the engine's flat result arrays, on several threads, into one contiguous buffer in coordinate order (each thread
 * owns a contiguous range of sites; the ranges are concatenated), which the caller writes (or compresses) in one go.
 *
 * Scope: substitution alleles (alleles 0..3) — the sites the device pileup produces.  A site whose called alleles include
 * an insertion or deletion allele needs the front end's allele strings and homopolymer lengths (print_VCF_first_5cols,
 * newcrispprint.c:25-96); such a site makes the call fail with -5 rather than being printed wrongly.
 */
#define _GNU_SOURCE
#include "vcf_writer.h"

#include <ctype.h>
#include <pthread.h>
#include <stdarg.h>
#include <stdio.h>
#include <stdlib.h>
#include <string.h>

typedef struct {
	char* p;
	size_t len, cap;
	int error;
} sbuf;

static void sb_reserve(sbuf* b, size_t extra)
{
	if (b->error || b->len + extra + 1 <= b->cap) return;
	size_t cap = b->cap ? b->cap * 2 : 1 << 16;
	while (cap < b->len + extra + 1) cap *= 2;
	char* n = (char*)realloc(b->p, cap);
	if (!n) { b->error = 1; return; }
	b->p = n;
	b->cap = cap;
}

static void sb_printf(sbuf* b, const char* fmt, ...)
{
	sb_reserve(b, 512);
	if (b->error) return;
	va_list ap;
	va_start(ap, fmt);
	const int n = vsnprintf(b->p + b->len, b->cap - b->len, fmt, ap);
	va_end(ap);
	if (n < 0 || (size_t)n >= b->cap - b->len) { b->error = 1; return; }
	b->len += (size_t)n;
}

static void sb_putc(sbuf* b, char c)
{
	sb_reserve(b, 1);
	if (b->error) return;
	b->p[b->len++] = c;
}

/* decimal without printf: the per-pool genotype columns are most of a record's text */
static void sb_int(sbuf* b, int v)
{
	char t[16];
	int n = 0;
	unsigned u = v < 0 ? 0u - (unsigned)v : (unsigned)v;
	do { t[n++] = (char)('0' + u % 10); u /= 10; } while (u);
	sb_reserve(b, 12);
	if (b->error) return;
	if (v < 0) b->p[b->len++] = '-';
	while (n) b->p[b->len++] = t[--n];
}

typedef struct {
	const crisp_vcf_sites* st;
	const crispgpu_results* res;
	const crisp_vcf_options* opt;
	int s0, s1;
	sbuf out;
	int records;
	int error;     /* -5: a called indel allele */
} job_t;

static const char ALLELE_CHAR[4] = {'A', 'C', 'G', 'T'};
static const char* const FILTER_STRINGS[5] = {"EMpass", "EMpass1", "EMpass2", "EMpass_indel", "EMfail"};

static int ref_char(const crisp_vcf_sites* st, long pos)
{
	const long k = pos - st->ref_start;
	return (k >= 0 && k < st->ref_len) ? (unsigned char)st->refseq[k] : 'N';
}

/* one record: print_pooledvariant (newcrispprint.c:347-476) on flat arrays */
static void format_site(job_t* jb, int s)
{
	const crisp_vcf_sites* st = jb->st;
	const crispgpu_results* r = jb->res;
	const crisp_vcf_options* opt = jb->opt;
	const int P = st->n_pools, em = opt->em_flag;
	const int nvar = r->varalleles[s];
	if (nvar < 1) return;
	sbuf* b = &jb->out;
	/* the called alleles in the order of variant->alleles[] / crispvar[] (host/crisp_shim.c: crisp_shim_fill_variant) */
	int slot_of[5] = {-1, -1, -1, -1, -1};
	for (int k = 0; k < 5; k++) {
		const int o = s * 5 + k, j = r->call_rank[o];
		if (r->status[o] != CRISPGPU_SLOT_CALLED || j < 0 || j >= 5) continue;
		slot_of[j] = o;
	}
	int al[5];
	for (int j = 0; j < nvar; j++) {
		if (slot_of[j] < 0) { jb->error = -1; return; }
		al[j] = r->allele[slot_of[j]];
		if (al[j] >= 4) { jb->error = -5; return; }
	}
	const int pos = st->pos[s], refbase = st->refbase[s];
	const int refraw = ref_char(st, pos), refb = toupper(refraw);
	const int iupac = !(refb == 'A' || refb == 'C' || refb == 'G' || refb == 'T' || refb == 'N');   /* init_variant.c:109 */
	const int32_t* counts = st->counts + (size_t)s * 24;
	const int32_t* ic = st->indcounts + (size_t)s * P * 24;
	const int32_t* depths = st->readdepths + (size_t)s * P;
	const int32_t* mq = st->mqcounts + (size_t)s * 4;
	/* columns 1-5 (print_VCF_first_5cols, substitutions only: type "SNV") */
	sb_printf(b, "%s\t%d\t.\t%c\t%c", st->chrom, pos + 1, refb, ALLELE_CHAR[al[0]]);
	for (int j = 1; j < nvar; j++) { sb_putc(b, ','); sb_putc(b, ALLELE_CHAR[al[j]]); }
	/* QUAL */
	double best = 0;
	for (int j = 0; j < nvar; j++) {
		const double sc = em ? r->delta[slot_of[j]] * 10 : r->ctpval[slot_of[0]] * (-10);
		if (sc > best) best = sc;
	}
	sb_printf(b, "\t%d\t", (int)best);
	/* FILTER */
	int K = 0, depth[3] = {0, 0, 0};
	for (int i = 0; i < P; i++) {
		K += opt->ploidy[i];
		const int32_t* row = ic + (size_t)i * 24;
		depth[0] += row[refbase]; depth[1] += row[refbase + 8]; depth[2] += row[refbase + 16];
		for (int j = 0; j < nvar; j++) { depth[0] += row[al[j]]; depth[1] += row[al[j] + 8]; depth[2] += row[al[j] + 16]; }
	}
	const int lowcov = depth[0] + depth[1] + depth[2] <= 2 * K;
	const double totalcount = mq[0] + mq[1] + mq[2] + mq[3];
	int mqflag = 0;
	if (mq[0] + mq[1] > 0.2 * totalcount) mqflag = 20;
	else if (mq[0] + mq[1] > 0.1 * totalcount) mqflag = 10;
	const int strandbias = r->strandpass[s] > 0 && em == 0;
	int filters = 0;
	if (mqflag > 0) filters++;
	if (strandbias) filters++;
	else if (lowcov) filters++;
	if (mqflag > 0) {
		sb_printf(b, "LowMQ%d", mqflag);
		if (strandbias) sb_printf(b, ";StrandBias");
		if (lowcov) sb_printf(b, ";LowDepth");
	} else if (strandbias) {
		sb_printf(b, "StrandBias");
		if (lowcov) sb_printf(b, ";LowDepth");
	} else if (lowcov) sb_printf(b, "LowDepth");
	int varfilters[5], pass_variant = 0;
	for (int j = 0; j < nvar; j++) {
		const double deltaf = r->deltaf[slot_of[j]], delta = r->delta[slot_of[j]];
		if (deltaf < 3.32 && delta >= 2) varfilters[j] = 0;
		else if (deltaf < 5.42 && delta >= 5) varfilters[j] = 1;
		else if (deltaf < 5.42 && delta >= 10) varfilters[j] = 2;
		else if (deltaf < 5.42 && delta >= 5 && al[j] >= 4) varfilters[j] = 3;
		else varfilters[j] = 4;
		if (varfilters[j] != 4) pass_variant++;
	}
	if (filters == 0 && pass_variant > 0) sb_printf(b, "PASS");
	else if (filters == 0) sb_printf(b, ".");
	/* INFO */
	sb_printf(b, "\tNP=%d;DP=%d,%d,%d;VT=%s;CT=", P, depth[0], depth[1], depth[2], "SNV");
	for (int j = 0; j < nvar; j++) {
		sb_printf(b, "%.1f", r->ctpval[slot_of[j]]);
		sb_printf(b, j < nvar - 1 ? "," : ";VP=");
	}
	for (int j = 0; j < nvar; j++) {
		sb_int(b, em == 1 ? r->varpools[slot_of[j]] : r->variantpools[slot_of[j]]);
		if (j < nvar - 1) sb_putc(b, ',');
	}
	if (em == 1) {
		sb_printf(b, ";VF=");
		for (int j = 0; j < nvar; j++) { sb_printf(b, "%s", FILTER_STRINGS[varfilters[j]]); if (j < nvar - 1) sb_putc(b, ','); }
		sb_printf(b, ";AC=");
		for (int j = 0; j < nvar; j++) { sb_int(b, r->allelecount[slot_of[j]]); if (j < nvar - 1) sb_putc(b, ','); }
		sb_printf(b, ";AF=");
		for (int j = 0; j < nvar; j++) { sb_printf(b, "%0.5f", r->af_ml[slot_of[j]]); if (j < nvar - 1) sb_putc(b, ','); }
		sb_printf(b, ";EMstats=");
		for (int j = 0; j < nvar; j++) { sb_printf(b, "%0.2f:%0.2f", r->delta[slot_of[j]], r->deltaf[slot_of[j]]); if (j < nvar - 1) sb_putc(b, ','); }
		sb_printf(b, ";HWEstats=");
		for (int j = 0; j < nvar; j++) { sb_printf(b, "%0.1f", -10 * r->hwe[slot_of[j]]); if (j < nvar - 1) sb_putc(b, ','); }
	}
	sb_printf(b, ";MQS=%d,%d,%d,%d;", mq[0], mq[1], mq[2], mq[3]);
	/* print_flanking_sequence, substitution branch (newcrispprint.c:333-344) */
	sb_printf(b, "FLANKSEQ=");
	if ((long)pos - 11 >= 0 && (long)pos + 11 <= st->chrom_len) {
		for (int j = -10; j < 0; j++) sb_putc(b, (char)tolower(ref_char(st, (long)pos + j)));
		sb_putc(b, ':');
		sb_putc(b, (char)toupper(refraw));
		sb_putc(b, ':');
		for (int j = 1; j < 11; j++) sb_putc(b, (char)tolower(ref_char(st, (long)pos + j)));
	} else sb_putc(b, '.');
	if (iupac) sb_printf(b, ";IUPAC=%c", refraw);
	/* genotype columns */
	const int bidir = counts[refbase + 16] + counts[al[0] + 16] >= 2;
	const int32_t* ac[5];
	const double* qv[5];
	for (int j = 0; j < nvar; j++) {
		const int row = r->call_row[slot_of[j]];
		ac[j] = (em && row >= 0) ? r->ac + (size_t)row * P : NULL;
		qv[j] = (em && row >= 0) ? r->qv + (size_t)row * P : NULL;
	}
#define AC(j, i) (ac[j] ? ac[j][i] : 0)
#define QV(j, i) (qv[j] ? qv[j][i] : 0.0)
	if (em == 0) {
		/* print_VCF_allelefreqs (newcrispprint.c:99-131) */
		sb_printf(b, "\tGT:AF:DP:ADf:ADr");
		for (int i = 0; i < P; i++) {
			const int32_t* row = ic + (size_t)i * 24;
			int coverage = row[refbase] + row[refbase + 8] + row[refbase + 16];
			for (int j = 0; j < nvar; j++) coverage += row[al[j]] + row[al[j] + 8] + row[al[j] + 16];
			sb_printf(b, "\t0/0:%0.3f", (float)(row[al[0]] + row[al[0] + 8] + row[al[0] + 16]) / (coverage + 0.001));
			for (int j = 1; j < nvar; j++) sb_printf(b, ",%0.3f", (float)(row[al[j]] + row[al[j] + 8] + row[al[j] + 16]) / (coverage + 0.001));
			sb_putc(b, ':'); sb_int(b, depths[i]);
			sb_putc(b, ':'); sb_int(b, row[refbase] + row[refbase + 16] / 2);
			for (int j = 0; j < nvar; j++) { sb_putc(b, ','); sb_int(b, row[al[j]] + row[al[j] + 16] / 2); }
			sb_putc(b, ':'); sb_int(b, row[refbase + 8] + (row[refbase + 16] - row[refbase + 16] / 2));
			for (int j = 0; j < nvar; j++) { sb_putc(b, ','); sb_int(b, row[al[j] + 8] + (row[al[j] + 16] - row[al[j] + 16] / 2)); }
		}
	} else {
		const int diploid = opt->ploidy[0] == 2 || opt->ploidy[0] == 1;
		sb_printf(b, diploid ? "\tGT:GQ:DP:ADf:ADr:ADb" : "\tAC:GQ:DP:ADf:ADr:ADb");
		for (int i = 0; i < P; i++) {
			const int32_t* row = ic + (size_t)i * 24;
			sb_putc(b, '\t');
			const int nodata = (double)depths[i] / opt->ploidy[i] <= 1;
			if (diploid) {
				/* print_VCF_genotypes_diploid (newcrispprint.c:134-180) */
				const char* gt = NULL;
				int q = 0;
				if (nodata) gt = NULL;
				else if (nvar == 1) {
					const int a0 = AC(0, i);
					gt = a0 == 0 ? "0/0" : a0 == 1 ? "0/1" : a0 == 2 ? "1/1" : "";
					q = (int)(0.5 + QV(0, i));
				} else if (nvar == 2) {
					const int a0 = AC(0, i), a1 = AC(1, i);
					q = (int)(0.5 + QV(0, i));
					if (a0 == 0 && a1 == 0) gt = "0/0";
					else if (a0 == 1 && a1 == 0) gt = "0/1";
					else if (a0 == 2 && a1 == 0) gt = "1/1";
					else if (a0 == 1 && a1 == 1) gt = "1/2";
					else if (a0 == 0 && a1 == 1) { gt = "0/2"; q = (int)(0.5 + QV(1, i)); }
					else if (a0 == 0 && a1 == 2) { gt = "2/2"; q = (int)(0.5 + QV(1, i)); }
				}
				if (gt == NULL) sb_printf(b, "./.:.");
				else if (gt[0]) { sb_printf(b, "%s:", gt); sb_int(b, q); }
				/* an allele count outside 0..2 prints nothing before the depth, as the reference's if-chain does */
			} else {
				/* print_VCF_genotypes_pooled (newcrispprint.c:248-298) */
				if (nodata) sb_putc(b, '.');
				else if (opt->ploidy[0] > 2 && opt->varpoolsize > 0) {
					int k = opt->ploidy[i];
					for (int j = 0; j < nvar; j++) k -= AC(j, i);
					if (k < 0) sb_putc(b, '.');
					else {
						sb_int(b, k); sb_putc(b, ','); sb_int(b, AC(0, i));
						for (int j = 1; j < nvar; j++) { sb_putc(b, ','); sb_int(b, AC(j, i)); }
					}
				} else if (opt->ploidy[0] > 2) {
					sb_int(b, AC(0, i));
					for (int j = 1; j < nvar; j++) { sb_putc(b, ','); sb_int(b, AC(j, i)); }
				} else sb_putc(b, '.');
				sb_putc(b, ':'); sb_int(b, (int)(QV(0, i) + 0.5));
			}
			sb_putc(b, ':'); sb_int(b, depths[i]);
			sb_putc(b, ':'); sb_int(b, row[refbase]);
			for (int j = 0; j < nvar; j++) { sb_putc(b, ','); sb_int(b, row[al[j]]); }
			sb_putc(b, ':'); sb_int(b, row[refbase + 8]);
			for (int j = 0; j < nvar; j++) { sb_putc(b, ','); sb_int(b, row[al[j] + 8]); }
			if (bidir) {
				sb_putc(b, ':'); sb_int(b, row[refbase + 16]);
				for (int j = 0; j < nvar; j++) { sb_putc(b, ','); sb_int(b, row[al[j] + 16]); }
			} else {
				sb_putc(b, ':'); sb_putc(b, '0');
				for (int j = 0; j < nvar; j++) { sb_putc(b, ','); sb_putc(b, '0'); }
			}
		}
	}
#undef AC
#undef QV
	sb_putc(b, '\n');
	jb->records++;
}

static void* worker(void* arg)
{
	job_t* jb = (job_t*)arg;
	for (int s = jb->s0; s < jb->s1 && !jb->error && !jb->out.error; s++) format_site(jb, s);
	return NULL;
}

int crisp_vcf_format(const crisp_vcf_sites* st, const crispgpu_results* res, const crisp_vcf_options* opt, char** out, size_t* out_len, int32_t* n_records)
{
	if (!st || !res || !opt || !out || !out_len || !opt->ploidy || st->n_sites != res->n_sites || st->n_pools <= 0 || !st->chrom) return -1;
	*out = NULL;
	*out_len = 0;
	if (n_records) *n_records = 0;
	const int n = st->n_sites;
	if (n > 0 && (!st->pos || !st->refbase || !st->counts || !st->indcounts || !st->readdepths || !st->mqcounts || !st->refseq)) return -1;
	int threads = opt->threads > 0 ? opt->threads : 1;
	if (threads > 64) threads = 64;
	if (threads > n) threads = n > 0 ? n : 1;
	job_t jobs[64];
	pthread_t th[64];
	memset(jobs, 0, sizeof(job_t) * (size_t)threads);
	for (int t = 0; t < threads; t++) {
		jobs[t].st = st; jobs[t].res = res; jobs[t].opt = opt;
		jobs[t].s0 = (int)((long long)n * t / threads);
		jobs[t].s1 = (int)((long long)n * (t + 1) / threads);
	}
	int started = 0;
	for (int t = 1; t < threads; t++) {
		if (pthread_create(&th[t], NULL, worker, &jobs[t]) != 0) break;
		started = t;
	}
	worker(&jobs[0]);
	for (int t = 1; t <= started; t++) pthread_join(th[t], NULL);
	for (int t = started + 1; t < threads; t++) worker(&jobs[t]);   /* threads that could not be created: run here */
	int rc = 0;
	size_t total = 0;
	for (int t = 0; t < threads; t++) {
		if (jobs[t].error && !rc) rc = jobs[t].error;
		if (jobs[t].out.error && !rc) rc = -4;
		total += jobs[t].out.len;
	}
	char* buf = rc ? NULL : (char*)malloc(total + 1);
	if (!rc && !buf) rc = -4;
	size_t o = 0;
	int recs = 0;
	for (int t = 0; t < threads; t++) {
		if (!rc && jobs[t].out.len) { memcpy(buf + o, jobs[t].out.p, jobs[t].out.len); o += jobs[t].out.len; }
		recs += jobs[t].records;
		free(jobs[t].out.p);
	}
	if (rc) return rc;
	buf[total] = 0;
	*out = buf;
	*out_len = total;
	if (n_records) *n_records = recs;
	return 0;
}

void crisp_vcf_free(char* p) { free(p); }
