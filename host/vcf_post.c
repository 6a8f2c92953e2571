/* vcf_post.c — see vcf_post.h.  The conversion follows scripts/convert_pooled_vcf.py line by line (cited below); the BGZF
 * framing follows samtools/bgzf.c (deflate_block: header :40-60 of the format, BSIZE = block size - 1, CRC32 + ISIZE). */
#include "vcf_post.h"

#include <pthread.h>
#include <stdio.h>
#include <stdlib.h>
#include <string.h>
#include <zlib.h>

/* ---------------------------------------------------------------------------------------------------------------- */
/* growable text buffer                                                                                             */
/* ---------------------------------------------------------------------------------------------------------------- */
typedef struct {
	char* p;
	size_t n, cap;
	int oom;
} buf_t;

static void buf_reserve(buf_t* b, size_t extra)
{
	if (b->oom || b->n + extra <= b->cap) return;
	size_t cap = b->cap ? b->cap : 1 << 16;
	while (cap < b->n + extra) cap *= 2;
	char* q = (char*)realloc(b->p, cap);
	if (!q) { b->oom = 1; return; }
	b->p = q;
	b->cap = cap;
}

static void buf_put(buf_t* b, const char* s, size_t n)
{
	buf_reserve(b, n);
	if (b->oom) return;
	memcpy(b->p + b->n, s, n);
	b->n += n;
}

static void buf_putc(buf_t* b, char c) { buf_put(b, &c, 1); }

/* n copies of the allele `text` separated by '/', continuing a genotype that already holds *have alleles */
static void put_alleles(buf_t* b, const char* text, size_t tl, long n, long* have)
{
	if (n <= 0) return;
	buf_reserve(b, (size_t)n * (tl + 1));
	if (b->oom) return;
	for (long k = 0; k < n; k++) {
		if (*have) b->p[b->n++] = '/';
		memcpy(b->p + b->n, text, tl);
		b->n += tl;
		(*have)++;
	}
}

static int is_space(char c) { return c == ' ' || c == '\t' || c == '\n' || c == '\r' || c == '\v' || c == '\f'; }

/* Python's int(str) on the slice [s, e): optional blanks, optional sign, digits, optional blanks */
static int parse_int(const char* s, const char* e, long* v)
{
	while (s < e && is_space(*s)) s++;
	while (e > s && is_space(e[-1])) e--;
	if (s == e) return -1;
	int neg = 0;
	if (*s == '+' || *s == '-') { neg = *s == '-'; s++; }
	if (s == e) return -1;
	long x = 0;
	for (; s < e; s++) {
		if (*s < '0' || *s > '9') return -1;
		if (x > 100000000L) return -1;   /* no pool holds that many alleles; keeps the arithmetic in range */
		x = x * 10 + (*s - '0');
	}
	*v = neg ? -x : x;
	return 0;
}

typedef struct { const char* s; const char* e; } span_t;

/* ---------------------------------------------------------------------------------------------------------------- */
/* convert_pooled_vcf.py                                                                                            */
/* ---------------------------------------------------------------------------------------------------------------- */
int crisp_vcf_expand_genotypes(const char* vcf, size_t len, int32_t poolsize, const int32_t* pool_sizes, int32_t n_pools, char** out, size_t* out_len,
                               crisp_vcf_expand_stats* stats)
{
	if ((!vcf && len) || !out || !out_len) return -1;
	if (poolsize <= 0 && (!pool_sizes || n_pools <= 0)) return -1;   /* script :24: no pool size from either source */
	crisp_vcf_expand_stats st;
	memset(&st, 0, sizeof st);
	buf_t b = {0, 0, 0, 0}, rec = {0, 0, 0, 0};
	span_t* col = NULL;
	size_t col_cap = 0;
	int rc = 0;
	const char* p = vcf;
	const char* end = vcf + len;
	while (p < end) {
		const char* nl = (const char*)memchr(p, '\n', (size_t)(end - p));
		const char* le = nl ? nl : end;                      /* line without its newline */
		const char* next = nl ? nl + 1 : end;
		if (*p == '#') {                                     /* :31-33 header lines pass through */
			buf_put(&b, p, (size_t)(le - p));
			buf_putc(&b, '\n');
			st.n_header_lines++;
			p = next;
			continue;
		}
		st.n_records_in++;
		/* :35 var = line.strip().split('\t') */
		const char* s = p;
		const char* e = le;
		while (s < e && is_space(*s)) s++;
		while (e > s && is_space(e[-1])) e--;
		size_t nc = 0;
		for (const char* q = s;;) {
			const char* t = (const char*)memchr(q, '\t', (size_t)(e - q));
			if (nc == col_cap) {
				size_t cap = col_cap ? col_cap * 2 : 64;
				span_t* c2 = (span_t*)realloc(col, cap * sizeof(span_t));
				if (!c2) { rc = -4; goto done; }
				col = c2;
				col_cap = cap;
			}
			col[nc].s = q;
			col[nc].e = t ? t : e;
			nc++;
			if (!t) break;
			q = t + 1;
		}
		if (nc < 9) { rc = -1; goto done; }                  /* var[4] / var[8]: the script stops with an IndexError */
		/* :37-39 multi-allelic indels are left out */
		{
			const char* a0 = col[4].s;
			const char* c1 = (const char*)memchr(a0, ',', (size_t)(col[4].e - a0));
			if (c1) {
				const char* c2 = (const char*)memchr(c1 + 1, ',', (size_t)(col[4].e - (c1 + 1)));
				const size_t lref = (size_t)(col[3].e - col[3].s), l0 = (size_t)(c1 - a0), l1 = (size_t)((c2 ? c2 : col[4].e) - (c1 + 1));
				if (lref != l0 || lref != l1) { st.n_multiallelic_indel++; p = next; continue; }
			}
		}
		/* :41-79 one genotype per sample column */
		rec.n = 0;
		int valid = 1;
		for (size_t i = 9; i < nc && valid; i++) {
			const char* gs = col[i].s;
			const char* ge = col[i].e;
			const char* colon = (const char*)memchr(gs, ':', (size_t)(ge - gs));
			const char* g0e = colon ? colon : ge;            /* genotype[0] */
			long have = 0;
			const size_t mark = rec.n;
			buf_putc(&rec, '\t');
			const int missing = g0e - gs == 1 && *gs == '.';
			long size = poolsize;
			if (poolsize <= 0) {
				if ((int64_t)(i - 9) >= n_pools) { rc = -1; goto done; }   /* pool_sizes[i-9]: IndexError */
				size = pool_sizes[i - 9];
			}
			if (missing) put_alleles(&rec, ".", 1, size, &have);            /* :50-51, :66-67 */
			else if (poolsize > 0) {
				/* :53-62 counts of the ALT alleles; the reference fills the pool */
				long refsum = poolsize;
				for (const char* q = gs;;) {
					const char* t = (const char*)memchr(q, ',', (size_t)(g0e - q));
					long v;
					if (parse_int(q, t ? t : g0e, &v)) { rc = -1; goto done; }
					refsum -= v;
					if (!t) break;
					q = t + 1;
				}
				if (refsum < 0) { valid = 0; rec.n = mark; break; }
				put_alleles(&rec, "0", 1, refsum, &have);
				long a = 0;
				for (const char* q = gs;; a++) {
					const char* t = (const char*)memchr(q, ',', (size_t)(g0e - q));
					long v;
					char name[24];
					parse_int(q, t ? t : g0e, &v);
					const int nl2 = snprintf(name, sizeof name, "%ld", a + 1);
					put_alleles(&rec, name, (size_t)nl2, v, &have);
					if (!t) break;
					q = t + 1;
				}
			} else {
				/* :72-78 counts of every allele, the reference first */
				long a = 0;
				for (const char* q = gs;; a++) {
					const char* t = (const char*)memchr(q, ',', (size_t)(g0e - q));
					long v;
					char name[24];
					if (parse_int(q, t ? t : g0e, &v)) { rc = -1; goto done; }
					const int nl2 = snprintf(name, sizeof name, "%ld", a);
					put_alleles(&rec, name, (size_t)nl2, v, &have);
					if (!t) break;
					q = t + 1;
				}
				if (have > size) { valid = 0; rec.n = mark; break; }
			}
			/* :82 '/'.join(GV) + ':' + ':'.join(genotype[1:]) */
			buf_putc(&rec, ':');
			if (colon) buf_put(&rec, colon + 1, (size_t)(ge - (colon + 1)));
		}
		if (!valid) { st.n_over_poolsize++; p = next; continue; }            /* :80 */
		/* :81-83 the first FORMAT key becomes GT; "\t".join(...).strip() */
		{
			const size_t start = b.n;
			buf_put(&b, col[0].s, (size_t)(col[7].e - col[0].s));           /* var[0:8], tabs between them as they were */
			buf_put(&b, "\tGT:", 4);
			const char* colon = (const char*)memchr(col[8].s, ':', (size_t)(col[8].e - col[8].s));
			if (colon) buf_put(&b, colon + 1, (size_t)(col[8].e - (colon + 1)));
			buf_put(&b, rec.p, rec.n);
			if (!b.oom) while (b.n > start && is_space(b.p[b.n - 1])) b.n--;
			buf_putc(&b, '\n');
		}
		st.n_records_out++;
		p = next;
	}
	if (b.oom || rec.oom) { rc = -4; goto done; }
	if (!b.p) { b.p = (char*)malloc(1); if (!b.p) { rc = -4; goto done; } }
	*out = b.p;
	*out_len = b.n;
	b.p = NULL;
	if (stats) *stats = st;
done:
	free(b.p);
	free(rec.p);
	free(col);
	return rc;
}

/* ---------------------------------------------------------------------------------------------------------------- */
/* BGZF                                                                                                             */
/* ---------------------------------------------------------------------------------------------------------------- */
#define BGZF_IN 0xff00u      /* input bytes per block (bgzf.c: BGZF_BLOCK_SIZE) */
#define BGZF_MAX 0x10000u    /* a block, header and trailer included, is at most 64 KiB */
#define BGZF_HEAD 18u
#define BGZF_TAIL 8u

typedef struct {
	const unsigned char* data;
	size_t len, n_blocks;
	int level;
	unsigned char* blocks;   /* n_blocks slots of BGZF_MAX bytes */
	uint32_t* sizes;
	size_t next;
	int error;
	pthread_mutex_t mu;
} bgzf_job;

static int deflate_raw(const unsigned char* in, size_t n, int level, unsigned char* out, size_t cap, size_t* produced)
{
	z_stream zs;
	memset(&zs, 0, sizeof zs);
	if (deflateInit2(&zs, level, Z_DEFLATED, -15, 8, Z_DEFAULT_STRATEGY) != Z_OK) return -6;
	zs.next_in = (Bytef*)in;
	zs.avail_in = (uInt)n;
	zs.next_out = out;
	zs.avail_out = (uInt)cap;
	const int r = deflate(&zs, Z_FINISH);
	*produced = cap - zs.avail_out;
	deflateEnd(&zs);
	return r == Z_STREAM_END ? 0 : 1;   /* 1: did not fit */
}

static int bgzf_block(const unsigned char* in, size_t n, int level, unsigned char* out, uint32_t* size)
{
	static const unsigned char head[16] = {0x1f, 0x8b, 8, 4, 0, 0, 0, 0, 0, 0xff, 6, 0, 'B', 'C', 2, 0};
	size_t body = 0;
	int r = deflate_raw(in, n, level, out + BGZF_HEAD, BGZF_MAX - BGZF_HEAD - BGZF_TAIL, &body);
	if (r == 1) r = deflate_raw(in, n, 0, out + BGZF_HEAD, BGZF_MAX - BGZF_HEAD - BGZF_TAIL, &body);   /* incompressible: stored always fits */
	if (r) return -6;
	const uint32_t total = (uint32_t)(BGZF_HEAD + body + BGZF_TAIL);
	memcpy(out, head, 16);
	out[16] = (unsigned char)((total - 1) & 0xff);
	out[17] = (unsigned char)((total - 1) >> 8);
	const uint32_t crc = (uint32_t)crc32(crc32(0L, Z_NULL, 0), in, (uInt)n);
	unsigned char* t = out + BGZF_HEAD + body;
	for (int k = 0; k < 4; k++) { t[k] = (unsigned char)(crc >> (8 * k)); t[4 + k] = (unsigned char)((uint32_t)n >> (8 * k)); }
	*size = total;
	return 0;
}

static void* bgzf_worker(void* arg)
{
	bgzf_job* jb = (bgzf_job*)arg;
	for (;;) {
		pthread_mutex_lock(&jb->mu);
		const size_t k = jb->next++;
		pthread_mutex_unlock(&jb->mu);
		if (k >= jb->n_blocks) return NULL;
		const size_t off = k * BGZF_IN, n = jb->len - off < BGZF_IN ? jb->len - off : BGZF_IN;
		if (bgzf_block(jb->data + off, n, jb->level, jb->blocks + k * BGZF_MAX, &jb->sizes[k])) jb->error = -6;
	}
}

int crisp_bgzf_compress(const void* data, size_t len, int level, int threads, unsigned char** out, size_t* out_len)
{
	static const unsigned char eof_block[28] = {0x1f, 0x8b, 8, 4, 0, 0, 0, 0, 0, 0xff, 6, 0, 'B', 'C', 2, 0, 0x1b, 0, 3, 0, 0, 0, 0, 0, 0, 0, 0, 0};
	if ((!data && len) || !out || !out_len || level < -1 || level > 9) return -1;
	bgzf_job jb;
	memset(&jb, 0, sizeof jb);
	jb.data = (const unsigned char*)data;
	jb.len = len;
	jb.level = level;
	jb.n_blocks = (len + BGZF_IN - 1) / BGZF_IN;
	int rc = 0;
	unsigned char* res = NULL;
	if (jb.n_blocks) {
		jb.blocks = (unsigned char*)malloc(jb.n_blocks * (size_t)BGZF_MAX);
		jb.sizes = (uint32_t*)calloc(jb.n_blocks, sizeof(uint32_t));
		if (!jb.blocks || !jb.sizes) { rc = -4; goto done; }
	}
	pthread_mutex_init(&jb.mu, NULL);
	{
		int nt = threads > 0 ? threads : 1;
		if ((size_t)nt > jb.n_blocks) nt = (int)jb.n_blocks;
		if (nt <= 1) bgzf_worker(&jb);
		else {
			pthread_t* th = (pthread_t*)malloc(sizeof(pthread_t) * (size_t)nt);
			int started = 0;
			if (th)
				for (int t = 0; t < nt; t++)
					if (pthread_create(&th[started], NULL, bgzf_worker, &jb) == 0) started++;
			if (!started) bgzf_worker(&jb);
			for (int t = 0; t < started; t++) pthread_join(th[t], NULL);
			free(th);
		}
	}
	pthread_mutex_destroy(&jb.mu);
	if (jb.error) { rc = jb.error; goto done; }
	{
		size_t total = sizeof eof_block;
		for (size_t k = 0; k < jb.n_blocks; k++) total += jb.sizes[k];
		res = (unsigned char*)malloc(total);
		if (!res) { rc = -4; goto done; }
		size_t o = 0;
		for (size_t k = 0; k < jb.n_blocks; k++) { memcpy(res + o, jb.blocks + k * BGZF_MAX, jb.sizes[k]); o += jb.sizes[k]; }
		memcpy(res + o, eof_block, sizeof eof_block);
		*out = res;
		*out_len = total;
	}
done:
	free(jb.blocks);
	free(jb.sizes);
	return rc;
}

void crisp_vcf_post_free(void* p) { free(p); }
