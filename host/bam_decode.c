/*
 * bam_decode.c — parallel BGZF inflate + BAM record split into the device pileup's alignment arrays.  See bam_decode.h.
 *
 * File format facts used here (SAM/BAM specification; the reference reads the same bytes through samtools/bgzf.c and
 * bam.c): a BGZF file is a series of gzip members of at most 64 KiB whose extra field "BC" carries the member's size
 * minus one; a BAM stream is "BAM\1", l_text, text, n_ref, n_ref x (l_name, name, l_ref), then records
 * block_size, refID, pos, l_read_name, mapq, bin, n_cigar_op, flag, l_seq, next_refID, next_pos, tlen, read_name,
 * cigar (op | len << 4), seq (4-bit, high nibble first), qual.
 */
#define _GNU_SOURCE
#include "bam_decode.h"

#include <pthread.h>
#include <stdio.h>
#include <stdlib.h>
#include <string.h>
#include <time.h>
#include <zlib.h>

typedef struct {
	size_t src, dst;
	uint32_t csize, isize;
} block_t;

typedef struct {
	unsigned char* file;
	size_t file_bytes;
	block_t* blocks;
	size_t n_blocks;
	unsigned char* data;    /* inflated stream */
	size_t data_bytes;
	size_t rec_start;       /* offset of the first alignment record */
	int32_t n_refs, first_ref_len;
	char first_ref_name[64];
	/* pass 1 */
	uint64_t kept, bases;   /* bases: padded to multiples of 8 per read */
	int64_t n_records, n_flag, n_outside, n_complex;
	int error;
} pool_t;

typedef struct {
	pool_t* pools;
	int n_pools;
	const crisp_bam_options* opt;
	/* work queues */
	size_t* item_pool;      /* inflate items: pool of item k */
	size_t* item_block;
	size_t n_items;
	size_t next_item;
	int next_pool;
	int phase;
	pthread_mutex_t mu;
	/* pass 2 targets */
	crispgpu_alnrec* rec;
	int32_t* mate_pos;
	int32_t* isize;
	uint8_t* seq4;
	uint8_t* qual;
	uint64_t* pool_read0;   /* first read / first base of each pool */
	uint64_t* pool_base0;
	int error;
} job_t;

static double now_s(void)
{
	struct timespec ts;
	clock_gettime(CLOCK_MONOTONIC, &ts);
	return (double)ts.tv_sec + 1e-9 * (double)ts.tv_nsec;
}

static uint32_t rd32(const unsigned char* p) { return (uint32_t)p[0] | ((uint32_t)p[1] << 8) | ((uint32_t)p[2] << 16) | ((uint32_t)p[3] << 24); }
static uint16_t rd16(const unsigned char* p) { return (uint16_t)(p[0] | (p[1] << 8)); }

static int read_file(const char* path, pool_t* pl)
{
	FILE* f = fopen(path, "rb");
	if (!f) return -2;
	if (fseek(f, 0, SEEK_END) != 0) { fclose(f); return -2; }
	const long n = ftell(f);
	if (n < 0 || fseek(f, 0, SEEK_SET) != 0) { fclose(f); return -2; }
	pl->file = (unsigned char*)malloc((size_t)n + 1);
	if (!pl->file) { fclose(f); return -4; }
	if (fread(pl->file, 1, (size_t)n, f) != (size_t)n) { fclose(f); return -2; }
	fclose(f);
	pl->file_bytes = (size_t)n;
	return 0;
}

/* walk the gzip members: sizes come from the BC extra field and the ISIZE trailer, nothing is inflated here */
static int index_blocks(pool_t* pl)
{
	size_t cap = pl->file_bytes / 16384 + 16, n = 0, o = 0, dst = 0;
	pl->blocks = (block_t*)malloc(cap * sizeof(block_t));
	if (!pl->blocks) return -4;
	while (o + 18 <= pl->file_bytes) {
		const unsigned char* h = pl->file + o;
		if (h[0] != 0x1f || h[1] != 0x8b || h[2] != 8 || !(h[3] & 4)) return -3;
		const uint32_t xlen = rd16(h + 10);
		if (o + 12 + xlen > pl->file_bytes) return -3;
		uint32_t bsize = 0, x = 0;
		while (x + 4 <= xlen) {
			const unsigned char* sf = h + 12 + x;
			const uint32_t slen = rd16(sf + 2);
			if (sf[0] == 'B' && sf[1] == 'C' && slen == 2) bsize = (uint32_t)rd16(sf + 4) + 1;
			x += 4 + slen;
		}
		if (bsize < 12 + xlen + 8 || o + bsize > pl->file_bytes) return -3;
		if (n == cap) {
			cap *= 2;
			block_t* nb = (block_t*)realloc(pl->blocks, cap * sizeof(block_t));
			if (!nb) return -4;
			pl->blocks = nb;
		}
		block_t* b = &pl->blocks[n++];
		b->src = o + 12 + xlen;
		b->csize = bsize - 12 - xlen - 8;
		b->isize = rd32(h + bsize - 4);
		b->dst = dst;
		dst += b->isize;
		o += bsize;
	}
	if (o != pl->file_bytes) return -3;
	pl->n_blocks = n;
	pl->data_bytes = dst;
	pl->data = (unsigned char*)malloc(dst + 16);
	return pl->data ? 0 : -4;
}

static int inflate_block(const pool_t* pl, const block_t* b)
{
	if (b->isize == 0) return 0;
	z_stream zs;
	memset(&zs, 0, sizeof zs);
	if (inflateInit2(&zs, -15) != Z_OK) return -4;
	zs.next_in = pl->file + b->src;
	zs.avail_in = b->csize;
	zs.next_out = pl->data + b->dst;
	zs.avail_out = b->isize;
	const int rc = inflate(&zs, Z_FINISH);
	const int ok = rc == Z_STREAM_END && zs.avail_out == 0;
	inflateEnd(&zs);
	if (!ok) return -3;
	if ((uint32_t)crc32(crc32(0L, Z_NULL, 0), pl->data + b->dst, b->isize) != rd32(pl->file + b->src + b->csize)) return -3;
	return 0;
}

static int parse_header(pool_t* pl)
{
	const unsigned char* d = pl->data;
	const size_t n = pl->data_bytes;
	if (n < 12 || memcmp(d, "BAM\1", 4) != 0) return -3;
	size_t o = 4;
	const uint32_t l_text = rd32(d + o);
	o += 4 + (size_t)l_text;
	if (o + 4 > n) return -3;
	const uint32_t n_ref = rd32(d + o);
	o += 4;
	pl->n_refs = (int32_t)n_ref;
	for (uint32_t r = 0; r < n_ref; r++) {
		if (o + 4 > n) return -3;
		const uint32_t l_name = rd32(d + o);
		o += 4;
		if (o + l_name + 4 > n) return -3;
		if (r == 0) {
			const size_t c = l_name < sizeof pl->first_ref_name ? l_name : sizeof pl->first_ref_name - 1;
			memcpy(pl->first_ref_name, d + o, c);
			pl->first_ref_name[sizeof pl->first_ref_name - 1] = 0;
			pl->first_ref_len = (int32_t)rd32(d + o + l_name);
		}
		o += l_name + 4;
	}
	pl->rec_start = o;
	return 0;
}

/* One record against the filters and the scope of the device pileup.  Returns 1 when kept and fills the fields. */
typedef struct {
	int32_t pos, mpos, tlen;
	uint32_t l_seq, clip, mlen;
	uint8_t mapq, strand;
	const unsigned char* seq;
	const unsigned char* qual;
} rec_view;

static int view_record(const unsigned char* r, uint32_t block_size, const crisp_bam_options* opt, rec_view* v, pool_t* stats)
{
	const int32_t ref_id = (int32_t)rd32(r);
	const int32_t pos = (int32_t)rd32(r + 4);
	const uint32_t l_name = r[8];
	const uint32_t n_cig = rd16(r + 12), flag = rd16(r + 14), l_seq = rd32(r + 16);
	if (32 + l_name + 4 * n_cig + (l_seq + 1) / 2 + l_seq > block_size) { if (stats) stats->error = -3; return 0; }
	if (stats) stats->n_records++;
	if (flag & opt->filter_mask) { if (stats) stats->n_flag++; return 0; }
	if (opt->tid >= 0 && ref_id != opt->tid) { if (stats) stats->n_outside++; return 0; }
	const unsigned char* cg = r + 32 + l_name;
	uint32_t clip = 0, mlen = 0, k = 0;
	int complex_cigar = n_cig == 0;
	/* [H] [S] M+ [S] [H]   (= and X count as M, as in parse_cigar, bamread.c:25) */
	while (k < n_cig && (rd32(cg + 4 * k) & 15) == 5) k++;
	if (k < n_cig && (rd32(cg + 4 * k) & 15) == 4) { clip = rd32(cg + 4 * k) >> 4; k++; }
	while (k < n_cig) {
		const uint32_t op = rd32(cg + 4 * k) & 15;
		if (op == 0 || op == 7 || op == 8) { mlen += rd32(cg + 4 * k) >> 4; k++; } else break;
	}
	if (k < n_cig && (rd32(cg + 4 * k) & 15) == 4) k++;
	while (k < n_cig && (rd32(cg + 4 * k) & 15) == 5) k++;
	if (k != n_cig || mlen == 0 || clip + mlen > l_seq || clip > 65535 || mlen > 65535) complex_cigar = 1;
	if (opt->tid >= 0 && !complex_cigar && (pos + (int32_t)mlen <= opt->start || pos >= opt->end)) { if (stats) stats->n_outside++; return 0; }
	if (complex_cigar) { if (stats) stats->n_complex++; return 0; }
	v->pos = pos;
	v->mpos = (int32_t)rd32(r + 24);
	v->tlen = (int32_t)rd32(r + 28);
	v->l_seq = l_seq;
	v->clip = clip;
	v->mlen = mlen;
	v->mapq = r[9];
	v->strand = (flag & 16) ? 1 : 0;
	v->seq = cg + 4 * n_cig;
	v->qual = v->seq + (l_seq + 1) / 2;
	return 1;
}

static void count_pool(pool_t* pl, const crisp_bam_options* opt)
{
	size_t o = pl->rec_start;
	rec_view v;
	while (o + 4 <= pl->data_bytes) {
		const uint32_t bs = rd32(pl->data + o);
		if (bs < 32 || o + 4 + bs > pl->data_bytes) { pl->error = -3; return; }
		if (view_record(pl->data + o + 4, bs, opt, &v, pl)) {
			pl->kept++;
			pl->bases += ((uint64_t)v.l_seq + 7) & ~(uint64_t)7;
		}
		if (pl->error) return;
		o += 4 + bs;
	}
	if (o != pl->data_bytes) pl->error = -3;
}

static void fill_pool(const job_t* jb, int p)
{
	const pool_t* pl = &jb->pools[p];
	size_t o = pl->rec_start;
	uint64_t r = jb->pool_read0[p], b = jb->pool_base0[p];
	rec_view v;
	while (o + 4 <= pl->data_bytes) {
		const uint32_t bs = rd32(pl->data + o);
		if (view_record(pl->data + o + 4, bs, jb->opt, &v, NULL)) {
			crispgpu_alnrec* a = &jb->rec[r];
			a->pos = v.pos;
			a->base_off = (uint32_t)b;
			a->clip = (uint16_t)v.clip;
			a->mlen = (uint16_t)v.mlen;
			a->mapq = v.mapq;
			a->strand = v.strand;
			a->reserved = 0;
			if (jb->mate_pos) { jb->mate_pos[r] = v.mpos; jb->isize[r] = v.tlen; }
			const uint64_t padded = ((uint64_t)v.l_seq + 7) & ~(uint64_t)7;
			memcpy(jb->qual + b, v.qual, v.l_seq);
			memset(jb->qual + b + v.l_seq, 0, padded - v.l_seq);
			memcpy(jb->seq4 + b / 2, v.seq, (v.l_seq + 1) / 2);
			memset(jb->seq4 + b / 2 + (v.l_seq + 1) / 2, 0, padded / 2 - (v.l_seq + 1) / 2);
			r++;
			b += padded;
		}
		o += 4 + bs;
	}
}

static void* worker(void* arg)
{
	job_t* jb = (job_t*)arg;
	for (;;) {
		if (jb->phase == 0) {   /* inflate: items in chunks of 16 blocks */
			pthread_mutex_lock(&jb->mu);
			const size_t k0 = jb->next_item;
			jb->next_item += 16;
			pthread_mutex_unlock(&jb->mu);
			if (k0 >= jb->n_items) return NULL;
			for (size_t k = k0; k < k0 + 16 && k < jb->n_items; k++) {
				const pool_t* pl = &jb->pools[jb->item_pool[k]];
				const int rc = inflate_block(pl, &pl->blocks[jb->item_block[k]]);
				if (rc) jb->error = rc;
			}
		} else {
			pthread_mutex_lock(&jb->mu);
			const int p = jb->next_pool++;
			pthread_mutex_unlock(&jb->mu);
			if (p >= jb->n_pools) return NULL;
			if (jb->phase == 1) {
				int rc = parse_header(&jb->pools[p]);
				if (rc) jb->pools[p].error = rc;
				else count_pool(&jb->pools[p], jb->opt);
			} else fill_pool(jb, p);
		}
	}
}

static void run_phase(job_t* jb, int phase, int threads)
{
	jb->phase = phase;
	jb->next_item = 0;
	jb->next_pool = 0;
	if (threads <= 1) { worker(jb); return; }
	pthread_t* th = (pthread_t*)malloc(sizeof(pthread_t) * (size_t)threads);
	int started = 0;
	if (th)
		for (int t = 0; t < threads; t++)
			if (pthread_create(&th[started], NULL, worker, jb) == 0) started++;
	if (!started) worker(jb);
	for (int t = 0; t < started; t++) pthread_join(th[t], NULL);
	free(th);
}

void crisp_bam_release(crisp_bam_reads* r, void (*release)(void*))
{
	if (!r) return;
	if (!release) release = free;
	if (r->aln.pool_off) release((void*)r->aln.pool_off);
	if (r->aln.rec) release((void*)r->aln.rec);
	if (r->aln.mate_pos) release((void*)r->aln.mate_pos);
	if (r->aln.isize) release((void*)r->aln.isize);
	if (r->aln.seq4) release((void*)r->aln.seq4);
	if (r->aln.qual) release((void*)r->aln.qual);
	memset(&r->aln, 0, sizeof r->aln);
}

int crisp_bam_decode(const char* const* paths, int32_t n_pools, const crisp_bam_options* opt, crisp_bam_reads* out)
{
	if (!paths || n_pools <= 0 || !opt || !out) return -1;
	crisp_bam_reads old;
	memset(&old, 0, sizeof old);
	if (opt->reuse) old = *out;   /* arrays of an earlier call */
	memset(out, 0, sizeof *out);
	void* (*alloc)(size_t) = opt->alloc ? opt->alloc : malloc;
	void (*release)(void*) = opt->release ? opt->release : free;
	const int threads = opt->threads > 0 ? opt->threads : 1;
	int rc = 0;
	job_t jb;
	memset(&jb, 0, sizeof jb);
	pthread_mutex_init(&jb.mu, NULL);
	jb.pools = (pool_t*)calloc((size_t)n_pools, sizeof(pool_t));
	jb.pool_read0 = (uint64_t*)calloc((size_t)n_pools + 1, sizeof(uint64_t));
	jb.pool_base0 = (uint64_t*)calloc((size_t)n_pools + 1, sizeof(uint64_t));
	jb.n_pools = n_pools;
	jb.opt = opt;
	if (!jb.pools || !jb.pool_read0 || !jb.pool_base0) { rc = -4; goto done; }
	double t0 = now_s();
	for (int p = 0; p < n_pools && !rc; p++) {
		rc = read_file(paths[p], &jb.pools[p]);
		if (!rc) rc = index_blocks(&jb.pools[p]);
		if (!rc) { jb.n_items += jb.pools[p].n_blocks; out->compressed_bytes += (int64_t)jb.pools[p].file_bytes; out->inflated_bytes += (int64_t)jb.pools[p].data_bytes; }
	}
	if (rc) goto done;
	out->read_s = now_s() - t0;
	jb.item_pool = (size_t*)malloc(sizeof(size_t) * (jb.n_items + 1));
	jb.item_block = (size_t*)malloc(sizeof(size_t) * (jb.n_items + 1));
	if (!jb.item_pool || !jb.item_block) { rc = -4; goto done; }
	{
		size_t k = 0;
		for (int p = 0; p < n_pools; p++)
			for (size_t b = 0; b < jb.pools[p].n_blocks; b++) { jb.item_pool[k] = (size_t)p; jb.item_block[k] = b; k++; }
	}
	t0 = now_s();
	run_phase(&jb, 0, threads);
	out->inflate_s = now_s() - t0;
	if (jb.error) { rc = jb.error; goto done; }
	t0 = now_s();
	run_phase(&jb, 1, threads);
	for (int p = 0; p < n_pools; p++) {
		const pool_t* pl = &jb.pools[p];
		if (pl->error) { rc = pl->error; goto done; }
		jb.pool_read0[p + 1] = jb.pool_read0[p] + pl->kept;
		jb.pool_base0[p + 1] = jb.pool_base0[p] + pl->bases;
		out->n_records += pl->n_records; out->n_flag_filtered += pl->n_flag; out->n_outside += pl->n_outside; out->n_complex += pl->n_complex;
	}
	{
		const uint64_t R = jb.pool_read0[n_pools], B = jb.pool_base0[n_pools];
		if (R >= (1ull << 32) || B >= (1ull << 32)) { rc = -1; goto done; }   /* the window is too large for 32-bit offsets: decode it in slices */
		uint32_t* pool_off;
		if (old.aln.rec && old.cap_reads >= R && old.cap_bases >= B && old.cap_pools >= n_pools && (!opt->paired || old.cap_paired)) {
			pool_off = (uint32_t*)old.aln.pool_off; jb.rec = (crispgpu_alnrec*)old.aln.rec; jb.seq4 = (uint8_t*)old.aln.seq4; jb.qual = (uint8_t*)old.aln.qual;
			if (opt->paired) { jb.mate_pos = (int32_t*)old.aln.mate_pos; jb.isize = (int32_t*)old.aln.isize; }
			else { if (old.aln.mate_pos) release((void*)old.aln.mate_pos); if (old.aln.isize) release((void*)old.aln.isize); }
			out->cap_reads = old.cap_reads; out->cap_bases = old.cap_bases; out->cap_pools = old.cap_pools; out->cap_paired = opt->paired ? old.cap_paired : 0;
			memset(&old, 0, sizeof old);
		} else {
			const uint64_t cr = R + R / 8 + 1, cb = (B + B / 8 + 64) & ~(uint64_t)7;   /* some slack for the next window */
			pool_off = (uint32_t*)alloc(sizeof(uint32_t) * ((size_t)n_pools + 1));
			jb.rec = (crispgpu_alnrec*)alloc(sizeof(crispgpu_alnrec) * cr);
			jb.seq4 = (uint8_t*)alloc(cb / 2 + 16);
			jb.qual = (uint8_t*)alloc(cb + 16);
			if (opt->paired) {
				jb.mate_pos = (int32_t*)alloc(sizeof(int32_t) * cr);
				jb.isize = (int32_t*)alloc(sizeof(int32_t) * cr);
			}
			out->cap_reads = cr; out->cap_bases = cb; out->cap_pools = n_pools; out->cap_paired = opt->paired;
		}
		out->aln.n_pools = n_pools;
		out->aln.pool_off = pool_off;
		out->aln.rec = jb.rec;
		out->aln.mate_pos = jb.mate_pos;
		out->aln.isize = jb.isize;
		out->aln.seq4 = jb.seq4;
		out->aln.qual = jb.qual;
		out->aln.n_bases = B;
		if (!pool_off || !jb.rec || !jb.seq4 || !jb.qual || (opt->paired && (!jb.mate_pos || !jb.isize))) { rc = -4; goto done; }
		for (int p = 0; p <= n_pools; p++) pool_off[p] = (uint32_t)jb.pool_read0[p];
	}
	run_phase(&jb, 2, threads);
	out->parse_s = now_s() - t0;
	out->n_refs = jb.pools[0].n_refs;
	out->first_ref_len = jb.pools[0].first_ref_len;
	memcpy(out->first_ref_name, jb.pools[0].first_ref_name, sizeof out->first_ref_name);
done:
	if (jb.pools)
		for (int p = 0; p < n_pools; p++) { free(jb.pools[p].file); free(jb.pools[p].blocks); free(jb.pools[p].data); }
	free(jb.pools);
	free(jb.pool_read0);
	free(jb.pool_base0);
	free(jb.item_pool);
	free(jb.item_block);
	pthread_mutex_destroy(&jb.mu);
	crisp_bam_release(&old, release);   /* arrays of the earlier call that were not taken over */
	if (rc) crisp_bam_release(out, release);
	return rc;
}
