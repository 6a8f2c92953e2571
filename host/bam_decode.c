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

#include <limits.h>
#include <pthread.h>
#include <stdio.h>
#include <stdlib.h>
#include <string.h>
#include <time.h>
#include <zlib.h>

typedef struct {
	size_t src, dst;
	size_t off;             /* file offset of the gzip member (the "coffset" of a BAI virtual offset) */
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
	/* region mode (a .bai next to the file and a region in the options): only the header blocks and the blocks the index
	 * names for the region are inflated; the record walk starts at the index's first record and may end inside a record */
	int region_mode;
	size_t n_seg;              /* region mode: runs of records in `data`, each starting at a record (whole file: one run, header end to data end) */
	size_t* seg_beg;
	size_t* seg_end;
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
		b->off = o;
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
	return 0;
}

static int alloc_data(pool_t* pl)
{
	pl->data = (unsigned char*)malloc(pl->data_bytes + 16);
	return pl->data ? 0 : -4;
}

static int inflate_to(const pool_t* pl, const block_t* b, unsigned char* dst)
{
	if (b->isize == 0) return 0;
	z_stream zs;
	memset(&zs, 0, sizeof zs);
	if (inflateInit2(&zs, -15) != Z_OK) return -4;
	zs.next_in = pl->file + b->src;
	zs.avail_in = b->csize;
	zs.next_out = dst;
	zs.avail_out = b->isize;
	const int rc = inflate(&zs, Z_FINISH);
	const int ok = rc == Z_STREAM_END && zs.avail_out == 0;
	inflateEnd(&zs);
	if (!ok) return -3;
	if ((uint32_t)crc32(crc32(0L, Z_NULL, 0), dst, b->isize) != rd32(pl->file + b->src + b->csize)) return -3;
	return 0;
}

static int inflate_block(const pool_t* pl, const block_t* b) { return inflate_to(pl, b, pl->data + b->dst); }

/* ---- region seek through the BAI index (SAM/BAM specification section 5.2; the reference's --regions goes through the
 * same index with bam_fetch, samtools/bam_index.c).  For the region of the options the index gives the virtual offsets
 * (file offset of a BGZF block << 16 | offset inside its inflated bytes) between which every overlapping alignment lies:
 * from the first chunk of the bins that can hold such an alignment (not before the linear index's entry for the 16 kb
 * window of `start`) to the last chunk's end.  Only the header's blocks and the blocks in that range are inflated. ---- */
static uint64_t rd64(const unsigned char* p) { return (uint64_t)rd32(p) | ((uint64_t)rd32(p + 4) << 32); }

static size_t block_at(const pool_t* pl, uint64_t coffset)
{
	size_t lo = 0, hi = pl->n_blocks;   /* first block with off >= coffset */
	while (lo < hi) { const size_t mid = (lo + hi) / 2; if (pl->blocks[mid].off < coffset) lo = mid + 1; else hi = mid; }
	return lo;
}

static int select_region(pool_t* pl, const char* path, const crisp_bam_options* opt)
{
	if (opt->tid < 0 || (opt->start <= 0 && opt->end == INT32_MAX) || pl->n_blocks == 0) return 0;
	const size_t pn = strlen(path);
	char* ip = (char*)malloc(pn + 8);
	if (!ip) return -4;
	memcpy(ip, path, pn);
	memcpy(ip + pn, ".bai", 5);
	FILE* f = fopen(ip, "rb");
	if (!f && pn > 4 && memcmp(path + pn - 4, ".bam", 4) == 0) { memcpy(ip + pn - 4, ".bai", 5); f = fopen(ip, "rb"); }   /* x.bam.bai or x.bai */
	free(ip);
	if (!f) return 0;   /* no index: the whole file is decoded and filtered */
	unsigned char* ix = NULL;
	size_t in = 0;
	{
		fseek(f, 0, SEEK_END);
		const long n = ftell(f);
		fseek(f, 0, SEEK_SET);
		if (n > 0) { ix = (unsigned char*)malloc((size_t)n); if (ix && fread(ix, 1, (size_t)n, f) == (size_t)n) in = (size_t)n; }
		fclose(f);
	}
	int rc = 0;
	if (in < 8 || memcmp(ix, "BAI\1", 4) != 0) { free(ix); return 0; }   /* not an index: as if there were none */
	const int32_t n_ref = (int32_t)rd32(ix + 4);
	const int64_t beg = opt->start > 0 ? opt->start : 0, end = opt->end > beg ? opt->end : beg + 1;
	uint64_t* ch_beg = NULL;
	uint64_t* ch_end = NULL;
	size_t n_ch = 0, ch_cap = 0;
	size_t o = 8;
	for (int32_t r = 0; r < n_ref && r <= opt->tid; r++) {
		if (o + 4 > in) goto bad;
		const int32_t n_bin = (int32_t)rd32(ix + o);
		o += 4;
		const size_t bins_at = o;
		for (int32_t b = 0; b < n_bin; b++) {   /* skip to the linear index */
			if (o + 8 > in) goto bad;
			const int32_t n_chunk = (int32_t)rd32(ix + o + 4);
			o += 8 + (size_t)n_chunk * 16;
			if (o > in) goto bad;
		}
		if (o + 4 > in) goto bad;
		const int32_t n_intv = (int32_t)rd32(ix + o);
		const size_t intv_at = o + 4;
		o = intv_at + (size_t)n_intv * 8;
		if (o > in) goto bad;
		if (r != opt->tid) continue;
		/* linear index: no overlapping alignment starts before the first alignment of the window of `beg` */
		uint64_t min_off = 0;
		if (n_intv > 0) {
			int64_t w = beg >> 14;
			if (w >= n_intv) w = n_intv - 1;
			min_off = rd64(ix + intv_at + (size_t)w * 8);
			for (int64_t k = w - 1; min_off == 0 && k >= 0; k--) min_off = rd64(ix + intv_at + (size_t)k * 8);
		}
		/* bins that can hold an alignment overlapping [beg, end) (reg2bins) */
		size_t q = bins_at;
		for (int32_t b = 0; b < n_bin; b++) {
			const uint32_t bin = rd32(ix + q);
			const int32_t n_chunk = (int32_t)rd32(ix + q + 4);
			const unsigned char* ch = ix + q + 8;
			q += 8 + (size_t)n_chunk * 16;
			if (bin == 37450) continue;   /* samtools' pseudo-bin: counts, not chunks */
			int hit = bin == 0;
			static const int first[5] = {1, 9, 73, 585, 4681}, shift[5] = {26, 23, 20, 17, 14};
			for (int l = 0; l < 5 && !hit; l++)
				if ((int64_t)bin >= first[l] + (beg >> shift[l]) && (int64_t)bin <= first[l] + ((end - 1) >> shift[l]) && (l == 4 || (int64_t)bin < first[l + 1])) hit = 1;
			if (!hit) continue;
			for (int32_t c = 0; c < n_chunk; c++) {
				const uint64_t cb = rd64(ch + (size_t)c * 16), ce = rd64(ch + (size_t)c * 16 + 8);
				if (ce <= min_off || ce <= cb) continue;
				if (n_ch == ch_cap) {
					ch_cap = ch_cap ? ch_cap * 2 : 64;
					uint64_t* nb2 = (uint64_t*)realloc(ch_beg, ch_cap * 8);
					uint64_t* ne2 = nb2 ? (uint64_t*)realloc(ch_end, ch_cap * 8) : NULL;
					if (nb2) ch_beg = nb2;
					if (ne2) ch_end = ne2;
					if (!nb2 || !ne2) { rc = -4; goto out; }
				}
				ch_beg[n_ch] = cb < min_off ? min_off : cb;   /* both are starts of records */
				ch_end[n_ch] = ce;
				n_ch++;
			}
		}
	}
	if (opt->tid >= n_ref) n_ch = 0;
	{
		/* the header's blocks: inflate from the start until the header is complete */
		size_t hb = 0, have = 0, need = 12, cap = 1 << 17;
		unsigned char* hd = (unsigned char*)malloc(cap);
		if (!hd) { rc = -4; goto out; }
		int stage = 0;          /* 0: magic + l_text, 1: text + n_ref, 2: reference entries */
		uint32_t refs_left = 0;
		size_t pos = 0;
		for (;;) {
			while (have < need) {
				if (hb >= pl->n_blocks) { free(hd); rc = -3; goto out; }
				const block_t* b = &pl->blocks[hb];
				if (have + b->isize > cap) {
					while (have + b->isize > cap) cap *= 2;
					unsigned char* h2 = (unsigned char*)realloc(hd, cap);
					if (!h2) { free(hd); rc = -4; goto out; }
					hd = h2;
				}
				const int r2 = inflate_to(pl, b, hd + have);
				if (r2) { free(hd); rc = r2; goto out; }
				have += b->isize;
				hb++;
			}
			if (stage == 0) {
				if (memcmp(hd, "BAM\1", 4) != 0) { free(hd); rc = -3; goto out; }
				pos = 8 + (size_t)rd32(hd + 4);
				need = pos + 4;
				stage = 1;
			} else if (stage == 1) {
				refs_left = rd32(hd + pos);
				pos += 4;
				need = pos + (refs_left ? 4 : 0);
				stage = 2;
				if (!refs_left) break;
			} else if (need == pos + 4) {
				need = pos + 4 + (size_t)rd32(hd + pos) + 4;   /* l_name, name, l_ref */
			} else {
				pos = need;
				if (--refs_left == 0) break;
				need = pos + 4;
			}
		}
		free(hd);
		/* hb blocks hold the header (and possibly the first records).  The chunks collected above become runs of blocks:
		 * sorted, clipped to the linear index's lower bound, merged where their blocks overlap or touch (a record between
		 * two chunks is then simply walked over, and none is seen twice). */
		pl->region_mode = 1;
		size_t nrun = 0;
		for (size_t a = 0; a < n_ch; a++)            /* insertion sort: a region has a few dozen chunks */
			for (size_t b2 = a; b2 > 0 && ch_beg[b2] < ch_beg[b2 - 1]; b2--) {
				uint64_t t = ch_beg[b2]; ch_beg[b2] = ch_beg[b2 - 1]; ch_beg[b2 - 1] = t;
				t = ch_end[b2]; ch_end[b2] = ch_end[b2 - 1]; ch_end[b2 - 1] = t;
			}
		size_t* run_b0 = (size_t*)malloc(sizeof(size_t) * (n_ch + 1));
		size_t* run_b1 = (size_t*)malloc(sizeof(size_t) * (n_ch + 1));
		size_t* run_in = (size_t*)malloc(sizeof(size_t) * (n_ch + 1));   /* offset of the run's first record inside its first block */
		pl->seg_beg = (size_t*)malloc(sizeof(size_t) * (n_ch + 1));
		pl->seg_end = (size_t*)malloc(sizeof(size_t) * (n_ch + 1));
		if (!run_b0 || !run_b1 || !run_in || !pl->seg_beg || !pl->seg_end) { free(run_b0); free(run_b1); free(run_in); rc = -4; goto out; }
		int fits = 1;
		for (size_t a = 0; a < n_ch && fits; a++) {
			size_t b0 = block_at(pl, ch_beg[a] >> 16), b1 = block_at(pl, ch_end[a] >> 16);
			if (b0 >= pl->n_blocks || pl->blocks[b0].off != (ch_beg[a] >> 16) || (ch_beg[a] & 0xffff) >= pl->blocks[b0].isize) { fits = 0; break; }   /* index of another file */
			if (b1 < pl->n_blocks && pl->blocks[b1].off == (ch_end[a] >> 16) && (ch_end[a] & 0xffff) > 0) b1++;   /* the block the chunk ends in */
			if (b1 <= b0) b1 = b0 + 1;
			size_t in0 = (size_t)(ch_beg[a] & 0xffff);
			if (b0 < hb) { b0 = hb; in0 = SIZE_MAX; if (b1 < hb) b1 = hb; }   /* starts inside the header's blocks: walked from the header's end */
			if (nrun && b0 <= run_b1[nrun - 1]) { if (b1 > run_b1[nrun - 1]) run_b1[nrun - 1] = b1; }
			else { run_b0[nrun] = b0; run_b1[nrun] = b1; run_in[nrun] = in0; nrun++; }
		}
		if (!fits) { free(run_b0); free(run_b1); free(run_in); free(pl->seg_beg); free(pl->seg_end); pl->seg_beg = pl->seg_end = NULL; pl->region_mode = 0; goto bad; }
		/* keep blocks [0, hb) and the runs; a run that begins in the header's blocks starts where parse_header ends */
		{
			block_t* nb = (block_t*)malloc(sizeof(block_t) * (pl->n_blocks + 1));
			if (!nb) { free(run_b0); free(run_b1); free(run_in); rc = -4; goto out; }
			size_t n = 0, dst = 0;
			for (size_t k = 0; k < hb; k++) { nb[n] = pl->blocks[k]; nb[n].dst = dst; dst += nb[n].isize; n++; }
			pl->n_seg = 0;
			for (size_t r2 = 0; r2 < nrun; r2++) {
				const size_t at = dst;
				if (run_in[r2] == SIZE_MAX && run_b0[r2] == hb) {
					/* continues the header's blocks: the segment starts at the header's end (filled in by parse_header) */
					pl->seg_beg[pl->n_seg] = SIZE_MAX;
				} else pl->seg_beg[pl->n_seg] = at + run_in[r2];
				for (size_t k = run_b0[r2]; k < run_b1[r2]; k++) { nb[n] = pl->blocks[k]; nb[n].dst = dst; dst += nb[n].isize; n++; }
				pl->seg_end[pl->n_seg] = dst;
				pl->n_seg++;
			}
			free(pl->blocks);
			pl->blocks = nb;
			pl->n_blocks = n;
			pl->data_bytes = dst;
		}
		free(run_b0); free(run_b1); free(run_in);
	}
out:
	free(ix);
	free(ch_beg);
	free(ch_end);
	return rc;
bad:
	free(ix);
	free(ch_beg);
	free(ch_end);
	return 0;   /* an index that does not fit the file is ignored: whole-file decode */
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
	if (!pl->region_mode) {
		pl->seg_beg = (size_t*)malloc(sizeof(size_t));
		pl->seg_end = (size_t*)malloc(sizeof(size_t));
		if (!pl->seg_beg || !pl->seg_end) return -4;
		pl->n_seg = 1;
		pl->seg_beg[0] = o;
		pl->seg_end[0] = n;
	} else
		for (size_t k = 0; k < pl->n_seg; k++)
			if (pl->seg_beg[k] == SIZE_MAX) pl->seg_beg[k] = o;   /* the run that continues the header's blocks */
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

/* In region mode the walk ends at the first alignment that starts at or after the region's end (the files are sorted by
 * coordinate), and a run of blocks may end inside a record that belongs to no chunk of the region. */
static int past_region(const unsigned char* r, const crisp_bam_options* opt)
{
	const int32_t ref_id = (int32_t)rd32(r), pos = (int32_t)rd32(r + 4);
	return ref_id > opt->tid || (ref_id == opt->tid && pos >= opt->end);
}

static void count_pool(pool_t* pl, const crisp_bam_options* opt)
{
	rec_view v;
	for (size_t g = 0; g < pl->n_seg; g++) {
		size_t o = pl->seg_beg[g];
		const size_t end = pl->seg_end[g];
		while (o + 4 <= end) {
			const uint32_t bs = rd32(pl->data + o);
			if (pl->region_mode && bs >= 32 && o + 4 + bs > end) break;
			if (bs < 32 || o + 4 + bs > end) { pl->error = -3; return; }
			if (pl->region_mode && past_region(pl->data + o + 4, opt)) return;
			if (view_record(pl->data + o + 4, bs, opt, &v, pl)) {
				pl->kept++;
				pl->bases += ((uint64_t)v.l_seq + 7) & ~(uint64_t)7;
			}
			if (pl->error) return;
			o += 4 + bs;
		}
		if (o != end && !pl->region_mode) pl->error = -3;
	}
}

static void fill_pool(const job_t* jb, int p)
{
	const pool_t* pl = &jb->pools[p];
	uint64_t r = jb->pool_read0[p], b = jb->pool_base0[p];
	rec_view v;
	for (size_t g = 0; g < pl->n_seg; g++) {
		size_t o = pl->seg_beg[g];
		const size_t end = pl->seg_end[g];
		while (o + 4 <= end) {
			const uint32_t bs = rd32(pl->data + o);
			if (o + 4 + bs > end) break;   /* region mode only (count_pool has refused it otherwise) */
			if (pl->region_mode && past_region(pl->data + o + 4, jb->opt)) return;
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
		if (!rc && opt->use_index) rc = select_region(&jb.pools[p], paths[p], opt);
		if (!rc) rc = alloc_data(&jb.pools[p]);
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
		for (int p = 0; p < n_pools; p++) { free(jb.pools[p].file); free(jb.pools[p].blocks); free(jb.pools[p].data); free(jb.pools[p].seg_beg); free(jb.pools[p].seg_end); }
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
