/*
 * crisp_shim.c — see crisp_shim.h.  Walks the reference's read lists the way the engine's three read-queue consumers
 * do (crisp/crispEM.c:104-136, parsebam/allelecounts.c:50-67, parsebam/process_indel_variant.c:25-62) and snapshots
 * what they would have read into the flat batch.
 */
#include <stdlib.h>
#include <string.h>
#include <math.h>
#include <ctype.h>

#include "bamread.h"
#include "bamsreader.h"
#include "variant.h"
#define CRISP_SHIM_WITH_REFERENCE_TYPES
#include "crisp_shim.h"

extern int maxalleles, MINQ, QVoffset, USE_QV;
extern unsigned int BTI[];
int calculate_indel_ambiguity(struct VARIANT* variant, REFLIST* reflist, int current, int allele);

/* grow an array geometrically; a failed allocation keeps the old block and makes the enclosing function return -1 */
#define GROW(ptr, cap, need, type)                                            \
	do {                                                                      \
		if ((size_t)(need) > (size_t)(cap)) {                                 \
			const size_t newcap_ = (size_t)(need) * 2 + 64;                   \
			type* np_ = (type*)realloc((ptr), newcap_ * sizeof(type));        \
			if (!np_) return -1;                                              \
			(ptr) = np_;                                                      \
			(cap) = newcap_;                                                  \
		}                                                                     \
	} while (0)

crisp_shim_batch* crisp_shim_batch_new(int n_pools, int want_qhigh, int want_stats)
{
	crisp_shim_batch* sb = (crisp_shim_batch*)calloc(1, sizeof(*sb));
	if (!sb) return NULL;
	sb->n_pools = n_pools;
	sb->want_qhigh = want_qhigh;
	sb->want_stats = want_stats;
	sb->read_off = (uint32_t*)calloc(1, sizeof(uint32_t));
	sb->alt_off = (uint32_t*)calloc(1, sizeof(uint32_t));
	sb->bin_off = (uint32_t*)calloc(1, sizeof(uint32_t));
	sb->ic_off = (uint32_t*)calloc(1, sizeof(uint32_t));
	sb->bin_count = (uint32_t*)calloc(4096, sizeof(uint32_t));
	sb->bin_touched = (uint16_t*)calloc(4096, sizeof(uint16_t));
	if (!sb->read_off || !sb->alt_off || !sb->bin_off || !sb->ic_off || !sb->bin_count || !sb->bin_touched) {
		crisp_shim_batch_free(sb);
		return NULL;
	}
	return sb;
}

void crisp_shim_batch_set_compact(crisp_shim_batch* sb, int compact) { sb->compact = compact != 0; }

void crisp_shim_batch_clear(crisp_shim_batch* sb)
{
	sb->n_sites = 0;
	sb->n_reads = sb->n_alt = 0;
	sb->n_amb = 0;
	sb->read_off[0] = 0;
	sb->alt_off[0] = 0;
	sb->n_bins = 0;
	sb->bin_off[0] = 0;
	sb->n_ic = 0;
	sb->ic_off[0] = 0;
}

void crisp_shim_batch_free(crisp_shim_batch* sb)
{
	if (!sb) return;
	free(sb->tid); free(sb->pos); free(sb->maxvec_ct); free(sb->counts); free(sb->indcounts); free(sb->amb_idx); free(sb->amb_dec);
	free(sb->refbase); free(sb->maxvec_al); free(sb->indel_len); free(sb->hplen); free(sb->read_off); free(sb->alt_off);
	free(sb->reads); free(sb->alt_reads); free(sb->amb_ep); free(sb->qhigh); free(sb->stats); free(sb->tstats);
	free(sb->bin_off); free(sb->bins); free(sb->bin_count); free(sb->bin_touched); free(sb->ic_off); free(sb->ic_sparse);
	free(sb);
}

void crisp_shim_batch_view(const crisp_shim_batch* sb, crispgpu_batch* b)
{
	memset(b, 0, sizeof(*b));
	b->n_sites = sb->n_sites;
	b->tid = sb->tid; b->pos = sb->pos; b->refbase = sb->refbase; b->maxvec_al = sb->maxvec_al; b->maxvec_ct = sb->maxvec_ct;
	b->counts = sb->counts; b->indcounts = sb->indcounts; b->indel_len = sb->indel_len; b->hplen = sb->hplen;
	b->read_off = sb->read_off; b->reads = sb->compact ? NULL : sb->reads;
	if (sb->compact) { b->bin_off = sb->bin_off; b->bins = sb->bins; b->ic_off = sb->ic_off; b->ic_sparse = sb->ic_sparse; b->indcounts = NULL; }
	b->alt_off = sb->alt_off; b->alt_reads = sb->alt_reads;
	b->amb_idx = sb->amb_idx; b->amb_dec = sb->amb_dec; b->amb_ep = sb->amb_ep; b->n_amb = sb->n_amb;
	b->qhigh = sb->want_qhigh ? sb->qhigh : NULL;
	b->stats = sb->want_stats ? sb->stats : NULL;
	b->tstats = sb->want_stats ? sb->tstats : NULL;
}

/* the arrays of a view in declaration order: address of the pointer field and its size in bytes */
typedef struct { const void** field; size_t bytes; } shim_span;

static int shim_spans(const crisp_shim_batch* sb, crispgpu_batch* v, shim_span* sp)
{
	const size_t n = (size_t)sb->n_sites, P = (size_t)sb->n_pools;
	int k = 0;
#define SPAN(name, nbytes) do { sp[k].field = (const void**)&v->name; sp[k].bytes = v->name ? (size_t)(nbytes) : 0; k++; } while (0)
	SPAN(tid, n * 4); SPAN(pos, n * 4); SPAN(refbase, n); SPAN(maxvec_al, n * 8); SPAN(maxvec_ct, n * 32); SPAN(counts, n * 96);
	SPAN(indcounts, n * P * 96); SPAN(indel_len, n * 16); SPAN(hplen, n * 16);
	SPAN(read_off, (n * P + 1) * 4); SPAN(reads, sb->n_reads * sizeof(crispgpu_read));
	SPAN(alt_off, (n * 5 + 1) * 4); SPAN(alt_reads, sb->n_alt * sizeof(crispgpu_altread));
	SPAN(amb_idx, n * 20); SPAN(amb_dec, (size_t)sb->n_amb * P * 16); SPAN(amb_ep, (size_t)sb->n_amb * P * 16);
	SPAN(qhigh, n * P * 128); SPAN(stats, n * P * 128); SPAN(tstats, n * 128);
	SPAN(bin_off, (n * P + 1) * 4); SPAN(bins, sb->n_bins * sizeof(crispgpu_bin));
	SPAN(ic_off, (n * P + 1) * 4); SPAN(ic_sparse, sb->n_ic * sizeof(crispgpu_count));
#undef SPAN
	return k;
}

size_t crisp_shim_batch_packed_bytes(const crisp_shim_batch* sb)
{
	crispgpu_batch v;
	shim_span sp[32];
	crisp_shim_batch_view(sb, &v);
	const int k = shim_spans(sb, &v, sp);
	size_t total = 0;
	for (int i = 0; i < k; i++) total += (sp[i].bytes + 255) & ~(size_t)255;
	return total;
}

int crisp_shim_batch_pack(const crisp_shim_batch* sb, void* arena, size_t arena_cap, crispgpu_batch* out)
{
	shim_span sp[32];
	crisp_shim_batch_view(sb, out);
	const int k = shim_spans(sb, out, sp);
	size_t off = 0;
	for (int i = 0; i < k; i++) {
		if (sp[i].bytes == 0) { *sp[i].field = NULL; continue; }  /* absent or empty */
		const size_t need = (sp[i].bytes + 255) & ~(size_t)255;
		if (off + need > arena_cap) return -1;
		memcpy((char*)arena + off, *sp[i].field, sp[i].bytes);
		*sp[i].field = (char*)arena + off;
		off += need;
	}
	out->arena_base = arena;
	out->arena_bytes = off;
	return 0;
}

/* room for one more site in every per-site array; -1 when an allocation fails (the batch stays valid) */
static int reserve_site(crisp_shim_batch* sb)
{
	const size_t n = (size_t)sb->n_sites + 1, P = sb->n_pools;
	if ((int)n <= sb->cap_sites) return 0;
	const size_t c = n * 2 + 64;
#define RESIZE(field, bytes) do { void* np_ = realloc(sb->field, (bytes)); if (!np_) return -1; sb->field = np_; } while (0)
	RESIZE(tid, c * 4); RESIZE(pos, c * 4); RESIZE(refbase, c); RESIZE(maxvec_al, c * 8); RESIZE(maxvec_ct, c * 32);
	RESIZE(counts, c * 96); RESIZE(indcounts, c * P * 96); RESIZE(indel_len, c * 16); RESIZE(hplen, c * 16);
	RESIZE(bin_off, (c * P + 1) * 4); RESIZE(ic_off, (c * P + 1) * 4); RESIZE(read_off, (c * P + 1) * 4);
	RESIZE(alt_off, (c * 5 + 1) * 4); RESIZE(amb_idx, c * 20);
	if (sb->want_qhigh) RESIZE(qhigh, c * P * 128);
	if (sb->want_stats) { RESIZE(stats, c * P * 128); RESIZE(tstats, c * 128); }
#undef RESIZE
	sb->cap_sites = (int)c;
	return 0;
}

static int append_site(crisp_shim_batch* sb, REFLIST* reflist, int current, READQUEUE* bq, struct BAMFILE_data* bd,
                       struct VARIANT* variant, ALLELE* maxvec, int tid);

int crisp_shim_append_site(crisp_shim_batch* sb, REFLIST* reflist, int current, READQUEUE* bq, struct BAMFILE_data* bd,
                           struct VARIANT* variant, ALLELE* maxvec, int tid)
{
	const size_t n_reads = sb->n_reads, n_alt = sb->n_alt, n_bins = sb->n_bins, n_ic = sb->n_ic;
	const int n_amb = sb->n_amb;
	const int rc = append_site(sb, reflist, current, bq, bd, variant, maxvec, tid);
	if (rc != 0) {  /* drop whatever the failed site had already appended */
		sb->n_reads = n_reads; sb->n_alt = n_alt; sb->n_bins = n_bins; sb->n_ic = n_ic; sb->n_amb = n_amb;
		memset(sb->bin_count, 0, 4096 * sizeof(uint32_t));
	}
	return rc;
}

static int append_site(crisp_shim_batch* sb, REFLIST* reflist, int current, READQUEUE* bq, struct BAMFILE_data* bd,
                       struct VARIANT* variant, ALLELE* maxvec, int tid)
{
	const int P = sb->n_pools, s = sb->n_sites;
	/* the offset arrays are 32-bit: refuse a site that could push a total past them (flush the batch first) */
	if (sb->n_reads > 0xf0000000u || sb->n_alt > 0xf0000000u || sb->n_bins > 0xf0000000u || sb->n_ic > 0xf0000000u) return -2;
	if (reserve_site(sb) != 0) return -1;
	sb->tid[s] = tid;
	sb->pos[s] = variant->position;
	sb->refbase[s] = (uint8_t)variant->refbase;
	for (int k = 0; k < 8; k++) { sb->maxvec_al[s * 8 + k] = (uint8_t)maxvec[k].al; sb->maxvec_ct[s * 8 + k] = maxvec[k].ct; }
	memcpy(sb->counts + (size_t)s * 24, variant->counts, 24 * sizeof(int));
	for (int i = 0; i < P; i++) memcpy(sb->indcounts + ((size_t)s * P + i) * 24, variant->indcounts[i], 24 * sizeof(int));
	if (sb->compact)
		for (int i = 0; i < P; i++) {
			for (int q = 0; q < 24; q++) {
				if (variant->indcounts[i][q] == 0) continue;
				GROW(sb->ic_sparse, sb->cap_ic, sb->n_ic + 1, crispgpu_count);
				sb->ic_sparse[sb->n_ic++] = CRISPGPU_COUNT(q, variant->indcounts[i][q]);
			}
			sb->ic_off[(size_t)s * P + i + 1] = (uint32_t)sb->n_ic;
		}
	for (int i = 0; i < P; i++)
		for (int q = 0; q < 16; q++) {
			if (sb->want_qhigh) sb->qhigh[((size_t)s * P + i) * 16 + q] = variant->Qhighcounts[i][q];
			if (sb->want_stats) sb->stats[((size_t)s * P + i) * 16 + q] = variant->stats[i][q];
		}
	if (sb->want_stats) for (int q = 0; q < 16; q++) sb->tstats[(size_t)s * 16 + q] = variant->tstats[q];
	/* homopolymer ambiguity of the tested indel alleles, newcrispcaller.c:191-198 */
	for (int al = 1; al <= 5; al++) {
		if (maxvec[al].ct < 3) continue;
		const int a = maxvec[al].al;
		if (current >= 0) {
			variant->HPlength[a] = 0;
			if (a >= 4 && (variant->itb[a][0] == '+' || variant->itb[a][0] == '-')) calculate_indel_ambiguity(variant, reflist, current, a);
		}
	}
	for (int a = 0; a < 8; a++) {
		int len = 0;
		if (a >= 4 && (variant->itb[a][0] == '+' || variant->itb[a][0] == '-')) len = (int)strlen(variant->itb[a]) - 1;
		if (len > 32767 || variant->HPlength[a] > 32767 || variant->HPlength[a] < -32768) return -2;
		sb->indel_len[s * 8 + a] = (int16_t)(a >= 4 && variant->itb[a][0] == '-' ? -len : len);
		sb->hplen[s * 8 + a] = (int16_t)variant->HPlength[a];
	}
	/* 1. likelihood stream: what calculate_pooled_likelihoods_snp reads (crispEM.c:104-110) */
	int ntouched = 0;
	for (int i = 0; i < P; i++) {
		for (struct alignedread* r = bd[i].first; r != NULL && r->position <= variant->position; r = r->nextread) {
			if (r->lastpos < variant->position) continue;
			if (r->filter != '0' || r->type != 0) continue;
			const int o = r->l1 + r->delta;
			const crispgpu_read w = CRISPGPU_READ((unsigned char)r->quality[o], BTI[(unsigned char)r->sequence[o]], r->strand, r->bidir == 1);
			if (sb->compact) {
				if (sb->bin_count[w & 4095]++ == 0) sb->bin_touched[ntouched++] = (uint16_t)(w & 4095);
				sb->n_reads++;
				continue;
			}
			GROW(sb->reads, sb->cap_reads, sb->n_reads + 1, crispgpu_read);
			sb->reads[sb->n_reads++] = w;
		}
		sb->read_off[(size_t)s * P + i + 1] = (uint32_t)sb->n_reads;
		if (sb->compact) {
			/* emit the cell's bins in ascending value order; a multiplicity above 65535 takes several bins */
			for (int x = 1; x < ntouched; x++) {
				const uint16_t t = sb->bin_touched[x];
				int y = x - 1;
				for (; y >= 0 && sb->bin_touched[y] > t; y--) sb->bin_touched[y + 1] = sb->bin_touched[y];
				sb->bin_touched[y + 1] = t;
			}
			for (int x = 0; x < ntouched; x++) {
				const uint16_t v = sb->bin_touched[x];
				uint32_t left = sb->bin_count[v];
				sb->bin_count[v] = 0;
				while (left > 0) {
					const uint32_t take = left > 65535 ? 65535 : left;
					GROW(sb->bins, sb->cap_bins, sb->n_bins + 1, crispgpu_bin);
					sb->bins[sb->n_bins++] = CRISPGPU_BIN(v, take);
					left -= take;
				}
			}
			ntouched = 0;
			sb->bin_off[(size_t)s * P + i + 1] = (uint32_t)sb->n_bins;
		}
	}
	/* 2. alt-read side list: what calculate_uniquereads collects (allelecounts.c:47-67), per tested slot and pool */
	for (int k = 0; k < 5; k++) {
		const int a2 = maxvec[k + 1].al;
		if (maxvec[k + 1].ct >= 3) {
			for (int i = 0; i < P; i++) {
				const int cap = variant->indcounts[i][a2] + variant->indcounts[i][a2 + maxalleles] + variant->indcounts[i][a2 + 2 * maxalleles];
				int taken = 0;
				for (struct alignedread* r = bd[i].first; r != NULL && r->position <= variant->position; r = r->nextread) {
					if (r->lastpos <= variant->position || r->filter == '1') continue;
					int hit = 0;
					if (r->quality[r->l1 + r->delta] >= MINQ + QVoffset && r->type == 0 && a2 < 4 && taken < cap) {
						hit = toupper(r->sequence[r->l1 + r->delta]) == variant->itb[a2][0];
					} else if (r->type > 0 && a2 >= 4 && taken < cap) {
						hit = variant->itb[a2][0] == '+' && (int)strlen(variant->itb[a2]) - 1 == r->type;
					} else if (r->type < 0 && a2 >= 4 && taken < cap) {
						hit = variant->itb[a2][0] == '-' && (int)strlen(variant->itb[a2]) - 1 == -1 * r->type;
					}
					if (!hit) continue;
					GROW(sb->alt_reads, sb->cap_alt, sb->n_alt + 1, crispgpu_altread);
					crispgpu_altread* e = &sb->alt_reads[sb->n_alt++];
					if (r->l1 + r->delta > 65535 || r->l1 + r->delta < 0 || r->alignedbases > 65535 || r->alignedbases < 0 || i > 65535) return -2;
					e->pool = (uint16_t)i; e->offset = (uint16_t)(r->l1 + r->delta); e->aligned = (uint16_t)r->alignedbases; e->strand = r->strand; e->pad = 0;
					taken++;
				}
			}
		}
		sb->alt_off[s * 5 + k + 1] = (uint32_t)sb->n_alt;
	}
	/* 3. ambiguity rows: what remove_ambiguous_bases(+1) would subtract (process_indel_variant.c:25-62) */
	for (int k = 0; k < 5; k++) {
		sb->amb_idx[s * 5 + k] = -1;
		if (maxvec[k + 1].ct < 3) continue;
		const int a2 = maxvec[k + 1].al, a3 = maxvec[k == 0 ? 2 : 1].al;
		if (!((a2 >= 4 && variant->HPlength[a2] > 0) || (a3 >= 4 && variant->HPlength[a3] >= 100))) continue;
		const int hp = variant->HPlength[a2];
		if (sb->n_amb + 1 > sb->cap_amb) {
			const int c = (sb->n_amb + 1) * 2 + 16;
			int32_t* nd = (int32_t*)realloc(sb->amb_dec, (size_t)c * P * 16);
			if (!nd) return -1;
			sb->amb_dec = nd;
			double* ne = (double*)realloc(sb->amb_ep, (size_t)c * P * 16);
			if (!ne) return -1;
			sb->amb_ep = ne;
			sb->cap_amb = c;
		}
		int32_t* dec = sb->amb_dec + (size_t)sb->n_amb * P * 4;
		double* ep = sb->amb_ep + (size_t)sb->n_amb * P * 2;
		memset(dec, 0, (size_t)P * 16);
		memset(ep, 0, (size_t)P * 16);
		for (struct alignedread* r = bq->first; r != NULL && r->position <= variant->position; r = r->next) {
			if (r->lastpos <= variant->position) continue;
			const int offset = r->strand == 1 ? maxalleles : 0;
			if (!(r->filter == '0' && r->allele - offset == variant->refbase)) continue;
			const int op = r->fcigarlist[r->cigoffset] & 0xf, ol = r->fcigarlist[r->cigoffset] >> 4;
			if (op != BAM_CMATCH || (ol - r->delta <= hp)) {
				const int st = r->strand == 1 ? 1 : 0;
				dec[r->sampleid * 4 + (r->bidir == 1 ? 2 + st : st)]++;
				ep[r->sampleid * 2 + st] += pow(0.1, ((double)r->quality[r->l1 + r->delta] - QVoffset) / 10);
			}
		}
		sb->amb_idx[s * 5 + k] = sb->n_amb++;
	}
	sb->n_sites++;
	return 0;
}

void crisp_shim_fill_variant(const crispgpu_results* r, int s, struct VARIANT* variant)
{
	const int P = variant->samples;
	variant->varalleles = r->varalleles[s];
	variant->strandpass = r->strandpass[s];
	for (int k = 0; k < 5; k++) {
		const int o = s * 5 + k, j = r->call_rank[o];
		if (r->status[o] != CRISPGPU_SLOT_CALLED || j < 0) continue;
		variant->alleles[j] = r->allele[o];
		variant->ctpval[j] = r->ctpval[o];
		variant->chi2pval[j] = r->chi2pval[o];
		variant->qvpvaluef[j] = r->qvpf[o];
		variant->qvpvaluer[j] = r->qvpr[o];
		variant->varpools[j] = variant->nvariants[j] = r->varpools[o];
		struct CRISPVAR* cv = &variant->crispvar[j];
		cv->allele = r->allele[o];
		cv->AF_ML = r->af_ml[o];
		cv->delta = r->delta[o];
		cv->deltaf = r->deltaf[o];
		cv->deltar = r->hwe[o];
		cv->allelecount = r->allelecount[o];
		cv->variantpools = r->variantpools[o];
		const int row = r->call_row[o];
		for (int i = 0; i < P; i++) {
			cv->AC[i] = row >= 0 ? r->ac[(size_t)row * P + i] : 0;
			cv->QV[i] = row >= 0 ? r->qv[(size_t)row * P + i] : 0;
		}
	}
}
