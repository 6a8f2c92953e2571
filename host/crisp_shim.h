/*
 * crisp_shim.h — host-side glue between CRISP's own data structures and the crispgpu C ABI.
 *
 * This is the code a CRISP maintainer compiles into readmultiplebams.c/variantcalls.c: it packs
 * the candidate site the front end has just piled up (struct VARIANT + the per-pool read lists) into a growing
 * struct-of-arrays batch at the point where the reference calls newCRISPcaller (variantcalls.c:115), and copies a
 * result block back into struct VARIANT so that the unchanged print_pooledvariant (crisp/newcrispprint.c:347) can
 * write the record.  It needs the reference's headers (parsebam/variant.h, bamsreader.h) and is therefore compiled
 * only where the reference sources are available (oracle/Makefile builds it into oracle/_ref for the round-trip
 * test tests/test_oracle.py::test_host_shim_round_trip).
 */
#ifndef CRISP_SHIM_H
#define CRISP_SHIM_H
#include "crispgpu.h"

typedef struct crisp_shim_batch {
	int n_pools;
	int n_sites, cap_sites;
	size_t n_reads, cap_reads, n_alt, cap_alt;
	int n_amb, cap_amb;
	int want_qhigh, want_stats;
	int32_t *tid, *pos, *maxvec_ct, *counts, *indcounts, *amb_idx, *amb_dec;
	uint8_t *refbase, *maxvec_al;
	int16_t *indel_len, *hplen;
	uint32_t *read_off, *alt_off;
	crispgpu_read* reads;
	crispgpu_altread* alt_reads;
	double *amb_ep, *qhigh, *stats, *tstats;
	/* compact read format (crispgpu_bin): per (site, pool) the distinct read values with their multiplicities */
	int compact;
	size_t n_bins, cap_bins;
	uint32_t* bin_off;
	crispgpu_bin* bins;
	size_t n_ic, cap_ic;
	uint32_t* ic_off;
	crispgpu_count* ic_sparse; /* compact mode: non-zero entries of every pool's indcounts row */
	uint32_t* bin_count;    /* [4096] scratch: multiplicity of each 12-bit read value in the current cell */
	uint16_t* bin_touched;  /* [4096] scratch: values seen in the current cell */
} crisp_shim_batch;

/* NULL when out of memory.  want_qhigh / want_stats: also snapshot Qhighcounts (read in permutation mode) / stats, tstats
 * (read with --EM 0). */
crisp_shim_batch* crisp_shim_batch_new(int n_pools, int want_qhigh, int want_stats);
/* compact = 1: emit `bins`/`bin_off` instead of one element per read and `ic_sparse`/`ic_off` instead of the dense
 * indcounts table (about a fifth of the bytes on binned-quality data; the engine rebuilds both on the device).
 * Choose before the first append of a batch. */
void crisp_shim_batch_set_compact(crisp_shim_batch* sb, int compact);
void crisp_shim_batch_clear(crisp_shim_batch* sb);
void crisp_shim_batch_free(crisp_shim_batch* sb);
/* view of the accumulated sites as the ABI batch (pointers into sb; valid until the next append/clear) */
void crisp_shim_batch_view(const crisp_shim_batch* sb, crispgpu_batch* out);
/* Bytes crisp_shim_batch_pack needs for the accumulated sites (every array rounded up to 256 bytes). */
size_t crisp_shim_batch_packed_bytes(const crisp_shim_batch* sb);
/* Copy the accumulated arrays back to back (256-byte aligned) into `arena` — pinned memory from crispgpu_host_alloc, at
 * least crisp_shim_batch_packed_bytes(sb) bytes, 256-byte aligned — and describe them in *out with arena_base /
 * arena_bytes set: the engine then moves the batch with one asynchronous copy, and sb can be cleared and refilled while
 * the batch is in flight.  Returns 0, or -1 if the arena is too small. */
int crisp_shim_batch_pack(const crisp_shim_batch* sb, void* arena, size_t arena_cap, crispgpu_batch* out);

#ifdef CRISP_SHIM_WITH_REFERENCE_TYPES
/* Append the site currently held in `variant` (after calculate_allelecounts + sort_allelecounts and the candidate
 * test of variantcalls.c:76-92).  reflist/current are used for calculate_indel_ambiguity exactly as
 * newcrispcaller.c:191-198 does; pass current < 0 to keep variant->HPlength as it is.
 * Returns 0; -1 when an allocation fails; -2 when a value does not fit the batch format (read offset / aligned length
 * above 65535, indel or homopolymer length above 32767, more than 2^32 - 2^28 elements in one batch: flush first).  On
 * failure the site is not part of the batch and the sites appended before stay valid. */
int crisp_shim_append_site(crisp_shim_batch* sb, REFLIST* reflist, int current, READQUEUE* bq, struct BAMFILE_data* bd,
                           struct VARIANT* variant, ALLELE* maxvec, int tid);
/* Copy the results of site s into struct VARIANT (fields read by print_pooledvariant, newcrispprint.c:357-473). */
void crisp_shim_fill_variant(const crispgpu_results* r, int s, struct VARIANT* variant);
#endif
#endif
