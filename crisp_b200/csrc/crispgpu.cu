// crispgpu.cu — C ABI (include/crispgpu.h) of the B200 statistics engine: context, batch transfer, kernel pipeline.
//
// Pipeline per batch, all on one stream:
//   H2D  ->  k_evaluate  ->  [k_permute_* -> k_filter]  ->  k_em | k_legacy  ->  k_finalize  ->  D2H
// The only host synchronisation is one event wait after the screening stage to learn how many (site, slot) tasks
// reached the caller (sizes the per-pool allele-count rows).  There is no CPU compute path.
#include <cuda_runtime.h>

#include <cmath>
#include <cstddef>
#include <cstdio>
#include <cstdlib>
#include <cstring>
#include <new>
#include <vector>

#include "../../include/crispgpu.h"
#include "dmath.cuh"
#include "engine_types.cuh"
#include "k_em.cuh"
#include "k_evaluate.cuh"
#include "k_legacy.cuh"
#include "k_permute.cuh"
#include "k_pileup.cuh"

using namespace crisp;

namespace {

struct DevBuf {
	void* p = nullptr;
	size_t cap = 0;
};

enum { EV_START, EV_H2D, EV_EVAL, EV_PERM, EV_FILTER, EV_CALL, EV_D2H, EV_COUNT };

}  // namespace

struct crispgpu_ctx {
	int device = 0;
	int num_sms = 0;
	size_t smem_optin = 0;
	cudaStream_t stream = nullptr;
	cudaStream_t side = nullptr;  // side stream: the quality-presence scan overlaps the screening kernel
	cudaEvent_t ev_fork = nullptr, ev_join = nullptr;
	bool own_stream = false;
	crispgpu_config cfg{};
	std::vector<int32_t> ploidy, phenotypes, pclass, class_ploidy;
	int P = 0, K = 0, nmax = 0;
	// device configuration tables
	DevBuf d_ploidy, d_pclass, d_class_ploidy, d_NC, d_T, d_ncr, d_pheno;
	DevBuf d_scr[8];               // crispgpu_screen_positions: inputs and outputs of the candidate screen
	DevBuf d_need_reads;           // per-site flags of the on-demand read fetch
	const uint16_t* lazy_reads = nullptr;  // device-visible address of the caller's pinned read array (on-demand fetch), or null
	int64_t copied_bytes = 0;
	bool bins_direct = false;      // k_em reads the bins themselves
	bool device_sort = false;      // maxvec_al/maxvec_ct are produced on the device (k_candidates)
	bool sparse_ic = false;        // the batch came with sparse per-pool allele counts
	bool binned = false;           // the batch came in the compact (value, count) read format
	bool narrow = false;           // ... in its narrow transport form (bins16 / ic16): widened on the device
	const void* nar[7] = {nullptr, nullptr, nullptr, nullptr, nullptr, nullptr, nullptr};  // device addresses of bins16, bin_len, bin_site, ic16, ic_len, ic_site, alt32
	DevBuf d_nar[7];
	bool narrow_counts = false;    // the narrow batch came without site totals: k_expand_counts16 writes them
	size_t n_alt32 = 0;            // alt-read records that came as 4-byte words (k_unpack_alt32)
	DevBuf d_exact_row, d_exact_cl, d_exact_done;  // exact_mode: term tables of the exact 2 x k test, per-task decided flags
	// device batch
	DevBuf d_b[24];
	DevBatch db{};
	int n_sites = 0;
	size_t n_reads = 0;
	size_t n_bins = 0;
	bool resident = false;
	// device results + pinned mirrors
	DevBuf d_o[32], d_ac, d_qv;     // d_o / h_o: views into the two result arenas (one device-to-host copy per batch)
	DevBuf h_o[32], h_ac, h_qv;
	DevBuf d_out_arena, h_out_arena;
	DevBuf d_arena;                 // device image of a batch the front end packed into one pinned arena
	DevBuf d_rows, h_rows;          // AC/QV rows of the batch: qv[rows][P] doubles followed by ac[rows][P] ints
	bool timing_pending = false;    // the stage durations of the last batch have not been read from the events yet
	size_t out_bytes = 0;
	DevOut dout{};
	// queues
	DevBuf d_perm_tasks, d_call_tasks, d_call_tasks2, d_order, d_counters, h_counters, d_gws, d_trace;
	DevBuf d_tail_task, d_tail_head, d_tail_total, d_tail_rec, d_tail_counters;  // long permutation tasks (k_permute_stranded_tail)
	size_t gws_stride = 0;
	// launch geometry
	int ev_warps = 8, ev_grid = 0;
	size_t ev_smem = 0;
	int em_threads = 256, em_grid = 0, em_warps = 8, em_key = -1;
	int nq = 0, dense_lik = 0, nobidir = 0;
	size_t em_smem2 = 0;       // shared memory of the two-plane (no bidirectional reads) instantiation
	bool em_smem_g2 = false;   // its planes fit in shared memory
	int em_grid2 = 0;
	uint8_t qlist[16] = {0};
	size_t em_smem = 0;
	bool em_smem_g = true;
	int pm_threads = 256, pm_grid = 0, pm_l10cap = 2048;
	size_t pm_smem = 0;
	int ps_l10cap = 2048;
	bool ps_many = false;              // many-pool geometry of the stranded kernels (hash tables / narrow columns over a fixed area)
	int ps_threads = 0, ps_area = 0;   // stranded permutation kernels: block size and bytes of the occupancy area (hash tables at many pools)
	size_t ps_smem = 0;
	int lg_grid = 0;
	size_t lg_smem = 0;
	// bookkeeping
	cudaEvent_t ev[EV_COUNT]{};
	cudaEvent_t ev_counts = nullptr;
	cudaEvent_t ev_done = nullptr;   // blocking-sync twin of EV_D2H (config.blocking_wait): crispgpu_wait sleeps on it
	cudaEvent_t ev_epoch = nullptr;  // recorded at creation (crispgpu_debug_timeline)
	int64_t ticket = 0;
	bool inflight = false;
	std::vector<const void*> prepared;  // kernels whose shared-memory attributes this context has set
	bool advanced = false;         // stage 2 of the batch in flight has been enqueued (crispgpu_advance)
	crispgpu_timing timing{};
	crispgpu_results res{};
	int n_call = 0, n_call1 = 0, n_call2 = 0, n_perm = 0;
	int64_t stage1_launches = 0;
	bool fetch_pending = false;
	bool rows_on_host = false;     // k_em wrote the AC/QV rows of the batch into pinned host memory
	std::vector<unsigned char> zeros;  // host view of result fields this configuration never produces
	// pileup on the device (crispgpu_pileup_run)
	DevBuf p_pool_off, p_rec, p_mate, p_isize, p_seq4, p_qual, p_ref, p_flags, p_state, p_win, p_flag, p_cand;
	DevBuf p_sites, p_work, p_reads, p_alt;   // site arrays (mirrored on the host), scan scratch, read stream, alt-read lists
	DevBuf ph_state, ph_sites, ph_reads, ph_alt;
	PileReads prd{};
	PileWindow pwin{};
	bool pile_resident = false, pile_timing = false, pile_streams = false;
	int64_t pile_h2d = 0;
	cudaEvent_t ev_pile[4]{};       // pileup start, k_pile_tile start / end, pileup end
	char err[512] = {0};
};

namespace {

#define CK(call)                                                                                             \
	do {                                                                                                     \
		cudaError_t e_ = (call);                                                                             \
		if (e_ != cudaSuccess) {                                                                             \
			snprintf(ctx->err, sizeof ctx->err, "%s:%d %s: %s", __FILE__, __LINE__, #call, cudaGetErrorString(e_)); \
			return CRISPGPU_ERR_CUDA;                                                                        \
		}                                                                                                    \
	} while (0)

// after every launch; CRISPGPU_SYNC_CHECK=1 also waits for the device so that a faulting kernel is reported at its own line
cudaError_t launch_check()
{
	static const bool sync = getenv("CRISPGPU_SYNC_CHECK") != nullptr;
	cudaError_t e = cudaGetLastError();
	if (e == cudaSuccess && sync) e = cudaDeviceSynchronize();
	return e;
}

int ensure_dev(crispgpu_ctx* ctx, DevBuf& b, size_t bytes)
{
	if (bytes <= b.cap) return 0;
	if (b.p) CK(cudaFree(b.p));
	b.p = nullptr;
	b.cap = 0;
	size_t cap = bytes + bytes / 4 + 256;
	CK(cudaMalloc(&b.p, cap));
	b.cap = cap;
	return 0;
}

int ensure_host(crispgpu_ctx* ctx, DevBuf& b, size_t bytes)
{
	if (bytes <= b.cap) return 0;
	if (b.p) CK(cudaFreeHost(b.p));
	b.p = nullptr;
	b.cap = 0;
	size_t cap = bytes + bytes / 4 + 256;
	CK(cudaMallocHost(&b.p, cap));
	b.cap = cap;
	return 0;
}

// result arrays in the order of DevOut / crispgpu_results; width = elements per site
// need: which configurations read the field on the host — 0 always (what print_pooledvariant consumes), 1 only with --EM 0,
// 2 diagnostics (dropped with config.minimal_results), 3 only in permutation mode, 4 only with --phenotypes.  A context lays
// its result arena out with the fields it needs first and copies only that prefix device-to-host; the others stay zero.
struct OutField { size_t elem; int width; int need; };
const OutField kOutFields[] = {
	{4, 1, 0}, {4, 1, 1},                 // varalleles, strandpass
	{1, 5, 0}, {1, 5, 0}, {1, 5, 0}, {4, 5, 0}, // status, allele, call_rank, call_row
	{8, 5, 2}, {8, 5, 0}, {8, 5, 2}, {8, 5, 0}, {8, 5, 0}, {8, 5, 0}, {8, 5, 0}, {8, 5, 1}, {8, 5, 1},  // paf ctpval chi2pval af_ml delta deltaf hwe qvpf qvpr
	{4, 5, 0}, {4, 5, 0}, {4, 5, 0}, {4, 5, 2}, {4, 5, 3}, {4, 5, 3},  // varpools allelecount variantpools em_iters perm_k perm_hits
	{8, 5, 4}, {8, 5, 4}, {8, 5, 4}, {8, 5, 4}, {8, 5, 4}, {8, 5, 4},  // assoc_p assoc_llr assoc_or assoc_p1 assoc_llr1 assoc_or1
};
constexpr int kNumOut = sizeof(kOutFields) / sizeof(kOutFields[0]);
// DevOut, crispgpu_results and DevBatch are walked as arrays of pointers in the order of kOutFields / collect_arrays
static_assert(sizeof(DevOut) == (kNumOut + 2) * sizeof(void*), "DevOut must be kNumOut result pointers followed by ac, qv");
static_assert(sizeof(crispgpu_results) == 8 + (kNumOut + 2) * sizeof(void*), "crispgpu_results layout");
static_assert(offsetof(crispgpu_results, varalleles) == 8 && offsetof(crispgpu_results, ac) == 8 + kNumOut * sizeof(void*), "crispgpu_results layout");
static_assert(offsetof(DevBatch, tid) == 8 && sizeof(DevBatch) == 8 + 23 * sizeof(void*), "DevBatch must be 23 array pointers after two ints");

void bind_out(crispgpu_ctx* ctx)
{
	void** d = reinterpret_cast<void**>(&ctx->dout);
	for (int f = 0; f < kNumOut; f++) d[f] = ctx->d_o[f].p;
	ctx->dout.ac = static_cast<int32_t*>(ctx->d_ac.p);
	ctx->dout.qv = static_cast<double*>(ctx->d_qv.p);
}

size_t em_smem_fixed(int P, int J, int nwarps, bool dense, int nq, bool nobidir = false)
{
	size_t off = (sizeof(EmShared) + 15) & ~size_t(15);
	off += sizeof(double) * P;
	off += sizeof(double) * J;
	off += sizeof(int32_t) * 6 * P;
	off = (off + 15) & ~size_t(15);
	off += dense ? sizeof(uint32_t) * 2 * nq * 3 * P : sizeof(uint16_t) * 6 * kNQ * nwarps;
	off = (off + 15) & ~size_t(15);
	off += dense ? sizeof(double) * 2 * nq * J + kNQ : 0;
	off = (off + 15) & ~size_t(15);
	const int rounds = (P + 31) / 32;
	off += sizeof(double) * 2 * 19 * rounds + sizeof(int32_t) * 2 * rounds;
	off = (off + 15) & ~size_t(15);
	off += sizeof(double) * 3 * P;
	off += nobidir ? sizeof(double) * J : 0;
	off = (off + 15) & ~size_t(15);
	return off;
}

int setup_geometry(crispgpu_ctx* ctx)
{
	const int P = ctx->P, J = ctx->nmax + 1;
	const size_t budget = ctx->smem_optin;
	// k_evaluate: warps per CTA limited by the staged count block
	ctx->ev_warps = 4;
	ctx->ev_smem = 0;
	// k_em
	const int rounds = (P + 31) / 32;
	if (rounds > kMaxRounds) return CRISPGPU_ERR_UNSUPPORTED;
	int warps = rounds * 3;
	if (warps > kEmMaxWarps) warps = kEmMaxWarps;
	ctx->em_threads = warps * 32;
	ctx->em_warps = warps;
	const size_t fixed = em_smem_fixed(P, J, warps, false, 0);
	const size_t gbytes = (size_t)3 * J * P * sizeof(double);
	ctx->em_smem_g = 2 * (fixed + gbytes + 1024) <= budget;  // at least two CTAs per SM, else planes go to the global workspace
	ctx->em_smem = ctx->em_smem_g ? fixed + gbytes : fixed;
	ctx->gws_stride = (size_t)3 * J * P;
	// k_permute: threads per CTA limited by the per-permutation occupancy columns
	ctx->pm_l10cap = P > 2048 ? P : 2048;
	size_t pfixed = ((sizeof(PermShared) + 15) & ~size_t(15)) + sizeof(uint32_t) * 3 * (P + 1) + sizeof(int32_t) * 4 * P + 32 + sizeof(double) * (ctx->pm_l10cap > P ? ctx->pm_l10cap : P);
	ctx->pm_threads = 256;
	while (ctx->pm_threads > 32 && pfixed + (size_t)P * ctx->pm_threads * 4 > budget - 1024) ctx->pm_threads >>= 1;
	ctx->pm_smem = pfixed + (size_t)P * ctx->pm_threads * 4;
	if (ctx->pm_smem > budget) ctx->pm_smem = 0;  // permutation modes unsupported for this many pools
	// stranded test: below kPermHashPools pools the occupancy area is the [P][T] columns; from there on full blocks over a fixed
	// area, in which a permutation keeps a hash table sized by the balls of its table (or a column, when the balls are many)
	ctx->ps_threads = ctx->pm_threads;
	ctx->ps_area = (int)((size_t)P * ctx->pm_threads * 4);
	ctx->ps_smem = ctx->pm_smem;
	ctx->ps_l10cap = ctx->pm_l10cap;
	if (P >= kPermHashPools && ctx->pm_smem) {
		// many pools are shallow pools: a short log table (deeper tables compute the logarithms), three CTAs per SM
		const int l10 = P > 512 ? P : 512;
		const size_t sfixed = pfixed - sizeof(double) * (ctx->pm_l10cap > P ? ctx->pm_l10cap : P) + sizeof(double) * l10;
		size_t area = (budget - 1024) / 3 > sfixed + 2048 ? ((budget - 1024) / 3 - sfixed - 1024) & ~size_t(255) : 0;
		if (area < (size_t)P * 4 * 32) area = (size_t)P * 4 * 32;   // at least one warp of 32-bit columns
		if (sfixed + area <= budget - 1024) { ctx->ps_threads = 256; ctx->ps_area = (int)area; ctx->ps_smem = sfixed + area; ctx->ps_l10cap = l10; ctx->ps_many = true; }
	}
	ctx->lg_smem = (size_t)4 * P * sizeof(double);
	return 0;
}

// Function attributes are process-wide: the dynamic shared-memory limit is always raised to everything the device allows
// (never a per-context size), so contexts with different pool counts on other threads cannot lower it under a launch.
template <typename K>
int prepare_kernel(crispgpu_ctx* ctx, K kernel, size_t smem)
{
	if (smem == 0) return 0;
	const void* key = reinterpret_cast<const void*>(kernel);
	for (const void* p : ctx->prepared) if (p == key) return 0;
	cudaFuncAttributes fa{};
	CK(cudaFuncGetAttributes(&fa, kernel));
	CK(cudaFuncSetAttribute(kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)(ctx->smem_optin - fa.sharedSizeBytes)));
	CK(cudaFuncSetAttribute(kernel, cudaFuncAttributePreferredSharedMemoryCarveout, cudaSharedmemCarveoutMaxShared));
	ctx->prepared.push_back(key);
	return 0;
}

template <typename K>
int occupancy_grid(crispgpu_ctx* ctx, K kernel, int threads, size_t smem, int* grid)
{
	int per_sm = 0;
	{ int r0 = prepare_kernel(ctx, kernel, smem); if (r0) return r0; }
	CK(cudaOccupancyMaxActiveBlocksPerMultiprocessor(&per_sm, kernel, threads, smem));
	if (per_sm < 1) per_sm = 1;
	*grid = per_sm * ctx->num_sms;
	return 0;
}

// Threads and shared memory of a permutation tail kernel whose per-thread occupancy columns take `occ_bytes` per pool
// (+ the touched-pool bitmaps of the low-coverage test): the largest power-of-two block that fits.
void perm_tail_geometry(const crispgpu_ctx* ctx, int occ_bytes, bool bitmaps, int* threads, size_t* smem)
{
	const int P = ctx->P;
	const size_t pfixed = ctx->pm_smem - (size_t)P * ctx->pm_threads * 4;
	const size_t per_thread = (size_t)P * occ_bytes + (bitmaps ? (size_t)((P + 31) / 32) * 4 : 0);
	int T = 256;
	while (T > 32 && pfixed + per_thread * T > ctx->smem_optin - 1024) T >>= 1;
	*threads = T;
	*smem = pfixed + per_thread * T;
}

int launch_evaluate(crispgpu_ctx* ctx, const EngineParams& prm)
{
	const int n = ctx->n_sites;
	auto go = [&](auto kernel, int warps) -> int {
		if (!ctx->ev_grid) {  // one instantiation per context (the mode is part of the configuration): query once
			int rc = occupancy_grid(ctx, kernel, warps * 32, ctx->ev_smem, &ctx->ev_grid);
			if (rc) return rc;
		}
		int grid = ctx->ev_grid * 4;  // several waves of short CTAs: sites differ in the number of tested slots
		int need = (n + warps - 1) / warps;
		if (grid > need) grid = need;
		if (grid < 1) grid = 1;
		kernel<<<grid, warps * 32, ctx->ev_smem, ctx->stream>>>(prm);
		return 0;
	};
	const bool perm = ctx->cfg.chisq_permutation == 1, pivot = ctx->cfg.pivot_sample > 0;
	int rc = perm ? (pivot ? go(k_evaluate<4, true, true>, 4) : go(k_evaluate<4, true, false>, 4))
	              : (pivot ? go(k_evaluate<4, false, true>, 4) : go(k_evaluate<4, false, false>, 4));
	if (rc) return rc;
	// p-values, filters, queues and the result zero fill: one thread per (site, slot)
	const int fgrid = (n * kSlots + 255) / 256;
	if (perm) { if (pivot) k_eval_finish<true, true><<<fgrid, 256, 0, ctx->stream>>>(prm); else k_eval_finish<true, false><<<fgrid, 256, 0, ctx->stream>>>(prm); }
	else { if (pivot) k_eval_finish<false, true><<<fgrid, 256, 0, ctx->stream>>>(prm); else k_eval_finish<false, false><<<fgrid, 256, 0, ctx->stream>>>(prm); }
	return 0;
}

struct BatchArray { const void* host; size_t bytes; };

int collect_arrays(crispgpu_ctx* ctx, const crispgpu_batch* b, BatchArray* a)
{
	const size_t n = b->n_sites, P = ctx->P;
	const bool narrow = b->bin_site != nullptr && b->ic_site != nullptr && b->bin_len != nullptr && b->ic_len != nullptr;
	if (narrow && ((b->bin_site[n] && !b->bins16) || (b->ic_site[n] && !b->ic16))) return CRISPGPU_ERR_ARG;
	const bool sparse_ic = !narrow && b->ic_off != nullptr && (b->ic_sparse != nullptr || (n && b->ic_off[n * P] == 0));
	const bool binned = !narrow && (b->bins != nullptr || (b->bin_off != nullptr && n && b->bin_off[n * P] == 0));
	if (!b->tid || !b->pos || !b->refbase || (!b->maxvec_al != !b->maxvec_ct) || (!b->counts && !narrow) || (!b->indcounts && !sparse_ic && !narrow) || (!b->read_off && !binned && !narrow)) return CRISPGPU_ERR_ARG;
	const size_t nreads = (n && !binned && !narrow) ? b->read_off[n * P] : 0;
	const size_t nalt = (b->alt_off && n) ? b->alt_off[n * 5] : 0;
	if (nreads && !b->reads) return CRISPGPU_ERR_ARG;
	if (binned && !b->bin_off) return CRISPGPU_ERR_ARG;
	const bool alt_narrow = narrow && b->alt32 != nullptr;
	if (nalt && !b->alt_reads && !alt_narrow) return CRISPGPU_ERR_ARG;
	if (b->amb_idx && b->n_amb > 0 && !b->amb_dec) return CRISPGPU_ERR_ARG;
	int k = 0;
	a[k++] = {b->tid, n * 4};
	a[k++] = {b->pos, n * 4};
	a[k++] = {b->refbase, n};
	a[k++] = {b->maxvec_al, n * 8};
	a[k++] = {b->maxvec_ct, n * 32};
	a[k++] = {b->counts, n * 96};
	a[k++] = {narrow ? nullptr : b->indcounts, n * P * 96};
	a[k++] = {b->indel_len, b->indel_len ? n * 16 : 0};
	a[k++] = {b->hplen, b->hplen ? n * 16 : 0};
	a[k++] = {(binned || narrow) ? nullptr : b->read_off, (n * P + 1) * 4};  // a compact batch has no read stream to index
	a[k++] = {narrow ? nullptr : b->reads, nreads * 2};
	a[k++] = {b->alt_off, b->alt_off ? (n * 5 + 1) * 4 : 0};
	a[k++] = {alt_narrow ? nullptr : b->alt_reads, nalt * sizeof(crispgpu_altread)};
	a[k++] = {b->amb_idx, b->amb_idx ? n * 20 : 0};
	a[k++] = {b->amb_dec, (b->amb_idx && b->amb_dec) ? (size_t)b->n_amb * P * 16 : 0};
	a[k++] = {b->amb_ep, (b->amb_idx && b->amb_ep) ? (size_t)b->n_amb * P * 16 : 0};
	a[k++] = {b->qhigh, b->qhigh ? n * P * 128 : 0};
	a[k++] = {b->stats, b->stats ? n * P * 128 : 0};
	a[k++] = {b->tstats, b->tstats ? n * 128 : 0};
	a[k++] = {binned ? b->bin_off : nullptr, binned ? (n * P + 1) * 4 : 0};
	a[k++] = {binned ? b->bins : nullptr, binned && n ? (size_t)b->bin_off[n * P] * 4 : 0};
	if (binned) a[10].host = nullptr;  // the read array is rebuilt on the device
	a[k++] = {sparse_ic ? b->ic_off : nullptr, sparse_ic ? (n * P + 1) * 4 : 0};
	a[k++] = {sparse_ic ? b->ic_sparse : nullptr, sparse_ic && n ? (size_t)b->ic_off[n * P] * 4 : 0};
	if (sparse_ic) a[6].host = nullptr;  // the dense table is rebuilt on the device
	// narrow transport arrays (widened on the device into a[19], a[20] and the dense a[6])
	a[k++] = {narrow ? b->bins16 : nullptr, narrow ? (size_t)b->bin_site[n] * 2 : 0};
	a[k++] = {narrow ? b->bin_len : nullptr, narrow ? n * P : 0};
	a[k++] = {narrow ? b->bin_site : nullptr, narrow ? (n + 1) * 4 : 0};
	a[k++] = {narrow ? b->ic16 : nullptr, narrow ? (size_t)b->ic_site[n] * 2 : 0};
	a[k++] = {narrow ? b->ic_len : nullptr, narrow ? n * P : 0};
	a[k++] = {narrow ? b->ic_site : nullptr, narrow ? (n + 1) * 4 : 0};
	a[k++] = {alt_narrow ? b->alt32 : nullptr, alt_narrow ? nalt * 4 : 0};
	return k;
}

// CRISPGPU_VALIDATE=1: host-side consistency checks of the offset arrays before anything is copied (the kernels trust them;
// a non-monotone offset or an out-of-range row index would be an out-of-bounds device read).  O(n P) on the host, so off
// by default; a front end under development switches it on.
int validate_batch(crispgpu_ctx* ctx, const crispgpu_batch* b)
{
	static const bool on = getenv("CRISPGPU_VALIDATE") != nullptr;
	if (!on) return 0;
	const size_t n = b->n_sites, P = ctx->P;
	auto monotone = [&](const uint32_t* off, size_t cells, const char* name) -> int {
		if (!off) return 0;
		if (off[0] != 0) { snprintf(ctx->err, sizeof ctx->err, "%s[0] is %u, not 0", name, off[0]); return CRISPGPU_ERR_ARG; }
		for (size_t k = 0; k < cells; k++)
			if (off[k + 1] < off[k]) { snprintf(ctx->err, sizeof ctx->err, "%s decreases at cell %zu (%u -> %u)", name, k, off[k], off[k + 1]); return CRISPGPU_ERR_ARG; }
		return 0;
	};
	int rc;
	if ((rc = monotone(b->read_off, n * P, "read_off"))) return rc;
	if ((rc = monotone(b->bin_off, n * P, "bin_off"))) return rc;
	if ((rc = monotone(b->ic_off, n * P, "ic_off"))) return rc;
	if ((rc = monotone(b->alt_off, n * 5, "alt_off"))) return rc;
	if (b->bin_site && b->bin_len && b->ic_site && b->ic_len) {
		if ((rc = monotone(b->bin_site, n, "bin_site")) || (rc = monotone(b->ic_site, n, "ic_site"))) return rc;
		for (size_t s2 = 0; s2 < n; s2++) {
			size_t nb = 0, ni = 0;
			for (size_t i = 0; i < P; i++) { nb += b->bin_len[s2 * P + i]; ni += b->ic_len[s2 * P + i]; }
			if (nb != b->bin_site[s2 + 1] - b->bin_site[s2] || ni != b->ic_site[s2 + 1] - b->ic_site[s2]) {
				snprintf(ctx->err, sizeof ctx->err, "narrow batch: the cell lengths of site %zu do not add up to its run of entries", s2);
				return CRISPGPU_ERR_ARG;
			}
		}
	}
	if (b->amb_idx)
		for (size_t k = 0; k < n * 5; k++)
			if (b->amb_idx[k] >= b->n_amb || b->amb_idx[k] < -1) { snprintf(ctx->err, sizeof ctx->err, "amb_idx[%zu] = %d outside [-1, %d)", k, b->amb_idx[k], b->n_amb); return CRISPGPU_ERR_ARG; }
	if (b->alt_off && b->alt_reads)
		for (size_t k = 0; k < b->alt_off[n * 5]; k++)
			if (b->alt_reads[k].pool >= P) { snprintf(ctx->err, sizeof ctx->err, "alt_reads[%zu].pool = %u with %zu pools", k, b->alt_reads[k].pool, P); return CRISPGPU_ERR_ARG; }
	if (b->maxvec_al)
		for (size_t k = 0; k < n * 8; k++)
			if (b->maxvec_al[k] >= 8) { snprintf(ctx->err, sizeof ctx->err, "maxvec_al[%zu] = %u is not an allele index", k, b->maxvec_al[k]); return CRISPGPU_ERR_ARG; }
	for (size_t k = 0; k < n; k++)
		if (b->refbase[k] >= 8) { snprintf(ctx->err, sizeof ctx->err, "refbase[%zu] = %u", k, b->refbase[k]); return CRISPGPU_ERR_ARG; }
	if (b->ic_off && b->ic_sparse)
		for (size_t k = 0; k < b->ic_off[n * P]; k++)
			if ((b->ic_sparse[k] >> 27) >= 24) { snprintf(ctx->err, sizeof ctx->err, "ic_sparse[%zu] has index %u (>= 24)", k, b->ic_sparse[k] >> 27); return CRISPGPU_ERR_ARG; }
	return 0;
}

// lazy: the caller's arrays outlive the call (submit/wait), so the packed reads — 80 % of a batch, needed only by the
// sites that reach the quality-value likelihood — may stay in pinned host memory and be fetched on demand.
int copy_batch(crispgpu_ctx* ctx, const crispgpu_batch* b, bool lazy)
{
	BatchArray a[32];
	const int k = collect_arrays(ctx, b, a);
	if (k < 0) {
		snprintf(ctx->err, sizeof ctx->err, "batch is missing a required array");
		return k;
	}
	{ const int vr = validate_batch(ctx, b); if (vr) return vr; }
	const void** dptr = reinterpret_cast<const void**>(&ctx->db.tid);
	int64_t bytes = 0;
	constexpr int kReadsArray = 10;
	constexpr size_t kLazyMinBytes = 1 << 20;
	const bool reads_used = ctx->cfg.em_flag >= 1 && ctx->cfg.use_base_qvs == 1;
	ctx->lazy_reads = nullptr;
	ctx->narrow = a[25].host != nullptr;   // bin_site present
	ctx->narrow_counts = false;
	ctx->n_alt32 = 0;
	ctx->binned = a[19].host != nullptr || ctx->narrow;
	ctx->device_sort = false;
	ctx->sparse_ic = a[21].host != nullptr;
	// one-copy path: every array lies in the arena the batch declares
	const unsigned char* abase = static_cast<const unsigned char*>(b->arena_base);
	const bool arena = abase != nullptr && b->arena_bytes > 0;
	if (arena) {
		if (reinterpret_cast<uintptr_t>(abase) & 255) { snprintf(ctx->err, sizeof ctx->err, "batch arena is not 256-byte aligned"); return CRISPGPU_ERR_ARG; }
		for (int i = 0; i < k; i++) {
			if (a[i].bytes == 0 || a[i].host == nullptr) continue;
			const unsigned char* p0 = static_cast<const unsigned char*>(a[i].host);
			if (p0 < abase || p0 + a[i].bytes > abase + b->arena_bytes || (reinterpret_cast<uintptr_t>(p0) & 15)) {
				snprintf(ctx->err, sizeof ctx->err, "batch array %d lies outside the declared arena (or is not 16-byte aligned)", i);
				return CRISPGPU_ERR_ARG;
			}
		}
		int rc = ensure_dev(ctx, ctx->d_arena, b->arena_bytes);
		if (rc) return rc;
		CK(cudaMemcpyAsync(ctx->d_arena.p, abase, b->arena_bytes, cudaMemcpyHostToDevice, ctx->stream));
		bytes += (int64_t)b->arena_bytes;
	}
	for (int i = 0; i < k; i++) {
		if (i >= 23) {   // narrow transport arrays: not part of the device batch, their widened forms are
			const void*& dst = ctx->nar[i - 23];
			dst = nullptr;
			if (!a[i].bytes || !a[i].host) continue;
			if (arena) { dst = static_cast<unsigned char*>(ctx->d_arena.p) + (static_cast<const unsigned char*>(a[i].host) - abase); continue; }
			int rc = ensure_dev(ctx, ctx->d_nar[i - 23], a[i].bytes);
			if (rc) return rc;
			CK(cudaMemcpyAsync(ctx->d_nar[i - 23].p, a[i].host, a[i].bytes, cudaMemcpyHostToDevice, ctx->stream));
			bytes += (int64_t)a[i].bytes;
			dst = ctx->d_nar[i - 23].p;
			continue;
		}
		if (ctx->narrow && i == 5 && a[5].host == nullptr) {   // site totals summed on the device
			int rc = ensure_dev(ctx, ctx->d_b[5], a[5].bytes ? a[5].bytes : 16);
			if (rc) return rc;
			dptr[5] = ctx->d_b[5].p;
			ctx->narrow_counts = true;
			continue;
		}
		if (ctx->narrow && i == 12 && a[29].host != nullptr) {   // 4-byte alt-read records widened on the device
			int rc = ensure_dev(ctx, ctx->d_b[12], a[12].bytes ? a[12].bytes : 16);
			if (rc) return rc;
			dptr[12] = ctx->d_b[12].p;
			ctx->n_alt32 = a[29].bytes / 4;
			continue;
		}
		if (ctx->narrow && (i == 6 || i == 19 || i == 20)) {   // widened on the device (k_expand_counts16, k_narrow_offsets, k_unpack_bins16)
			const size_t nb16 = a[23].bytes / 2;
			const size_t need = i == 6 ? a[6].bytes : (i == 19 ? ((size_t)b->n_sites * ctx->P + 1) * 4 : nb16 * 4);
			int rc = ensure_dev(ctx, ctx->d_b[i], need ? need : 16);
			if (rc) return rc;
			dptr[i] = ctx->d_b[i].p;
			if ((i == 19 || i == 20) && !reads_used) dptr[i] = nullptr;
			continue;
		}
		if (arena && a[i].bytes && a[i].host != nullptr) {
			dptr[i] = static_cast<unsigned char*>(ctx->d_arena.p) + (static_cast<const unsigned char*>(a[i].host) - abase);
			continue;
		}
		if ((i == 3 || i == 4) && a[i].host == nullptr && a[i].bytes) {  // no maxvec from the host: k_candidates sorts the alleles
			int rc = ensure_dev(ctx, ctx->d_b[i], a[i].bytes);
			if (rc) return rc;
			dptr[i] = ctx->d_b[i].p;
			ctx->device_sort = true;
			continue;
		}
		if (i == 6 && ctx->sparse_ic) {
			int rc = ensure_dev(ctx, ctx->d_b[i], a[i].bytes ? a[i].bytes : 1);
			if (rc) return rc;
			dptr[i] = ctx->d_b[i].p;
			continue;
		}
		if (i == kReadsArray && ctx->binned) { dptr[i] = nullptr; continue; }  // k_em consumes the bins; no read stream on the device
		if (i == 20 && ctx->binned && a[i].bytes == 0) {  // a compact batch without a single bin still takes the bins path (all cells empty)
			int rc = ensure_dev(ctx, ctx->d_b[i], 16);
			if (rc) return rc;
			dptr[i] = ctx->d_b[i].p;
			continue;
		}
		if (a[i].bytes == 0 || a[i].host == nullptr) { dptr[i] = nullptr; continue; }
		int rc = ensure_dev(ctx, ctx->d_b[i], a[i].bytes);
		if (rc) return rc;
		if ((i == 19 || i == 20) && !reads_used) { dptr[i] = nullptr; continue; }
		if (i == kReadsArray && !reads_used) { dptr[i] = ctx->d_b[i].p; continue; }  // --EM 0 / --usebq 0: no kernel touches the reads
		static const bool no_fetch = getenv("CRISPGPU_NO_FETCH") != nullptr;  // measurement switch: copy the whole read array instead
		if (i == kReadsArray && lazy && !no_fetch) {
			cudaPointerAttributes at{};
			if (cudaPointerGetAttributes(&at, a[i].host) == cudaSuccess && at.type == cudaMemoryTypeHost && at.devicePointer &&
			    (reinterpret_cast<uintptr_t>(at.devicePointer) & 15) == 0 && a[i].bytes >= kLazyMinBytes) {
				ctx->lazy_reads = static_cast<const uint16_t*>(at.devicePointer);
				dptr[i] = ctx->d_b[i].p;
				continue;
			}
			(void)cudaGetLastError();  // pageable memory: cudaPointerGetAttributes may flag an error on older drivers
		}
		CK(cudaMemcpyAsync(ctx->d_b[i].p, a[i].host, a[i].bytes, cudaMemcpyHostToDevice, ctx->stream));
		dptr[i] = ctx->d_b[i].p;
		bytes += (int64_t)a[i].bytes;
	}
	ctx->db.n_sites = b->n_sites;
	ctx->db.P = ctx->P;
	ctx->n_sites = b->n_sites;
	ctx->n_reads = (b->n_sites && !ctx->binned) ? b->read_off[(size_t)b->n_sites * ctx->P] : 0;
	ctx->n_bins = !(ctx->binned && b->n_sites) ? 0 : (ctx->narrow ? b->bin_site[b->n_sites] : b->bin_off[(size_t)b->n_sites * ctx->P]);
	ctx->timing.h2d_bytes = bytes;
	ctx->copied_bytes = bytes;
	return 0;
}

int ensure_outputs(crispgpu_ctx* ctx, int n)
{
	// all per-site / per-slot result arrays live back to back (256-byte aligned) in one device arena mirrored by one pinned
	// host arena, so that a batch's results come home in a single copy
	const crispgpu_config& g = ctx->cfg;
	auto needed = [&](int need) {
		return need == 0 || (need == 1 && g.em_flag == 0) || (need == 2 && !g.minimal_results) || (need == 3 && g.chisq_permutation == 1) ||
		       (need == 4 && g.association == 1);
	};
	size_t off[kNumOut], total = 0, copy_bytes = 0;
	for (int pass = 0; pass < 2; pass++) {  // needed fields first, then the rest (device-only scratch the kernels still write)
		for (int f = 0; f < kNumOut; f++) {
			if (needed(kOutFields[f].need) != (pass == 0)) continue;
			off[f] = total;
			total += ((size_t)n * kOutFields[f].width * kOutFields[f].elem + 255) & ~size_t(255);
		}
		if (pass == 0) copy_bytes = total;
	}
	int rc = ensure_dev(ctx, ctx->d_out_arena, total);
	if (rc) return rc;
	rc = ensure_host(ctx, ctx->h_out_arena, copy_bytes);
	if (rc) return rc;
	if (ctx->zeros.size() < (size_t)n * 40) ctx->zeros.assign((size_t)n * 40 + 4096, 0);  // what the fields that are never copied read as
	for (int f = 0; f < kNumOut; f++) {
		ctx->d_o[f].p = static_cast<unsigned char*>(ctx->d_out_arena.p) + off[f];
		ctx->h_o[f].p = needed(kOutFields[f].need) ? static_cast<void*>(static_cast<unsigned char*>(ctx->h_out_arena.p) + off[f]) : static_cast<void*>(ctx->zeros.data());
	}
	ctx->out_bytes = copy_bytes;
	rc = ensure_dev(ctx, ctx->d_perm_tasks, (size_t)n * 5 * 4);
	if (rc) return rc;
	rc = ensure_dev(ctx, ctx->d_call_tasks, (size_t)n * 5 * 4);
	if (rc) return rc;
	rc = ensure_dev(ctx, ctx->d_call_tasks2, (size_t)n * 5 * 4);
	if (rc) return rc;
	rc = ensure_dev(ctx, ctx->d_order, (size_t)n * 5 * 4);
	if (rc) return rc;
	return 0;
}

EngineParams make_params(crispgpu_ctx* ctx)
{
	EngineParams prm{};
	prm.b = ctx->db;
	DevCfg& c = prm.c;
	const crispgpu_config& g = ctx->cfg;
	c.P = ctx->P; c.K = ctx->K; c.nmax = ctx->nmax; c.n_pclass = (int)ctx->class_ploidy.size(); c.ploidy0 = ctx->ploidy[0];
	c.varpoolsize = g.varpoolsize;
	c.em_flag = g.em_flag; c.max_iter = g.max_iter; c.chisq_permutation = g.chisq_permutation; c.fast_filter = g.fast_filter;
	c.use_base_qvs = g.use_base_qvs; c.min_reads = g.min_reads; c.qv_offset = g.qv_offset; c.min_q = g.min_q; c.pivot_sample = g.pivot_sample;
	c.use_qv = g.use_qv; c.exact_mode = g.exact_mode;
	c.association = (g.association == 1 && ctx->d_pheno.p) ? 1 : 0;
	c.phenotypes = static_cast<const int32_t*>(ctx->d_pheno.p);
	c.thresh1 = g.thresh1; c.min_qpv = g.min_qpv; c.thresh3 = g.thresh3; c.ser = g.ser; c.relax = g.relax; c.qmin = g.qmin;
	c.ref_bias = g.agilent_ref_bias; c.theta = g.theta;
	double alpha = 0.0;
	for (int j = 1; j <= ctx->K; j++) alpha += 1.0 / j;  // crispEM.c:223-224
	alpha *= g.theta;
	c.alpha_prior = alpha;
	c.rng_seed = g.rng_seed;
	c.ploidy = static_cast<const int32_t*>(ctx->d_ploidy.p);
	c.pclass = static_cast<const int32_t*>(ctx->d_pclass.p);
	c.NC = static_cast<const double*>(ctx->d_NC.p);
	c.T = static_cast<const double*>(ctx->d_T.p);
	c.ncr_tab = static_cast<const double*>(ctx->d_ncr.p);
	c.nq = ctx->nq;
	c.dense_lik = ctx->dense_lik;
	memcpy(c.qlist, ctx->qlist, sizeof c.qlist);
	bind_out(ctx);
	prm.o = ctx->dout;
	prm.q.perm_tasks = static_cast<int32_t*>(ctx->d_perm_tasks.p);
	prm.q.call_tasks = static_cast<int32_t*>(ctx->d_call_tasks.p);
	prm.q.call_tasks2 = static_cast<int32_t*>(ctx->d_call_tasks2.p);
	prm.q.counters = static_cast<int32_t*>(ctx->d_counters.p);
	if (!ctx->bins_direct) { prm.b.bins = nullptr; prm.b.bin_off = nullptr; }
	prm.q.need_reads = ctx->lazy_reads ? static_cast<uint8_t*>(ctx->d_need_reads.p) : nullptr;
	prm.gws = static_cast<double*>(ctx->d_gws.p);
	prm.gws_stride = ctx->gws_stride;
	return prm;
}

// Stage 1 (enqueued by submit): screening kernels up to the copy of the task counters.
// Stage 2 (enqueued by wait, after the counters arrived): caller kernels, finalisation, D2H.
// Splitting at the only host synchronisation keeps crispgpu_submit non-blocking, so a front end can keep a second
// context busy with the next batch's H2D while this one computes.
int run_stage1(crispgpu_ctx* ctx)
{
	const int n = ctx->n_sites;
	const crispgpu_config& g = ctx->cfg;
	int rc = ensure_outputs(ctx, n > 0 ? n : 1);
	if (rc) return rc;
	int64_t launches = 0;
	CK(cudaMemsetAsync(ctx->d_counters.p, 0, 16 * sizeof(int32_t), ctx->stream));
	const bool lazy = ctx->lazy_reads != nullptr && n > 0;
	const bool use_bins = ctx->binned && n > 0 && g.em_flag >= 1 && g.use_base_qvs == 1;
	if (!use_bins) ctx->binned = false;
	if (lazy) {
		rc = ensure_dev(ctx, ctx->d_need_reads, (size_t)n); if (rc) return rc;
		CK(cudaMemsetAsync(ctx->d_need_reads.p, 0, (size_t)n, ctx->stream));
	} else ctx->lazy_reads = nullptr;
	EngineParams prm = make_params(ctx);
	if (ctx->narrow && n > 0) {
		const int need = (n + 7) / 8;
		k_expand_counts16<<<need < ctx->num_sms * 8 ? need : ctx->num_sms * 8, 256, 0, ctx->stream>>>(static_cast<const uint8_t*>(ctx->nar[4]), static_cast<const uint32_t*>(ctx->nar[5]),
		                                                                                 static_cast<const uint16_t*>(ctx->nar[3]), static_cast<int32_t*>(ctx->d_b[6].p),
		                                                                                 ctx->narrow_counts ? static_cast<int32_t*>(ctx->d_b[5].p) : nullptr, n, ctx->P);
		launches++;
		CK(launch_check());
	}
	if (ctx->sparse_ic && n > 0) {
		const size_t cells = (size_t)n * ctx->P;
		k_expand_counts<<<ctx->num_sms * 8, 256, 0, ctx->stream>>>(ctx->db.ic_off, ctx->db.ic_sparse, static_cast<int32_t*>(ctx->d_b[6].p), cells);
		launches++;
		CK(launch_check());
	}
	if (ctx->device_sort && n > 0) {
		k_candidates<<<(n + 7) / 8, 256, 0, ctx->stream>>>(ctx->db.refbase, ctx->db.counts, ctx->db.indcounts, static_cast<const int32_t*>(ctx->d_ploidy.p), n, ctx->P,
		                                                   static_cast<uint8_t*>(ctx->d_b[3].p), static_cast<int32_t*>(ctx->d_b[4].p), nullptr);
		launches++;
		CK(launch_check());
	}
	CK(cudaEventRecord(ctx->ev[EV_H2D], ctx->stream));
	const bool scan = n > 0 && g.em_flag >= 1 && g.use_base_qvs == 1 && !lazy && ((use_bins && (ctx->n_bins > 0 || ctx->narrow)) || (!use_bins && ctx->n_reads > 0 && ctx->db.reads != nullptr));
	// 4-byte alt-read words -> records: only the caller kernels read them, so the side stream takes them when it is in use
	const bool alt32 = ctx->narrow && n > 0 && ctx->n_alt32 > 0;
	if (alt32 && !scan) {
		k_unpack_alt32<<<ctx->num_sms * 2, 256, 0, ctx->stream>>>(static_cast<const uint32_t*>(ctx->nar[6]), static_cast<crispgpu_altread*>(ctx->d_b[12].p), ctx->n_alt32);
		launches++;
		CK(launch_check());
	}
	if (scan) {
		// fork: the scan of the read stream (or of its compact bins) is independent of the count-based screening
		CK(cudaEventRecord(ctx->ev_fork, ctx->stream));
		CK(cudaStreamWaitEvent(ctx->side, ctx->ev_fork, 0));
		if (alt32) {
			k_unpack_alt32<<<ctx->num_sms * 2, 256, 0, ctx->side>>>(static_cast<const uint32_t*>(ctx->nar[6]), static_cast<crispgpu_altread*>(ctx->d_b[12].p), ctx->n_alt32);
			launches++;
		}
		if (use_bins && ctx->narrow) {
			const int need = (n + 7) / 8;
			k_narrow_offsets<<<need < ctx->num_sms * 8 ? need : ctx->num_sms * 8, 256, 0, ctx->side>>>(static_cast<const uint8_t*>(ctx->nar[1]), static_cast<const uint32_t*>(ctx->nar[2]),
			                                                                                static_cast<uint32_t*>(ctx->d_b[19].p), n, ctx->P);
			if (ctx->n_bins > 0) k_unpack_bins16<<<ctx->num_sms * 4, 256, 0, ctx->side>>>(static_cast<const uint16_t*>(ctx->nar[0]), static_cast<uint32_t*>(ctx->d_b[20].p), ctx->n_bins, static_cast<int32_t*>(ctx->d_counters.p));
			launches++;
		} else if (use_bins) k_scan_bins<<<ctx->num_sms * 2, 256, 0, ctx->side>>>(ctx->db.bins, ctx->n_bins, static_cast<int32_t*>(ctx->d_counters.p));
		else k_qualmap<<<ctx->num_sms * 2, 256, 0, ctx->side>>>(ctx->db.reads, ctx->n_reads, static_cast<int32_t*>(ctx->d_counters.p));
		launches++;
		CK(launch_check());
		CK(cudaEventRecord(ctx->ev_join, ctx->side));
	}
	if (n > 0) {
		rc = launch_evaluate(ctx, prm);
		if (rc) return rc;
		launches += 2;
	}
	CK(launch_check());
	if (scan) CK(cudaStreamWaitEvent(ctx->stream, ctx->ev_join, 0));
	CK(cudaEventRecord(ctx->ev[EV_EVAL], ctx->stream));
	if (g.chisq_permutation == 1 && n > 0) {
		if (ctx->pm_smem == 0) {
			snprintf(ctx->err, sizeof ctx->err, "permutation test: %d pools exceed the shared-memory layout", ctx->P);
			return CRISPGPU_ERR_UNSUPPORTED;
		}
		if (ctx->ploidy[0] > 2) {
			const bool many = ctx->ps_many;
			if (!ctx->pm_grid) {
				rc = many ? occupancy_grid(ctx, k_permute_stranded<true>, ctx->ps_threads, ctx->ps_smem, &ctx->pm_grid)
				          : occupancy_grid(ctx, k_permute_stranded<false>, ctx->ps_threads, ctx->ps_smem, &ctx->pm_grid);
				if (rc) return rc;
			}
			// head: first 1024 permutations of every task; tasks still undecided go to the chunked tail
			PermTail tail{};
			rc = ensure_dev(ctx, ctx->d_tail_task, (size_t)n * 5 * 4); if (rc) return rc;
			rc = ensure_dev(ctx, ctx->d_tail_head, (size_t)n * 5 * 8); if (rc) return rc;
			rc = ensure_dev(ctx, ctx->d_tail_total, (size_t)n * 5 * 8); if (rc) return rc;
			rc = ensure_dev(ctx, ctx->d_tail_counters, 4 * 4); if (rc) return rc;
			CK(cudaMemsetAsync(ctx->d_tail_counters.p, 0, 4 * 4, ctx->stream));
			tail.task = static_cast<int32_t*>(ctx->d_tail_task.p);
			tail.hits_head = static_cast<int32_t*>(ctx->d_tail_head.p);
			tail.total = static_cast<int32_t*>(ctx->d_tail_total.p);
			tail.counters = static_cast<int32_t*>(ctx->d_tail_counters.p);
			const int32_t* exact_done = nullptr;
			if (ctx->d_exact_row.p) {
				// small tables are decided exactly; the Monte-Carlo kernels skip what this pass marks done
				rc = ensure_dev(ctx, ctx->d_exact_done, (size_t)n * 5 * 4); if (rc) return rc;
				k_exact_2xk<<<ctx->num_sms * 4, 128, 0, ctx->stream>>>(prm, static_cast<const double*>(ctx->d_exact_row.p), static_cast<const double*>(ctx->d_exact_cl.p), static_cast<int32_t*>(ctx->d_exact_done.p));
				CK(launch_check());
				launches++;
				exact_done = static_cast<const int32_t*>(ctx->d_exact_done.p);
			}
			if (many) k_permute_stranded<true><<<ctx->pm_grid, ctx->ps_threads, ctx->ps_smem, ctx->stream>>>(prm, ctx->ps_l10cap, ctx->ps_area, tail, exact_done);
			else k_permute_stranded<false><<<ctx->pm_grid, ctx->ps_threads, ctx->ps_smem, ctx->stream>>>(prm, ctx->ps_l10cap, ctx->ps_area, tail, exact_done);
			CK(launch_check());
			if (g.max_iter > kPermHead) {
				// the number of tail tasks sizes the chunk records (the one extra host wait of the permutation mode)
				int32_t* hc2 = static_cast<int32_t*>(ctx->h_counters.p) + 12;
				CK(cudaMemcpyAsync(hc2, ctx->d_tail_counters.p, 4, cudaMemcpyDeviceToHost, ctx->stream));
				CK(cudaMemcpyAsync(hc2 + 1, static_cast<int32_t*>(ctx->d_counters.p) + 11, 4, cudaMemcpyDeviceToHost, ctx->stream));
				CK(cudaStreamSynchronize(ctx->stream));
				const int n_tail = hc2[0];
				if (n_tail > 0) {
					// no pool deeper than 255 reads per strand: 16-bit occupancy words, twice the permutations in flight
					const bool narrow = hc2[1] <= 255;
					int T = 0, grid = 0;
					size_t smem = 0;
					perm_tail_geometry(ctx, narrow ? 2 : 4, false, &T, &smem);
					int area = (int)((size_t)ctx->P * T * (narrow ? 2 : 4));
					if (many) { T = 256; area = ctx->ps_area; smem = ctx->ps_smem; }
					rc = many ? occupancy_grid(ctx, k_permute_stranded_tail<uint32_t, true>, T, smem, &grid)
					   : narrow ? occupancy_grid(ctx, k_permute_stranded_tail<uint16_t, false>, T, smem, &grid) : occupancy_grid(ctx, k_permute_stranded_tail<uint32_t, false>, T, smem, &grid);
					if (rc) return rc;
					const int rem = g.max_iter - kPermHead;
					int chunk = ((rem + 255) / 256 + T - 1) / T * T;
					if (chunk < 4 * T) chunk = 4 * T;
					tail.chunk = chunk;
					tail.nchunks = (rem + chunk - 1) / chunk;
					rc = ensure_dev(ctx, ctx->d_tail_rec, (size_t)n_tail * tail.nchunks * kPermRec * 4); if (rc) return rc;
					tail.rec = static_cast<int32_t*>(ctx->d_tail_rec.p);
					if (many) k_permute_stranded_tail<uint32_t, true><<<grid, T, smem, ctx->stream>>>(prm, ctx->ps_l10cap, area, tail);
					else if (narrow) k_permute_stranded_tail<uint16_t, false><<<grid, T, smem, ctx->stream>>>(prm, ctx->pm_l10cap, area, tail);
					else k_permute_stranded_tail<uint32_t, false><<<grid, T, smem, ctx->stream>>>(prm, ctx->pm_l10cap, area, tail);
					CK(launch_check());
					k_permute_stranded_resolve<<<(n_tail + 127) / 128, 128, 0, ctx->stream>>>(prm, tail);
					launches += 2;
				}
			}
		} else {
			const size_t low_smem = ctx->pm_smem + (size_t)((ctx->P + 31) / 32) * ctx->pm_threads * 4;  // + the touched-pool bitmaps
			if (low_smem > ctx->smem_optin) {
				snprintf(ctx->err, sizeof ctx->err, "low-coverage permutation test: %d pools exceed the shared-memory layout", ctx->P);
				return CRISPGPU_ERR_UNSUPPORTED;
			}
			if (!ctx->pm_grid) {
				rc = occupancy_grid(ctx, k_permute_lowcov, ctx->pm_threads, low_smem, &ctx->pm_grid); if (rc) return rc;
			}
			PermTail tail{};
			rc = ensure_dev(ctx, ctx->d_tail_task, (size_t)n * 5 * 4); if (rc) return rc;
			rc = ensure_dev(ctx, ctx->d_tail_head, (size_t)n * 5 * 8); if (rc) return rc;
			rc = ensure_dev(ctx, ctx->d_tail_total, (size_t)n * 5 * 8); if (rc) return rc;
			rc = ensure_dev(ctx, ctx->d_tail_counters, 4 * 4); if (rc) return rc;
			CK(cudaMemsetAsync(ctx->d_tail_counters.p, 0, 4 * 4, ctx->stream));
			tail.task = static_cast<int32_t*>(ctx->d_tail_task.p);
			tail.hits_head = static_cast<int32_t*>(ctx->d_tail_head.p);
			tail.total = static_cast<int32_t*>(ctx->d_tail_total.p);
			tail.counters = static_cast<int32_t*>(ctx->d_tail_counters.p);
			k_permute_lowcov<<<ctx->pm_grid, ctx->pm_threads, low_smem, ctx->stream>>>(prm, tail);
			CK(launch_check());
			if (g.max_iter > kPermHead) {
				int32_t* hc2 = static_cast<int32_t*>(ctx->h_counters.p) + 12;
				CK(cudaMemcpyAsync(hc2, ctx->d_tail_counters.p, 4, cudaMemcpyDeviceToHost, ctx->stream));
				CK(cudaMemcpyAsync(hc2 + 1, static_cast<int32_t*>(ctx->d_counters.p) + 11, 4, cudaMemcpyDeviceToHost, ctx->stream));
				CK(cudaStreamSynchronize(ctx->stream));
				const int n_tail = hc2[0];
				if (n_tail > 0) {
					// no pool deeper than 31 reads per stratum: three 5-bit counters in 16-bit words
					const bool narrow = hc2[1] <= 31;
					int T = 0, grid = 0;
					size_t smem = 0;
					perm_tail_geometry(ctx, narrow ? 2 : 4, true, &T, &smem);
					rc = narrow ? occupancy_grid(ctx, k_permute_lowcov_tail<uint16_t>, T, smem, &grid) : occupancy_grid(ctx, k_permute_lowcov_tail<uint32_t>, T, smem, &grid);
					if (rc) return rc;
					const int rem = g.max_iter - kPermHead;
					int chunk = ((rem + 255) / 256 + T - 1) / T * T;
					if (chunk < 4 * T) chunk = 4 * T;
					tail.chunk = chunk;
					tail.nchunks = (rem + chunk - 1) / chunk;
					rc = ensure_dev(ctx, ctx->d_tail_rec, (size_t)n_tail * tail.nchunks * kLowRec * 4); if (rc) return rc;
					tail.rec = static_cast<int32_t*>(ctx->d_tail_rec.p);
					if (narrow) k_permute_lowcov_tail<uint16_t><<<grid, T, smem, ctx->stream>>>(prm, tail);
					else k_permute_lowcov_tail<uint32_t><<<grid, T, smem, ctx->stream>>>(prm, tail);
					CK(launch_check());
					k_permute_lowcov_resolve<<<(n_tail + 127) / 128, 128, 0, ctx->stream>>>(prm, tail);
					launches += 2;
				}
			}
		}
		launches++;
		CK(launch_check());
		CK(cudaEventRecord(ctx->ev[EV_PERM], ctx->stream));
		k_filter<<<ctx->num_sms * 2, 256, 0, ctx->stream>>>(prm);
		launches++;
		CK(launch_check());
	} else {
		CK(cudaEventRecord(ctx->ev[EV_PERM], ctx->stream));
	}
	if (lazy) {
		// pull the reads of the sites that reached the quality-value likelihood straight from the caller's pinned array
		k_fetch_reads<<<ctx->num_sms * 4, 256, 0, ctx->stream>>>(ctx->lazy_reads, static_cast<uint16_t*>(ctx->d_b[10].p), ctx->db.read_off,
		                                                        static_cast<const uint8_t*>(ctx->d_need_reads.p), n, ctx->P, ctx->n_reads, static_cast<int32_t*>(ctx->d_counters.p));
		launches++;
		CK(launch_check());
	}
	CK(cudaEventRecord(ctx->ev[EV_FILTER], ctx->stream));
	// how many tasks reached the caller: sizes the per-pool rows
	CK(cudaMemcpyAsync(ctx->h_counters.p, ctx->d_counters.p, 16 * sizeof(int32_t), cudaMemcpyDeviceToHost, ctx->stream));
	CK(cudaEventRecord(ctx->ev_counts, ctx->stream));
	ctx->stage1_launches = launches;
	return 0;
}

int run_stage2(crispgpu_ctx* ctx, bool fetch)
{
	const int n = ctx->n_sites;
	const crispgpu_config& g = ctx->cfg;
	int rc = 0;
	int64_t launches = ctx->stage1_launches;
	EngineParams prm = make_params(ctx);
	CK(cudaEventSynchronize(ctx->ev_counts));
	const int32_t* hc = static_cast<const int32_t*>(ctx->h_counters.p);
	if (g.chisq_permutation == 1 && n > 0) {
		// the occupancy counters of the permutation kernels are packed fields: 16 bits per strand (pool size > 2), 10 bits per
		// stratum (low-coverage test).  A pool that could receive more balls than a field holds would corrupt its neighbour.
		const int cap = ctx->ploidy[0] > 2 ? 65535 : 1023;
		if (hc[13] > cap) {
			snprintf(ctx->err, sizeof ctx->err, "permutation test: a pool can receive %d alternate reads of one %s in a permutation, more than the %d the packed occupancy "
			         "counters hold (deeper data needs a wider counter layout)", hc[13], ctx->ploidy[0] > 2 ? "strand" : "stratum", cap);
			return CRISPGPU_ERR_UNSUPPORTED;
		}
	}
	ctx->n_perm = hc[0];
	ctx->n_call1 = hc[1];
	ctx->n_call2 = hc[8];
	ctx->n_call = hc[1] + hc[8];
	ctx->timing.fetched_bytes = ctx->lazy_reads ? 2 * (int64_t)(static_cast<uint32_t>(hc[14]) | (static_cast<uint64_t>(static_cast<uint32_t>(hc[15])) << 32)) : 0;
	ctx->timing.h2d_bytes = ctx->copied_bytes + ctx->timing.fetched_bytes;
	{
		// distinct quality characters of the batch -> dense likelihood path when few enough and one pool size
		int nq = 0;
		uint8_t ql[16] = {0};
		for (int q = 0; q < kNQ; q++)
			if ((static_cast<const uint32_t*>(ctx->h_counters.p)[4 + (q >> 5)] >> (q & 31)) & 1u) { if (nq < 16) ql[nq] = (uint8_t)q; nq++; }
		const size_t planes = ctx->em_smem_g ? (size_t)3 * (ctx->nmax + 1) * ctx->P * 8 : 0;  // planes in shared memory or in the global workspace
		const bool dense = g.em_flag >= 1 && g.use_base_qvs == 1 && ctx->class_ploidy.size() == 1 && nq >= 1 && nq <= kDenseNQ &&
		                   em_smem_fixed(ctx->P, ctx->nmax + 1, ctx->em_warps, true, nq) + planes <= ctx->smem_optin;
		ctx->nq = dense ? nq : 0;
		ctx->dense_lik = dense ? 1 : 0;
		ctx->bins_direct = ctx->binned;  // both likelihood paths of k_em take the (value, count) bins as they are
		memcpy(ctx->qlist, ql, sizeof ql);
		// two-plane variant: no bidirectional read in the batch, one pool size, no association pass
		const bool scanned = g.em_flag >= 1 && g.use_base_qvs == 1 && (ctx->n_reads > 0 || (ctx->binned && ctx->n_bins > 0));
		const bool nobidir = scanned && hc[10] == 0 && ctx->class_ploidy.size() == 1 && !(g.association == 1 && ctx->d_pheno.p);
		ctx->nobidir = nobidir ? 1 : 0;
		const int key = (dense ? nq : 0) + (nobidir ? 64 : 0) + (ctx->bins_direct ? 128 : 0);
		if (key != ctx->em_key) {
			ctx->em_key = key;
			ctx->em_grid = 0;
			ctx->em_grid2 = 0;
			const int J = ctx->nmax + 1;
			const size_t fixed = em_smem_fixed(ctx->P, J, ctx->em_warps, dense, nq);
			ctx->em_smem = ctx->em_smem_g ? fixed + (size_t)3 * J * ctx->P * 8 : fixed;
			const size_t fixed2 = em_smem_fixed(ctx->P, J, ctx->em_warps, dense, nq, true);
			// planes in shared memory only when at least two CTAs still fit per SM; a single resident CTA hides less latency
			// than several CTAs working out of the L2-resident workspace (measured on config 3: 1.11 ms vs 0.98 ms)
			ctx->em_smem_g2 = 2 * (fixed2 + (size_t)2 * J * ctx->P * 8 + 1024) <= ctx->smem_optin;
			ctx->em_smem2 = ctx->em_smem_g2 ? fixed2 + (size_t)2 * J * ctx->P * 8 : fixed2;
		}
		prm = make_params(ctx);
	}
	const int rows = g.em_flag >= 1 ? ctx->n_call : 0;
	if (rows > 0) {
		// The per-pool rows (allele counts, genotype qualities) are indexed by EM task, but only called alleles have one:
		// when the results go to the host k_em writes them straight into the pinned host buffer (mapped, same address under
		// unified addressing), so only the rows of called alleles cross the link — a third of the tasks at the benchmark
		// shape — and no copy is enqueued for them.  Device-resident runs keep them in device memory.
		if (fetch) {
			rc = ensure_host(ctx, ctx->h_rows, (size_t)rows * ctx->P * 12); if (rc) return rc;
			ctx->h_qv.p = ctx->h_rows.p;
			ctx->h_ac.p = static_cast<unsigned char*>(ctx->h_rows.p) + (size_t)rows * ctx->P * 8;
			ctx->d_qv.p = ctx->h_qv.p;
			ctx->d_ac.p = ctx->h_ac.p;
		} else {
			rc = ensure_dev(ctx, ctx->d_rows, (size_t)rows * ctx->P * 12); if (rc) return rc;
			ctx->d_qv.p = ctx->d_rows.p;
			ctx->d_ac.p = static_cast<unsigned char*>(ctx->d_rows.p) + (size_t)rows * ctx->P * 8;
		}
		prm = make_params(ctx);
	}
	if (ctx->n_call > 0) {
		if (g.em_flag >= 1) {
			const bool assoc = g.association == 1 && ctx->d_pheno.p != nullptr;
			int32_t* cnt = static_cast<int32_t*>(ctx->d_counters.p);
			static const bool keep_order = getenv("CRISPGPU_EM_SITE_ORDER") != nullptr;  // measurement switch
			static const char* trace_path = getenv("CRISPGPU_EM_TRACE");                 // measurement switch: task timeline of the last launch
			if (trace_path) {
				rc = ensure_dev(ctx, ctx->d_trace, (size_t)ctx->n_call * 8 * sizeof(long long)); if (rc) return rc;
				CK(cudaMemsetAsync(ctx->d_trace.p, 0, (size_t)ctx->n_call * 8 * sizeof(long long), ctx->stream));
				prm.em_trace = static_cast<long long*>(ctx->d_trace.p);
			}
			auto launch_em = [&](auto kernel, const int32_t* list, int ntasks, int32_t* head, int row0, size_t smem, bool smem_g, int* grid_cache) -> int {
				if (ntasks <= 0) return 0;
				if (!*grid_cache) {
					int r2 = occupancy_grid(ctx, kernel, ctx->em_threads, smem, grid_cache);
					if (r2) return r2;
				} else {
					int r2 = prepare_kernel(ctx, kernel, smem);  // another instantiation may have filled the shared grid cache
					if (r2) return r2;
				}
				int grid = *grid_cache < ntasks ? *grid_cache : ntasks;
				if (ntasks > *grid_cache && !keep_order) {
					// more tasks than resident CTAs: slow tasks first (k_order_tasks), so the launch does not end on one
					int32_t* ordered = static_cast<int32_t*>(ctx->d_order.p) + row0;
					k_order_tasks<<<1, 1024, 0, ctx->stream>>>(list, ntasks, prm.o.paf, ordered);
					launches++;
					list = ordered;
				}
				if (!smem_g) {
					int r2 = ensure_dev(ctx, ctx->d_gws, (size_t)*grid_cache * ctx->gws_stride * sizeof(double));
					if (r2) return r2;
					prm = make_params(ctx);
				}
				kernel<<<grid, ctx->em_threads, smem, ctx->stream>>>(prm, list, ntasks, head, row0);
				launches++;
				return 0;
			};
			const int32_t* l1 = prm.q.call_tasks;
			const int32_t* l2 = prm.q.call_tasks2;
			const size_t sm3 = ctx->em_smem, sm2 = ctx->em_smem2;
			const bool g3 = ctx->em_smem_g, g2 = ctx->em_smem_g2;
			// quality-value likelihood tasks (two-plane kernel when the batch allows it) ...
			if (ctx->nobidir) {
				static const size_t pad = getenv("CRISPGPU_EM_PAD_SMEM") ? (size_t)atoi(getenv("CRISPGPU_EM_PAD_SMEM")) : 0;  // occupancy experiment
				static const bool generic = getenv("CRISPGPU_EM_GENERIC") != nullptr;  // measurement switch
				if (g2 && ctx->dense_lik && ctx->bins_direct && !generic) rc = launch_em(k_em<true, false, false, true, 1>, l1, ctx->n_call1, cnt + 3, 0, sm2 + pad, true, &ctx->em_grid2);
				else
				rc = g2 ? launch_em(k_em<true, false, false, true>, l1, ctx->n_call1, cnt + 3, 0, sm2, true, &ctx->em_grid2)
				        : launch_em(k_em<false, false, false, true>, l1, ctx->n_call1, cnt + 3, 0, sm2, false, &ctx->em_grid2);
			} else if (g3) {
				rc = assoc ? launch_em(k_em<true, true, false, false>, l1, ctx->n_call1, cnt + 3, 0, sm3, true, &ctx->em_grid)
				           : launch_em(k_em<true, false, false, false>, l1, ctx->n_call1, cnt + 3, 0, sm3, true, &ctx->em_grid);
			} else {
				rc = assoc ? launch_em(k_em<false, true, false, false>, l1, ctx->n_call1, cnt + 3, 0, sm3, false, &ctx->em_grid)
				           : launch_em(k_em<false, false, false, false>, l1, ctx->n_call1, cnt + 3, 0, sm3, false, &ctx->em_grid);
			}
			// ... then count-based likelihood tasks (always three planes; their AC/QV rows follow the first list's)
			if (!rc) {
				if (g3) rc = assoc ? launch_em(k_em<true, true, true, false>, l2, ctx->n_call2, cnt + 9, ctx->n_call1, sm3, true, &ctx->em_grid)
				                   : launch_em(k_em<true, false, true, false>, l2, ctx->n_call2, cnt + 9, ctx->n_call1, sm3, true, &ctx->em_grid);
				else rc = assoc ? launch_em(k_em<false, true, true, false>, l2, ctx->n_call2, cnt + 9, ctx->n_call1, sm3, false, &ctx->em_grid)
				                : launch_em(k_em<false, false, true, false>, l2, ctx->n_call2, cnt + 9, ctx->n_call1, sm3, false, &ctx->em_grid);
			}
			if (rc) return rc;
			launches--;  // the common increment below counts one
		} else {
			if (!ctx->lg_grid) { rc = occupancy_grid(ctx, k_legacy<4>, 128, ctx->lg_smem, &ctx->lg_grid); if (rc) return rc; }
			int need = (ctx->n_call + 3) / 4;
			k_legacy<4><<<ctx->lg_grid < need ? ctx->lg_grid : need, 128, ctx->lg_smem, ctx->stream>>>(prm);
		}
		launches++;
		CK(launch_check());
	}
	if (prm.em_trace) {
		std::vector<long long> tr((size_t)ctx->n_call * 8);
		CK(cudaStreamSynchronize(ctx->stream));
		CK(cudaMemcpy(tr.data(), ctx->d_trace.p, tr.size() * sizeof(long long), cudaMemcpyDeviceToHost));
		if (FILE* f = fopen(getenv("CRISPGPU_EM_TRACE"), "wb")) { fwrite(tr.data(), sizeof(long long), tr.size(), f); fclose(f); }
	}
	if (n > 0) {
		k_finalize<<<(n + 255) / 256, 256, 0, ctx->stream>>>(prm);
		launches++;
		CK(launch_check());
	}
	CK(cudaEventRecord(ctx->ev[EV_CALL], ctx->stream));
	int64_t d2h = 0;
	if (fetch) {
		if (n > 0) {
			CK(cudaMemcpyAsync(ctx->h_out_arena.p, ctx->d_out_arena.p, ctx->out_bytes, cudaMemcpyDeviceToHost, ctx->stream));
			d2h += (int64_t)ctx->out_bytes;
		}
		ctx->rows_on_host = rows > 0;  // their bytes are counted once the number of called alleles is known (crispgpu_wait)
	}
	CK(cudaEventRecord(ctx->ev[EV_D2H], ctx->stream));
	if (ctx->ev_done) CK(cudaEventRecord(ctx->ev_done, ctx->stream));
	ctx->timing.launches = launches;
	ctx->timing.d2h_bytes = d2h;
	ctx->timing.n_tasks_em = ctx->n_call;
	ctx->timing.n_tasks_perm = ctx->n_perm;
	return 0;
}

void fill_results(crispgpu_ctx* ctx, crispgpu_results* out)
{
	crispgpu_results& r = ctx->res;
	r.n_sites = ctx->n_sites;
	r.n_calls = ctx->cfg.em_flag >= 1 ? ctx->n_call : 0;
	const void** hp = reinterpret_cast<const void**>(&r.varalleles);
	for (int f = 0; f < kNumOut; f++) hp[f] = ctx->h_o[f].p;
	r.ac = static_cast<const int32_t*>(ctx->h_ac.p);
	r.qv = static_cast<const double*>(ctx->h_qv.p);
	if (out) *out = r;
}

int finish_timing(crispgpu_ctx* ctx)
{
	crispgpu_timing& t = ctx->timing;
	if (ctx->pile_timing) {
		CK(cudaEventElapsedTime(&t.h2d_ms, ctx->ev[EV_START], ctx->ev_pile[0]));
		CK(cudaEventElapsedTime(&t.pileup_ms, ctx->ev_pile[0], ctx->ev_pile[3]));
		CK(cudaEventElapsedTime(&t.pile_tile_ms, ctx->ev_pile[1], ctx->ev_pile[2]));
	} else {
		CK(cudaEventElapsedTime(&t.h2d_ms, ctx->ev[EV_START], ctx->ev[EV_H2D]));
		t.pileup_ms = t.pile_tile_ms = 0.f;
	}
	CK(cudaEventElapsedTime(&t.evaluate_ms, ctx->ev[EV_H2D], ctx->ev[EV_EVAL]));
	CK(cudaEventElapsedTime(&t.permute_ms, ctx->ev[EV_EVAL], ctx->ev[EV_PERM]));
	CK(cudaEventElapsedTime(&t.filter_ms, ctx->ev[EV_PERM], ctx->ev[EV_FILTER]));
	float call_ms = 0;
	CK(cudaEventElapsedTime(&call_ms, ctx->ev[EV_FILTER], ctx->ev[EV_CALL]));
	t.em_ms = ctx->cfg.em_flag >= 1 ? call_ms : 0.f;
	t.legacy_ms = ctx->cfg.em_flag >= 1 ? 0.f : call_ms;
	CK(cudaEventElapsedTime(&t.d2h_ms, ctx->ev[EV_CALL], ctx->ev[EV_D2H]));
	CK(cudaEventElapsedTime(&t.total_ms, ctx->ev[EV_START], ctx->ev[EV_D2H]));
	return 0;
}

__global__ void k_dfma_peak(double* out, int iters)
{
	double a0 = threadIdx.x * 1e-9, a1 = a0 + 1, a2 = a0 + 2, a3 = a0 + 3, a4 = a0 + 4, a5 = a0 + 5, a6 = a0 + 6, a7 = a0 + 7;
	const double m = 1.0000001, c = 1e-7;
	for (int i = 0; i < iters; i++) {
		a0 = fma(a0, m, c); a1 = fma(a1, m, c); a2 = fma(a2, m, c); a3 = fma(a3, m, c);
		a4 = fma(a4, m, c); a5 = fma(a5, m, c); a6 = fma(a6, m, c); a7 = fma(a7, m, c);
	}
	if (a0 + a1 + a2 + a3 + a4 + a5 + a6 + a7 == 12345.678) out[0] = a0;
}

// 32-bit integer multiply-add chains (IMAD, the instruction the permutation kernels' Philox rounds and index arithmetic are
// made of): eight independent chains per thread, register-resident
__global__ void k_imad_peak(uint32_t* out, int iters)
{
	uint32_t a0 = threadIdx.x, a1 = a0 + 1, a2 = a0 + 2, a3 = a0 + 3, a4 = a0 + 4, a5 = a0 + 5, a6 = a0 + 6, a7 = a0 + 7;
	const uint32_t m = 0xD2511F53u + blockIdx.x, c = 0x9E3779B9u;
	for (int i = 0; i < iters; i++) {
		a0 = a0 * m + c; a1 = a1 * m + c; a2 = a2 * m + c; a3 = a3 * m + c;
		a4 = a4 * m + c; a5 = a5 * m + c; a6 = a6 * m + c; a7 = a7 * m + c;
	}
	if ((a0 ^ a1 ^ a2 ^ a3 ^ a4 ^ a5 ^ a6 ^ a7) == 0x12345678u) out[0] = a0;
}

}  // namespace

extern "C" {

int crispgpu_measure_int32_tops(int device_id, double* tops)
{
	if (!tops) return CRISPGPU_ERR_ARG;
	cudaDeviceProp prop;
	if (cudaSetDevice(device_id) != cudaSuccess || cudaGetDeviceProperties(&prop, device_id) != cudaSuccess) return CRISPGPU_ERR_NO_DEVICE;
	uint32_t* d = nullptr;
	cudaEvent_t e0, e1;
	if (cudaMalloc(&d, 8) != cudaSuccess) return CRISPGPU_ERR_CUDA;
	cudaEventCreate(&e0);
	cudaEventCreate(&e1);
	const int blocks = prop.multiProcessorCount * 8, threads = 256, iters = 1 << 15;
	double best = 0.0;
	for (int rep = 0; rep < 4; rep++) {
		cudaEventRecord(e0);
		k_imad_peak<<<blocks, threads>>>(d, iters);
		cudaEventRecord(e1);
		if (cudaEventSynchronize(e1) != cudaSuccess) { cudaFree(d); return CRISPGPU_ERR_CUDA; }
		float ms = 0;
		cudaEventElapsedTime(&ms, e0, e1);
		double t = 8.0 * (double)iters * blocks * threads / (ms * 1e-3) / 1e12;   // integer multiply-add instructions per lane
		if (rep > 0 && t > best) best = t;
	}
	cudaEventDestroy(e0);
	cudaEventDestroy(e1);
	cudaFree(d);
	*tops = best;
	return CRISPGPU_OK;
}

int crispgpu_measure_fp64_tflops(int device_id, double* tflops)
{
	if (!tflops) return CRISPGPU_ERR_ARG;
	cudaDeviceProp prop;
	if (cudaSetDevice(device_id) != cudaSuccess || cudaGetDeviceProperties(&prop, device_id) != cudaSuccess) return CRISPGPU_ERR_NO_DEVICE;
	double* d = nullptr;
	cudaEvent_t e0, e1;
	if (cudaMalloc(&d, 8) != cudaSuccess) return CRISPGPU_ERR_CUDA;
	cudaEventCreate(&e0);
	cudaEventCreate(&e1);
	const int blocks = prop.multiProcessorCount * 8, threads = 256, iters = 1 << 15;
	double best = 0.0;
	for (int rep = 0; rep < 4; rep++) {
		cudaEventRecord(e0);
		k_dfma_peak<<<blocks, threads>>>(d, iters);
		cudaEventRecord(e1);
		if (cudaEventSynchronize(e1) != cudaSuccess) { cudaFree(d); return CRISPGPU_ERR_CUDA; }
		float ms = 0;
		cudaEventElapsedTime(&ms, e0, e1);
		double tf = 2.0 * 8.0 * (double)iters * blocks * threads / (ms * 1e-3) / 1e12;
		if (rep > 0 && tf > best) best = tf;
	}
	cudaEventDestroy(e0);
	cudaEventDestroy(e1);
	cudaFree(d);
	*tflops = best;
	return CRISPGPU_OK;
}

void crispgpu_config_defaults(crispgpu_config* cfg)
{
	if (!cfg) return;
	memset(cfg, 0, sizeof(*cfg));
	cfg->abi_version = CRISPGPU_ABI_VERSION;
	cfg->em_flag = 1;
	cfg->max_iter = 20000;
	cfg->chisq_permutation = 0;
	cfg->fast_filter = 1;
	cfg->use_base_qvs = 1;
	cfg->min_reads = 4;
	cfg->qv_offset = 33;
	cfg->min_q = 13;
	cfg->thresh1 = -3.5;
	cfg->min_qpv = -5;
	cfg->thresh3 = -2;
	cfg->ser = -1;
	cfg->relax = 0.75;
	cfg->qmin = 23;
	cfg->agilent_ref_bias = 0.5;
	cfg->theta = 0.001;
	cfg->rng_seed = 1;
}

int crispgpu_abi_version(void) { return CRISPGPU_ABI_VERSION; }

int crispgpu_device_count(void)
{
	int n = 0;
	if (cudaGetDeviceCount(&n) != cudaSuccess) return 0;
	return n;
}

int crispgpu_create(const crispgpu_config* cfg, int device_id, void* stream, crispgpu_ctx** out)
{
	if (!cfg || !out || cfg->abi_version != CRISPGPU_ABI_VERSION || cfg->n_pools <= 0 || !cfg->ploidy) return CRISPGPU_ERR_ARG;
	*out = nullptr;
	int ndev = 0;
	if (cudaGetDeviceCount(&ndev) != cudaSuccess || ndev <= 0 || device_id < 0 || device_id >= ndev) return CRISPGPU_ERR_NO_DEVICE;
	crispgpu_ctx* ctx = new (std::nothrow) crispgpu_ctx();
	if (!ctx) return CRISPGPU_ERR_NOMEM;
	auto fail = [&](int rc) { crispgpu_destroy(ctx); return rc; };
	ctx->device = device_id;
	if (cudaSetDevice(device_id) != cudaSuccess) return fail(CRISPGPU_ERR_NO_DEVICE);
	cudaDeviceProp prop;
	if (cudaGetDeviceProperties(&prop, device_id) != cudaSuccess) return fail(CRISPGPU_ERR_NO_DEVICE);
	if (prop.major < 10) return fail(CRISPGPU_ERR_NO_DEVICE);  // built for sm_100a only
	ctx->num_sms = prop.multiProcessorCount;
	ctx->smem_optin = prop.sharedMemPerBlockOptin;
	ctx->cfg = *cfg;
	ctx->P = cfg->n_pools;
	ctx->ploidy.assign(cfg->ploidy, cfg->ploidy + ctx->P);
	ctx->cfg.ploidy = ctx->ploidy.data();
	if (cfg->phenotypes) ctx->phenotypes.assign(cfg->phenotypes, cfg->phenotypes + ctx->P);
	ctx->cfg.phenotypes = ctx->phenotypes.empty() ? nullptr : ctx->phenotypes.data();
	ctx->pclass.resize(ctx->P);
	for (int i = 0; i < ctx->P; i++) {
		const int n = ctx->ploidy[i];
		if (n < 1) return fail(CRISPGPU_ERR_ARG);
		ctx->K += n;
		if (n > ctx->nmax) ctx->nmax = n;
		int pc = -1;
		for (size_t k = 0; k < ctx->class_ploidy.size(); k++) if (ctx->class_ploidy[k] == n) pc = (int)k;
		if (pc < 0) { pc = (int)ctx->class_ploidy.size(); ctx->class_ploidy.push_back(n); }
		ctx->pclass[i] = pc;
	}
	if (ctx->class_ploidy.size() > 64) return fail(CRISPGPU_ERR_UNSUPPORTED);
	if (setup_geometry(ctx) != 0) return fail(CRISPGPU_ERR_UNSUPPORTED);
	if (stream) ctx->stream = static_cast<cudaStream_t>(stream);
	else {
		if (cudaStreamCreateWithFlags(&ctx->stream, cudaStreamNonBlocking) != cudaSuccess) return fail(CRISPGPU_ERR_CUDA);
		ctx->own_stream = true;
	}
	for (int i = 0; i < EV_COUNT; i++) if (cudaEventCreate(&ctx->ev[i]) != cudaSuccess) return fail(CRISPGPU_ERR_CUDA);
	if (cudaEventCreateWithFlags(&ctx->ev_counts, cudaEventDisableTiming | (cfg->blocking_wait == 1 ? cudaEventBlockingSync : 0)) != cudaSuccess) return fail(CRISPGPU_ERR_CUDA);
	if (cfg->blocking_wait == 1 && cudaEventCreateWithFlags(&ctx->ev_done, cudaEventBlockingSync | cudaEventDisableTiming) != cudaSuccess) return fail(CRISPGPU_ERR_CUDA);
	if (cudaEventCreate(&ctx->ev_epoch) != cudaSuccess || cudaEventRecord(ctx->ev_epoch, ctx->stream) != cudaSuccess) return fail(CRISPGPU_ERR_CUDA);
	if (cudaEventCreateWithFlags(&ctx->ev_fork, cudaEventDisableTiming) != cudaSuccess) return fail(CRISPGPU_ERR_CUDA);
	if (cudaEventCreateWithFlags(&ctx->ev_join, cudaEventDisableTiming) != cudaSuccess) return fail(CRISPGPU_ERR_CUDA);
	if (cudaStreamCreateWithFlags(&ctx->side, cudaStreamNonBlocking) != cudaSuccess) return fail(CRISPGPU_ERR_CUDA);
	// configuration tables: NC on the host exactly as init_variant does (init_variant.c:19-25), T on the device
	const int J = ctx->nmax + 1, ncl = (int)ctx->class_ploidy.size();
	std::vector<double> NC((size_t)ncl * J, 0.0);
	for (int k = 0; k < ncl; k++) {
		const int n = ctx->class_ploidy[k];
		double* row = NC.data() + (size_t)k * J;
		for (int j = 1; j < n; j++) row[j] = row[j - 1] + log10((double)(n - j + 1)) - log10((double)j);
	}
	auto up = [&](DevBuf& b, const void* src, size_t bytes) -> int {
		int rc = ensure_dev(ctx, b, bytes);
		if (rc) return rc;
		if (cudaMemcpyAsync(b.p, src, bytes, cudaMemcpyHostToDevice, ctx->stream) != cudaSuccess) return CRISPGPU_ERR_CUDA;
		return 0;
	};
	int rc = 0;
	if ((rc = up(ctx->d_ploidy, ctx->ploidy.data(), (size_t)ctx->P * 4))) return fail(rc);
	if ((rc = up(ctx->d_pclass, ctx->pclass.data(), (size_t)ctx->P * 4))) return fail(rc);
	if ((rc = up(ctx->d_class_ploidy, ctx->class_ploidy.data(), (size_t)ncl * 4))) return fail(rc);
	if ((rc = up(ctx->d_NC, NC.data(), NC.size() * 8))) return fail(rc);
	if (!ctx->phenotypes.empty() && (rc = up(ctx->d_pheno, ctx->phenotypes.data(), (size_t)ctx->P * 4))) return fail(rc);
	if ((rc = ensure_dev(ctx, ctx->d_T, (size_t)ncl * 2 * kNQ * J * 8))) return fail(rc);
	if (cfg->chisq_permutation == 1 && ctx->ploidy[0] == 2) {
		// constant table log10 C(n,r) in the reference's own summation order, evaluated with the host libm
		const int H = kNcrTabN / 2 + 1;
		std::vector<double> tab((size_t)kNcrTabN * H, 0.0);
		for (int n = 1; n < kNcrTabN; n++) {
			double ll = 0.0;
			for (int r = 1; r <= n / 2; r++) {
				ll += log10((double)(n - (r - 1)) / (double)r);
				tab[(size_t)n * H + r] = ll;
			}
		}
		if ((rc = up(ctx->d_ncr, tab.data(), tab.size() * 8))) return fail(rc);
	}
	if (cfg->chisq_permutation == 1 && cfg->exact_mode == 1 && ctx->ploidy[0] > 2 && ctx->P >= 2 && ctx->P < 8) {
		// term tables of the exact test, host libm, the reference's two formulas: the running sums of
		// compute_NCRtable (contables.c:244-257) for the enumerated tables, ncr() (contables.c:80-87) for the observed one
		std::vector<double> rowt((size_t)kExactMaxN * kExactOnes, 0.0), clt((size_t)kExactMaxN * kExactOnes, 0.0);
		for (int n = 1; n < kExactMaxN; n++) {
			double value = 0.0;
			for (int j = 1; j < kExactOnes; j++) {
				if (j < n) { value += log10((double)(n - j + 1)) - log10((double)j); rowt[(size_t)n * kExactOnes + j] = value; }
				if (j <= n) {
					int r = j;
					double ll = 0.0;
					if (n != r) {
						if (2 * r > n) r = n - r;
						for (int i = 0; i < r; i++) ll += log10((double)(n - i) / (double)(i + 1));
					}
					clt[(size_t)n * kExactOnes + j] = ll;
				}
			}
		}
		if ((rc = up(ctx->d_exact_row, rowt.data(), rowt.size() * 8))) return fail(rc);
		if ((rc = up(ctx->d_exact_cl, clt.data(), clt.size() * 8))) return fail(rc);
	}
	if ((rc = ensure_dev(ctx, ctx->d_counters, 16 * 4))) return fail(rc);
	if ((rc = ensure_host(ctx, ctx->h_counters, 16 * 4))) return fail(rc);
	k_build_T<<<ctx->num_sms, 256, 0, ctx->stream>>>(static_cast<double*>(ctx->d_T.p), static_cast<const int32_t*>(ctx->d_class_ploidy.p), ncl, J, cfg->qv_offset, cfg->agilent_ref_bias);
	if (cudaGetLastError() != cudaSuccess || cudaStreamSynchronize(ctx->stream) != cudaSuccess) return fail(CRISPGPU_ERR_CUDA);
	*out = ctx;
	return CRISPGPU_OK;
}

void crispgpu_destroy(crispgpu_ctx* ctx)
{
	if (!ctx) return;
	cudaSetDevice(ctx->device);
	if (ctx->stream) cudaStreamSynchronize(ctx->stream);
	auto fd = [](DevBuf& b) { if (b.p) cudaFree(b.p); b.p = nullptr; };
	auto fh = [](DevBuf& b) { if (b.p) cudaFreeHost(b.p); b.p = nullptr; };
	fd(ctx->d_ploidy); fd(ctx->d_pclass); fd(ctx->d_class_ploidy); fd(ctx->d_NC); fd(ctx->d_T); fd(ctx->d_ncr); fd(ctx->d_pheno);
	fd(ctx->d_exact_row); fd(ctx->d_exact_cl); fd(ctx->d_exact_done); fd(ctx->d_need_reads);
	for (auto& b : ctx->d_scr) fd(b);
	for (auto& b : ctx->d_b) fd(b);
	fd(ctx->d_out_arena); fh(ctx->h_out_arena);
	fd(ctx->d_rows); fh(ctx->h_rows); fd(ctx->d_arena);
	for (auto& b : ctx->d_nar) fd(b);
	fd(ctx->d_tail_task); fd(ctx->d_tail_head); fd(ctx->d_tail_total); fd(ctx->d_tail_rec); fd(ctx->d_tail_counters);
	fd(ctx->d_perm_tasks); fd(ctx->d_call_tasks); fd(ctx->d_call_tasks2); fd(ctx->d_order); fd(ctx->d_trace); fd(ctx->d_counters); fh(ctx->h_counters); fd(ctx->d_gws);
	fd(ctx->p_pool_off); fd(ctx->p_rec); fd(ctx->p_mate); fd(ctx->p_isize); fd(ctx->p_seq4); fd(ctx->p_qual); fd(ctx->p_ref); fd(ctx->p_flags);
	fd(ctx->p_state); fd(ctx->p_win); fd(ctx->p_flag); fd(ctx->p_cand); fd(ctx->p_sites); fd(ctx->p_work); fd(ctx->p_reads); fd(ctx->p_alt);
	fh(ctx->ph_state); fh(ctx->ph_sites); fh(ctx->ph_reads); fh(ctx->ph_alt);
	for (auto& e : ctx->ev_pile) if (e) cudaEventDestroy(e);
	for (auto& e : ctx->ev) if (e) cudaEventDestroy(e);
	if (ctx->ev_counts) cudaEventDestroy(ctx->ev_counts);
	if (ctx->ev_epoch) cudaEventDestroy(ctx->ev_epoch);
	if (ctx->ev_done) cudaEventDestroy(ctx->ev_done);
	if (ctx->ev_fork) cudaEventDestroy(ctx->ev_fork);
	if (ctx->ev_join) cudaEventDestroy(ctx->ev_join);
	if (ctx->side) { cudaStreamSynchronize(ctx->side); cudaStreamDestroy(ctx->side); }
	if (ctx->own_stream && ctx->stream) cudaStreamDestroy(ctx->stream);
	delete ctx;
}

int crispgpu_submit(crispgpu_ctx* ctx, const crispgpu_batch* batch, int64_t* ticket)
{
	if (!ctx || !batch || batch->n_sites < 0) return CRISPGPU_ERR_ARG;
	if (ctx->inflight) {
		snprintf(ctx->err, sizeof ctx->err, "a batch is already in flight; call crispgpu_wait first");
		return CRISPGPU_ERR_TICKET;
	}
	CK(cudaSetDevice(ctx->device));
	ctx->timing_pending = false;  // the stage events are about to be reused: read crispgpu_get_timing between wait and the next submit
	ctx->pile_timing = false;
	CK(cudaEventRecord(ctx->ev[EV_START], ctx->stream));
	int rc = copy_batch(ctx, batch, true);
	if (rc) return rc;
	ctx->resident = false;
	rc = run_stage1(ctx);
	if (rc) return rc;
	ctx->inflight = true;
	ctx->advanced = false;
	ctx->ticket++;
	if (ticket) *ticket = ctx->ticket;
	return CRISPGPU_OK;
}

int crispgpu_advance(crispgpu_ctx* ctx, int64_t ticket)
{
	if (!ctx) return CRISPGPU_ERR_ARG;
	if (!ctx->inflight || ticket != ctx->ticket) return CRISPGPU_ERR_TICKET;
	if (ctx->advanced) return CRISPGPU_OK;
	CK(cudaSetDevice(ctx->device));
	int rc = run_stage2(ctx, true);
	if (rc) { ctx->inflight = false; return rc; }
	ctx->advanced = true;
	return CRISPGPU_OK;
}

int crispgpu_wait(crispgpu_ctx* ctx, int64_t ticket, crispgpu_results* out)
{
	if (!ctx) return CRISPGPU_ERR_ARG;
	if (!ctx->inflight || ticket != ctx->ticket) return CRISPGPU_ERR_TICKET;
	CK(cudaSetDevice(ctx->device));
	ctx->inflight = false;
	int rc = ctx->advanced ? 0 : run_stage2(ctx, true);
	ctx->advanced = false;
	if (rc) return rc;
	CK(cudaEventSynchronize(ctx->ev_done ? ctx->ev_done : ctx->ev[EV_D2H]));
	ctx->timing_pending = true;
	fill_results(ctx, out);
	if (ctx->rows_on_host) {
		// the rows k_em wrote across the link: one per called allele
		int64_t called = 0;
		const int32_t* va = ctx->res.varalleles;
		for (int s2 = 0; s2 < ctx->n_sites; s2++) called += va[s2];
		ctx->timing.d2h_bytes += called * ctx->P * 12;
		ctx->rows_on_host = false;
	}
	return CRISPGPU_OK;
}

int crispgpu_screen_positions(crispgpu_ctx* ctx, int32_t n, const uint8_t* refbase, const int32_t* counts, const int32_t* indcounts,
                              const uint32_t* ic_off, const crispgpu_count* ic_sparse, uint8_t* maxvec_al, int32_t* maxvec_ct, int32_t* potential)
{
	if (!ctx || n < 0 || !refbase || !counts || (!indcounts && !ic_off) || !maxvec_al || !maxvec_ct || !potential) return CRISPGPU_ERR_ARG;
	if (ctx->inflight) return CRISPGPU_ERR_TICKET;
	if (n == 0) return CRISPGPU_OK;
	CK(cudaSetDevice(ctx->device));
	const size_t cells = (size_t)n * ctx->P;
	DevBuf &d_ref = ctx->d_scr[0], &d_cnt = ctx->d_scr[1], &d_ic = ctx->d_scr[2], &d_off = ctx->d_scr[3], &d_sp = ctx->d_scr[4], &d_al = ctx->d_scr[5], &d_ct = ctx->d_scr[6], &d_pot = ctx->d_scr[7];
	int rc;
	if ((rc = ensure_dev(ctx, d_ref, n)) || (rc = ensure_dev(ctx, d_cnt, (size_t)n * kNC * 4)) || (rc = ensure_dev(ctx, d_ic, cells * kNC * 4)) ||
	    (rc = ensure_dev(ctx, d_al, (size_t)n * 8)) || (rc = ensure_dev(ctx, d_ct, (size_t)n * 32)) || (rc = ensure_dev(ctx, d_pot, (size_t)n * 4))) return rc;
	CK(cudaMemcpyAsync(d_ref.p, refbase, n, cudaMemcpyHostToDevice, ctx->stream));
	CK(cudaMemcpyAsync(d_cnt.p, counts, (size_t)n * kNC * 4, cudaMemcpyHostToDevice, ctx->stream));
	if (indcounts) CK(cudaMemcpyAsync(d_ic.p, indcounts, cells * kNC * 4, cudaMemcpyHostToDevice, ctx->stream));
	else {
		const size_t nsp = ic_off[cells];
		if (nsp && !ic_sparse) return CRISPGPU_ERR_ARG;
		if ((rc = ensure_dev(ctx, d_off, (cells + 1) * 4)) || (rc = ensure_dev(ctx, d_sp, (nsp ? nsp : 1) * 4))) return rc;
		CK(cudaMemcpyAsync(d_off.p, ic_off, (cells + 1) * 4, cudaMemcpyHostToDevice, ctx->stream));
		if (nsp) CK(cudaMemcpyAsync(d_sp.p, ic_sparse, nsp * 4, cudaMemcpyHostToDevice, ctx->stream));
		k_expand_counts<<<ctx->num_sms * 8, 256, 0, ctx->stream>>>(static_cast<const uint32_t*>(d_off.p), static_cast<const uint32_t*>(d_sp.p), static_cast<int32_t*>(d_ic.p), cells);
		CK(launch_check());
	}
	k_candidates<<<(n + 7) / 8, 256, 0, ctx->stream>>>(static_cast<const uint8_t*>(d_ref.p), static_cast<const int32_t*>(d_cnt.p), static_cast<const int32_t*>(d_ic.p),
	                                                   static_cast<const int32_t*>(ctx->d_ploidy.p), n, ctx->P, static_cast<uint8_t*>(d_al.p), static_cast<int32_t*>(d_ct.p), static_cast<int32_t*>(d_pot.p));
	CK(launch_check());
	CK(cudaMemcpyAsync(maxvec_al, d_al.p, (size_t)n * 8, cudaMemcpyDeviceToHost, ctx->stream));
	CK(cudaMemcpyAsync(maxvec_ct, d_ct.p, (size_t)n * 32, cudaMemcpyDeviceToHost, ctx->stream));
	CK(cudaMemcpyAsync(potential, d_pot.p, (size_t)n * 4, cudaMemcpyDeviceToHost, ctx->stream));
	CK(cudaStreamSynchronize(ctx->stream));
	return CRISPGPU_OK;
}


namespace {

struct SiteLayout {   // the candidate-site arrays of a window in one arena (device) mirrored by one pinned arena (host)
	size_t tid, pos, refbase, maxvec_al, maxvec_ct, counts, indcounts, readdepths, mqcounts, filtered, read_off, alt_off, copy_bytes;
	size_t cell_reads, alt_len, alt_cell_off, stats, tstats, total;
};

SiteLayout site_layout(size_t n, size_t P, bool with_stats)
{
	SiteLayout L{};
	size_t o = 0;
	auto take = [&](size_t bytes) { const size_t at = o; o += (bytes + 255) & ~size_t(255); return at; };
	L.tid = take(n * 4); L.pos = take(n * 4); L.refbase = take(n); L.maxvec_al = take(n * 8); L.maxvec_ct = take(n * 32);
	L.counts = take(n * kNC * 4); L.indcounts = take(n * P * kNC * 4); L.readdepths = take(n * P * 4); L.mqcounts = take(n * 16);
	L.filtered = take(n * 8); L.read_off = take((n * P + 1) * 4); L.alt_off = take((n * 5 + 1) * 4);
	L.copy_bytes = o;
	L.cell_reads = take(n * P * 4); L.alt_len = take(n * 5 * P * 4); L.alt_cell_off = take((n * 5 * P + 1) * 4);
	L.stats = take(with_stats ? n * P * 128 : 0); L.tstats = take(with_stats ? n * 128 : 0);
	L.total = o;
	return L;
}

int pile_upload(crispgpu_ctx* ctx, const crispgpu_alignments* aln, const crispgpu_window* win)
{
	if (!aln || !win || aln->n_pools != ctx->P || !aln->pool_off || !win->refseq || win->end < win->start || win->ref_len < 0) {
		snprintf(ctx->err, sizeof ctx->err, "pileup: missing array, pool count different from the context's, or an inverted window");
		return CRISPGPU_ERR_ARG;
	}
	if (ctx->cfg.use_qv == 1) {
		snprintf(ctx->err, sizeof ctx->err, "pileup: --weighted counts (Qhighcounts of 1 - error probability per read) are not accumulated by the device pileup");
		return CRISPGPU_ERR_UNSUPPORTED;
	}
	const size_t P = ctx->P, R = aln->pool_off[P];
	if (R && (!aln->rec || !aln->seq4 || !aln->qual)) return CRISPGPU_ERR_ARG;
	if ((aln->n_bases & 7) || aln->n_bases >= (1ull << 32)) { snprintf(ctx->err, sizeof ctx->err, "pileup: n_bases must be a multiple of 8 below 2^32"); return CRISPGPU_ERR_ARG; }
	int rc;
	int64_t bytes = 0;
	auto up = [&](DevBuf& b, const void* src, size_t n) -> int {
		int r2 = ensure_dev(ctx, b, n ? n : 16);
		if (r2) return r2;
		if (n) CK(cudaMemcpyAsync(b.p, src, n, cudaMemcpyHostToDevice, ctx->stream));
		bytes += (int64_t)n;
		return 0;
	};
	if ((rc = up(ctx->p_pool_off, aln->pool_off, (P + 1) * 4)) || (rc = up(ctx->p_rec, aln->rec, R * sizeof(crispgpu_alnrec))) ||
	    (rc = up(ctx->p_seq4, aln->seq4, aln->n_bases / 2)) || (rc = up(ctx->p_qual, aln->qual, aln->n_bases)) ||
	    (rc = up(ctx->p_ref, win->refseq, (size_t)win->ref_len))) return rc;
	if (aln->mate_pos && (rc = up(ctx->p_mate, aln->mate_pos, R * 4))) return rc;
	if (aln->mate_pos && aln->isize && (rc = up(ctx->p_isize, aln->isize, R * 4))) return rc;
	if ((rc = ensure_dev(ctx, ctx->p_flags, R ? R : 16)) || (rc = ensure_dev(ctx, ctx->p_state, 64)) || (rc = ensure_host(ctx, ctx->ph_state, 64))) return rc;
	PileReads& rd = ctx->prd;
	rd.P = (int)P; rd.n_reads = (uint32_t)R; rd.n_bases = aln->n_bases;
	rd.pool_off = static_cast<const uint32_t*>(ctx->p_pool_off.p);
	rd.rec = static_cast<const crispgpu_alnrec*>(ctx->p_rec.p);
	rd.mate_pos = aln->mate_pos ? static_cast<const int32_t*>(ctx->p_mate.p) : nullptr;
	rd.isize = (aln->mate_pos && aln->isize) ? static_cast<const int32_t*>(ctx->p_isize.p) : nullptr;
	rd.seq4 = static_cast<const uint8_t*>(ctx->p_seq4.p);
	rd.qual = static_cast<const uint8_t*>(ctx->p_qual.p);
	rd.flags = static_cast<uint8_t*>(ctx->p_flags.p);
	rd.state = static_cast<int32_t*>(ctx->p_state.p);
	PileWindow& w = ctx->pwin;
	w.tid = win->tid; w.w0 = win->start; w.W = win->end - win->start; w.ref0 = win->ref_start; w.reflen = win->ref_len;
	w.refseq = static_cast<const uint8_t*>(ctx->p_ref.p);
	w.min_q = ctx->cfg.min_q; w.min_m = win->min_mapq;
	ctx->pile_streams = win->want_streams != 0;
	ctx->pile_h2d = bytes;
	return 0;
}

// alignments resident in device memory -> candidate sites packed as the engine's batch -> the engine's two stages
int pile_run(crispgpu_ctx* ctx, bool fetch, crispgpu_pileup_sites* sites, crispgpu_results* results)
{
	const PileReads& rd = ctx->prd;
	const PileWindow& w = ctx->pwin;
	const size_t P = ctx->P, W = (size_t)w.W;
	int rc;
	for (auto& e : ctx->ev_pile) if (!e) CK(cudaEventCreate(&e));
	CK(cudaEventRecord(ctx->ev_pile[0], ctx->stream));
	CK(cudaMemsetAsync(ctx->p_state.p, 0, 64, ctx->stream));
	int64_t launches = 0;
	if (rd.n_reads) {
		const int need = (int)((rd.n_reads + 255) / 256);
		k_pile_check<<<need < ctx->num_sms * 8 ? need : ctx->num_sms * 8, 256, 0, ctx->stream>>>(rd);
		CK(launch_check());
		launches++;
	}
	const size_t n_words = (W + 31) / 32;
	if ((rc = ensure_dev(ctx, ctx->p_win, W * P * kPileCell * 4 + 16)) || (rc = ensure_dev(ctx, ctx->p_flag, n_words * 4 + 16)) || (rc = ensure_dev(ctx, ctx->p_cand, W * 4 + 16))) return rc;
	uint32_t* d_win = static_cast<uint32_t*>(ctx->p_win.p);
	CK(cudaEventRecord(ctx->ev_pile[1], ctx->stream));
	if (W) {
		const dim3 grid((unsigned)((W + kPileTile - 1) / kPileTile), (unsigned)P);
		k_pile_tile<<<grid, 256, 0, ctx->stream>>>(rd, w, d_win);
		CK(launch_check());
		launches++;
	}
	CK(cudaEventRecord(ctx->ev_pile[2], ctx->stream));
	int32_t* hs = static_cast<int32_t*>(ctx->ph_state.p);
	if (W) {
		uint32_t* bits = static_cast<uint32_t*>(ctx->p_flag.p);
		k_pile_screen<<<(unsigned)n_words, 256, 0, ctx->stream>>>(d_win, w, (int)P, static_cast<const int32_t*>(ctx->d_ploidy.p), bits);
		CK(launch_check());
		k_pile_select<<<(unsigned)((n_words + kPileSelectWords - 1) / kPileSelectWords), kPileSelectWords, 0, ctx->stream>>>(bits, (int)n_words, static_cast<int32_t*>(ctx->p_cand.p), rd.state + 2);
		CK(launch_check());
		launches += 2;
	}
	CK(cudaMemcpyAsync(hs, ctx->p_state.p, 16, cudaMemcpyDeviceToHost, ctx->stream));
	CK(cudaStreamSynchronize(ctx->stream));
	if (hs[0]) {
		snprintf(ctx->err, sizeof ctx->err, "pileup: alignments outside the interface%s%s%s%s%s", (hs[0] & kPileErrAlign) ? "; base_off not a multiple of 8" : "",
		         (hs[0] & kPileErrBounds) ? "; a read reaches past n_bases (or has an empty match block)" : "", (hs[0] & kPileErrOrder) ? "; a pool's alignments are not in coordinate order" : "",
		         (hs[0] & kPileErrMate) ? "; overlapping mates (the reference's mate merge is not piled up on the device)" : "", (hs[0] & kPileErrPool) ? "; pool_off is not a partition of the alignments" : "");
		return (hs[0] & kPileErrMate) ? CRISPGPU_ERR_UNSUPPORTED : CRISPGPU_ERR_ARG;
	}
	const size_t n = (size_t)hs[2];
	const bool with_stats = ctx->cfg.em_flag == 0;   // the legacy caller reads the per-pool error-probability sums (chernoff_bound)
	const SiteLayout L = site_layout(n ? n : 1, P, with_stats);
	if ((rc = ensure_dev(ctx, ctx->p_sites, L.total)) || (rc = ensure_host(ctx, ctx->ph_sites, L.copy_bytes))) return rc;
	unsigned char* ds = static_cast<unsigned char*>(ctx->p_sites.p);
	PileSites ps{};
	ps.n = (int)n;
	ps.cand = static_cast<const int32_t*>(ctx->p_cand.p);
	ps.tid = reinterpret_cast<int32_t*>(ds + L.tid); ps.pos = reinterpret_cast<int32_t*>(ds + L.pos); ps.refbase = ds + L.refbase;
	ps.maxvec_al = ds + L.maxvec_al; ps.maxvec_ct = reinterpret_cast<int32_t*>(ds + L.maxvec_ct); ps.counts = reinterpret_cast<int32_t*>(ds + L.counts);
	ps.indcounts = reinterpret_cast<int32_t*>(ds + L.indcounts); ps.readdepths = reinterpret_cast<int32_t*>(ds + L.readdepths);
	ps.mqcounts = reinterpret_cast<int32_t*>(ds + L.mqcounts); ps.filtered = reinterpret_cast<int32_t*>(ds + L.filtered);
	ps.read_off = reinterpret_cast<uint32_t*>(ds + L.read_off); ps.alt_off = reinterpret_cast<uint32_t*>(ds + L.alt_off);
	ps.cell_reads = reinterpret_cast<uint32_t*>(ds + L.cell_reads); ps.alt_len = reinterpret_cast<uint32_t*>(ds + L.alt_len);
	ps.alt_cell_off = reinterpret_cast<uint32_t*>(ds + L.alt_cell_off);
	ps.stats = with_stats ? reinterpret_cast<double*>(ds + L.stats) : nullptr;
	ps.tstats = with_stats ? reinterpret_cast<double*>(ds + L.tstats) : nullptr;
	size_t n_reads = 0, n_alt = 0;
	if (n) {
		const size_t cells = n * P;
		const int wgrid = (int)((cells + 7) / 8 < (size_t)ctx->num_sms * 16 ? (cells + 7) / 8 : (size_t)ctx->num_sms * 16);
		k_pile_gather_counts<<<(int)(n < (size_t)ctx->num_sms * 8 ? n : (size_t)ctx->num_sms * 8), 128, 0, ctx->stream>>>(ps, w, (int)P, d_win);
		CK(launch_check());
		k_pile_scan<<<1, 1024, 0, ctx->stream>>>(ps.cell_reads, ps.read_off, cells);
		k_pile_gather_reads<0><<<wgrid, 256, 0, ctx->stream>>>(rd, w, ps);
		CK(launch_check());
		k_pile_scan<<<1, 1024, 0, ctx->stream>>>(ps.alt_len, ps.alt_cell_off, cells * 5);
		CK(launch_check());
		CK(cudaMemcpyAsync(hs + 4, ps.read_off + cells, 4, cudaMemcpyDeviceToHost, ctx->stream));
		CK(cudaMemcpyAsync(hs + 5, ps.alt_cell_off + cells * 5, 4, cudaMemcpyDeviceToHost, ctx->stream));
		CK(cudaStreamSynchronize(ctx->stream));
		n_reads = static_cast<uint32_t>(hs[4]);
		n_alt = static_cast<uint32_t>(hs[5]);
		if ((rc = ensure_dev(ctx, ctx->p_reads, n_reads * 2 + 16)) || (rc = ensure_dev(ctx, ctx->p_alt, n_alt * sizeof(crispgpu_altread) + 16))) return rc;
		ps.reads = static_cast<uint16_t*>(ctx->p_reads.p);
		ps.alt_reads = static_cast<crispgpu_altread*>(ctx->p_alt.p);
		k_pile_gather_reads<1><<<wgrid, 256, 0, ctx->stream>>>(rd, w, ps);
		CK(launch_check());
		k_pile_alt_off<<<(int)((n * 5 + 256) / 256), 256, 0, ctx->stream>>>(ps, (int)P);
		CK(launch_check());
		if (with_stats) {
			k_pile_tstats<<<(int)((n * 16 + 255) / 256), 256, 0, ctx->stream>>>(ps, (int)P);
			CK(launch_check());
			launches++;
		}
		launches += 6;
	}
	CK(cudaEventRecord(ctx->ev_pile[3], ctx->stream));
	// the packed sites are the engine's batch
	DevBatch& b = ctx->db;
	b = DevBatch{};
	b.n_sites = (int)n; b.P = (int)P;
	b.tid = ps.tid; b.pos = ps.pos; b.refbase = ps.refbase; b.maxvec_al = ps.maxvec_al; b.maxvec_ct = ps.maxvec_ct; b.counts = ps.counts;
	b.indcounts = ps.indcounts; b.read_off = ps.read_off; b.reads = ps.reads; b.alt_off = ps.alt_off; b.alt_reads = ps.alt_reads;
	b.stats = ps.stats; b.tstats = ps.tstats;
	ctx->n_sites = (int)n; ctx->n_reads = n_reads; ctx->n_bins = 0;
	ctx->binned = false; ctx->sparse_ic = false; ctx->device_sort = false; ctx->lazy_reads = nullptr; ctx->resident = false;
	ctx->copied_bytes = ctx->pile_h2d;
	ctx->timing.h2d_bytes = ctx->pile_h2d;
	if ((rc = run_stage1(ctx)) || (rc = run_stage2(ctx, fetch))) return rc;
	ctx->timing.launches += launches;
	ctx->timing.n_positions = (int32_t)W;
	ctx->timing.n_alignments = (int32_t)rd.n_reads;
	ctx->timing.pile_bytes = (int64_t)rd.n_reads * 16 + (int64_t)(rd.n_bases + rd.n_bases / 2) + (int64_t)(W * P * kPileCell * 4);
	if (fetch && n) {
		CK(cudaMemcpyAsync(ctx->ph_sites.p, ds, L.copy_bytes, cudaMemcpyDeviceToHost, ctx->stream));
		ctx->timing.d2h_bytes += (int64_t)L.copy_bytes;
		if (ctx->pile_streams) {
			if ((rc = ensure_host(ctx, ctx->ph_reads, n_reads * 2 + 16)) || (rc = ensure_host(ctx, ctx->ph_alt, n_alt * sizeof(crispgpu_altread) + 16))) return rc;
			if (n_reads) CK(cudaMemcpyAsync(ctx->ph_reads.p, ps.reads, n_reads * 2, cudaMemcpyDeviceToHost, ctx->stream));
			if (n_alt) CK(cudaMemcpyAsync(ctx->ph_alt.p, ps.alt_reads, n_alt * sizeof(crispgpu_altread), cudaMemcpyDeviceToHost, ctx->stream));
		}
	}
	CK(cudaStreamSynchronize(ctx->stream));
	ctx->pile_timing = true;
	ctx->timing_pending = true;
	if (fetch) {
		fill_results(ctx, results);
		if (ctx->rows_on_host) {
			int64_t called = 0;
			for (size_t s2 = 0; s2 < n; s2++) called += ctx->res.varalleles[s2];
			ctx->timing.d2h_bytes += called * (int64_t)P * 12;
			ctx->rows_on_host = false;
		}
	}
	if (sites) {
		const unsigned char* hsb = static_cast<const unsigned char*>(ctx->ph_sites.p);
		*sites = crispgpu_pileup_sites{};
		sites->n_sites = (int32_t)n;
		sites->n_positions = (int32_t)W;
		if (fetch && n) {
			sites->pos = reinterpret_cast<const int32_t*>(hsb + L.pos); sites->refbase = hsb + L.refbase; sites->maxvec_al = hsb + L.maxvec_al;
			sites->maxvec_ct = reinterpret_cast<const int32_t*>(hsb + L.maxvec_ct); sites->counts = reinterpret_cast<const int32_t*>(hsb + L.counts);
			sites->indcounts = reinterpret_cast<const int32_t*>(hsb + L.indcounts); sites->readdepths = reinterpret_cast<const int32_t*>(hsb + L.readdepths);
			sites->mqcounts = reinterpret_cast<const int32_t*>(hsb + L.mqcounts); sites->filtered = reinterpret_cast<const int32_t*>(hsb + L.filtered);
			if (ctx->pile_streams) {
				sites->read_off = reinterpret_cast<const uint32_t*>(hsb + L.read_off); sites->alt_off = reinterpret_cast<const uint32_t*>(hsb + L.alt_off);
				sites->reads = static_cast<const crispgpu_read*>(ctx->ph_reads.p); sites->alt_reads = static_cast<const crispgpu_altread*>(ctx->ph_alt.p);
			}
		}
	}
	return CRISPGPU_OK;
}

}  // namespace

int crispgpu_pileup_run(crispgpu_ctx* ctx, const crispgpu_alignments* aln, const crispgpu_window* win, crispgpu_pileup_sites* sites, crispgpu_results* results)
{
	if (!ctx) return CRISPGPU_ERR_ARG;
	if (ctx->inflight) return CRISPGPU_ERR_TICKET;
	CK(cudaSetDevice(ctx->device));
	ctx->timing_pending = false;
	CK(cudaEventRecord(ctx->ev[EV_START], ctx->stream));
	int rc = pile_upload(ctx, aln, win);
	if (rc) return rc;
	ctx->pile_resident = false;
	return pile_run(ctx, true, sites, results);
}

int crispgpu_pileup_upload(crispgpu_ctx* ctx, const crispgpu_alignments* aln, const crispgpu_window* win)
{
	if (!ctx) return CRISPGPU_ERR_ARG;
	if (ctx->inflight) return CRISPGPU_ERR_TICKET;
	CK(cudaSetDevice(ctx->device));
	int rc = pile_upload(ctx, aln, win);
	if (rc) return rc;
	CK(cudaStreamSynchronize(ctx->stream));
	ctx->pile_resident = true;
	return CRISPGPU_OK;
}

int crispgpu_pileup_set_window(crispgpu_ctx* ctx, const crispgpu_window* win)
{
	if (!ctx || !win) return CRISPGPU_ERR_ARG;
	if (ctx->inflight) return CRISPGPU_ERR_TICKET;
	if (!ctx->pile_resident) { snprintf(ctx->err, sizeof ctx->err, "no resident alignments: call crispgpu_pileup_upload first"); return CRISPGPU_ERR_ARG; }
	if (win->end < win->start) return CRISPGPU_ERR_ARG;
	CK(cudaSetDevice(ctx->device));
	PileWindow& w = ctx->pwin;
	if (win->refseq) {   // a new reference stretch; NULL keeps the one that is on the device
		if (win->ref_len < 0) return CRISPGPU_ERR_ARG;
		int rc = ensure_dev(ctx, ctx->p_ref, win->ref_len ? (size_t)win->ref_len : 16);
		if (rc) return rc;
		if (win->ref_len) CK(cudaMemcpyAsync(ctx->p_ref.p, win->refseq, (size_t)win->ref_len, cudaMemcpyHostToDevice, ctx->stream));
		w.ref0 = win->ref_start; w.reflen = win->ref_len;
		w.refseq = static_cast<const uint8_t*>(ctx->p_ref.p);
	}
	w.tid = win->tid; w.w0 = win->start; w.W = win->end - win->start;
	w.min_m = win->min_mapq;
	ctx->pile_streams = win->want_streams != 0;
	ctx->pile_h2d = 0;   // nothing of the alignments moves for this window
	return CRISPGPU_OK;
}

int crispgpu_pileup_run_resident(crispgpu_ctx* ctx, int fetch_results, crispgpu_pileup_sites* sites, crispgpu_results* results)
{
	if (!ctx) return CRISPGPU_ERR_ARG;
	if (!ctx->pile_resident) { snprintf(ctx->err, sizeof ctx->err, "no resident alignments: call crispgpu_pileup_upload first"); return CRISPGPU_ERR_ARG; }
	CK(cudaSetDevice(ctx->device));
	ctx->timing_pending = false;
	CK(cudaEventRecord(ctx->ev[EV_START], ctx->stream));
	return pile_run(ctx, fetch_results != 0, sites, results);
}


namespace {

struct NarrowPlan { size_t nb16, nic16; bool ok; bool alt32; size_t nalt; };

NarrowPlan narrow_plan(const crispgpu_batch* b, size_t P)
{
	NarrowPlan pl{0, 0, false, false, 0};
	const size_t n = (size_t)b->n_sites;
	pl.nalt = (b->alt_off && n) ? b->alt_off[n * 5] : 0;
	pl.alt32 = b->alt_off != nullptr && (pl.nalt == 0 || b->alt_reads != nullptr);
	for (size_t k = 0; k < pl.nalt && pl.alt32; k++)
		if (b->alt_reads[k].pool > 0x7ff || b->alt_reads[k].offset > 0x3ff || b->alt_reads[k].aligned > 0x3ff || b->alt_reads[k].strand > 1) pl.alt32 = false;
	if (!b->bin_off || !b->ic_off || (n && b->bin_off[n * P] && !b->bins) || (n && b->ic_off[n * P] && !b->ic_sparse)) return pl;
	for (size_t c = 0; c < n * P; c++) {
		size_t e = 0;
		for (uint32_t k = b->bin_off[c]; k < b->bin_off[c + 1]; k++) e += ((b->bins[k] >> 16) + 31) / 32;
		if (e > 255) return pl;
		pl.nb16 += e;
		e = 0;
		for (uint32_t k = b->ic_off[c]; k < b->ic_off[c + 1]; k++) e += ((b->ic_sparse[k] & 0x7ffffffu) + 2046) / 2047;
		if (e > 255) return pl;
		pl.nic16 += e;
	}
	pl.ok = true;
	return pl;
}

// the arrays of a batch that travel unchanged in a narrow arena (everything but the read stream and the four compact arrays)
struct Span { const void** field; size_t bytes; };
int narrow_spans(crispgpu_batch* v, size_t P, Span* sp, bool alt32)
{
	const size_t n = (size_t)v->n_sites;
	int k = 0;
	auto add = [&](const void** f, size_t bytes) { sp[k].field = f; sp[k].bytes = *f ? bytes : 0; k++; };
	add((const void**)&v->tid, n * 4); add((const void**)&v->pos, n * 4); add((const void**)&v->refbase, n);
	// the site totals are recomputed on the device from the per-pool counts (k_expand_counts16): they do not travel.  The allele
	// order does (40 B per site): sorting every site again on the device costs the single-GPU step more than the bytes cost
	v->counts = nullptr;
	add((const void**)&v->maxvec_al, n * 8); add((const void**)&v->maxvec_ct, n * 32);
	add((const void**)&v->indel_len, n * 16); add((const void**)&v->hplen, n * 16);
	add((const void**)&v->alt_off, (n * 5 + 1) * 4);
	if (alt32) v->alt_reads = nullptr;
	add((const void**)&v->alt_reads, (v->alt_off && n) ? (size_t)v->alt_off[n * 5] * sizeof(crispgpu_altread) : 0);
	add((const void**)&v->amb_idx, n * 20); add((const void**)&v->amb_dec, (size_t)v->n_amb * P * 16); add((const void**)&v->amb_ep, (size_t)v->n_amb * P * 16);
	add((const void**)&v->qhigh, n * P * 128); add((const void**)&v->stats, n * P * 128); add((const void**)&v->tstats, n * 128);
	return k;
}

size_t pad256(size_t x) { return (x + 255) & ~size_t(255); }

}  // namespace

namespace {

size_t narrow_bytes_of(const crispgpu_batch* compact, size_t P, const NarrowPlan& pl)
{
	const size_t n = (size_t)compact->n_sites;
	crispgpu_batch v = *compact;
	Span sp[20];
	const int k = narrow_spans(&v, P, sp, pl.alt32);
	size_t total = pl.alt32 ? pad256(pl.nalt * 4) : 0;
	for (int i = 0; i < k; i++) total += pad256(sp[i].bytes);
	total += pad256(pl.nb16 * 2) + pad256(n * P) + pad256((n + 1) * 4) + pad256(pl.nic16 * 2) + pad256(n * P) + pad256((n + 1) * 4);
	return total + 256;
}

}  // namespace

size_t crispgpu_batch_narrow_bytes(const crispgpu_batch* compact, int32_t n_pools)
{
	if (!compact || n_pools <= 0 || compact->n_sites < 0) return 0;
	const NarrowPlan pl = narrow_plan(compact, (size_t)n_pools);
	return pl.ok ? narrow_bytes_of(compact, (size_t)n_pools, pl) : 0;
}

int crispgpu_batch_pack_narrow(const crispgpu_batch* compact, int32_t n_pools, void* arena, size_t arena_cap, crispgpu_batch* out)
{
	if (!compact || !arena || !out || n_pools <= 0 || compact->n_sites < 0 || (reinterpret_cast<uintptr_t>(arena) & 255)) return CRISPGPU_ERR_ARG;
	const size_t P = (size_t)n_pools, n = (size_t)compact->n_sites;
	const NarrowPlan pl = narrow_plan(compact, P);   // one pass over the entries sizes the narrow arrays; the second pass writes them
	if (!pl.ok || narrow_bytes_of(compact, P, pl) > arena_cap) return CRISPGPU_ERR_ARG;
	*out = *compact;
	Span sp[20];
	const int k = narrow_spans(out, P, sp, pl.alt32);
	unsigned char* base = static_cast<unsigned char*>(arena);
	size_t off = 0;
	for (int i = 0; i < k; i++) {
		if (*sp[i].field == nullptr) continue;                          // absent
		if (sp[i].bytes == 0) { *sp[i].field = base + off; continue; }   // present but empty (a batch without sites)
		memcpy(base + off, *sp[i].field, sp[i].bytes);
		*sp[i].field = base + off;
		off += pad256(sp[i].bytes);
	}
	auto take = [&](size_t bytes) { unsigned char* p = base + off; off += pad256(bytes); return p; };
	uint16_t* b16 = reinterpret_cast<uint16_t*>(take(pl.nb16 * 2));
	uint8_t* blen = take(n * P);
	uint32_t* bsite = reinterpret_cast<uint32_t*>(take((n + 1) * 4));
	uint16_t* i16 = reinterpret_cast<uint16_t*>(take(pl.nic16 * 2));
	uint8_t* ilen = take(n * P);
	uint32_t* isite = reinterpret_cast<uint32_t*>(take((n + 1) * 4));
	size_t kb = 0, ki = 0;
	for (size_t s = 0; s < n; s++) {
		bsite[s] = (uint32_t)kb;
		isite[s] = (uint32_t)ki;
		for (size_t i = 0; i < P; i++) {
			const size_t c = s * P + i, kb0 = kb, ki0 = ki;
			for (uint32_t e = compact->bin_off[c]; e < compact->bin_off[c + 1]; e++) {
				const uint32_t w = compact->bins[e];
				for (uint32_t left = w >> 16; left > 0;) {
					const uint32_t take_n = left > 32 ? 32 : left;
					b16[kb++] = CRISPGPU_BIN16(w & 0xffffu, take_n);
					left -= take_n;
				}
			}
			for (uint32_t e = compact->ic_off[c]; e < compact->ic_off[c + 1]; e++) {
				const uint32_t w = compact->ic_sparse[e];
				for (uint32_t left = w & 0x7ffffffu; left > 0;) {
					const uint32_t take_n = left > 2047 ? 2047 : left;
					i16[ki++] = CRISPGPU_COUNT16(w >> 27, take_n);
					left -= take_n;
				}
			}
			blen[c] = (uint8_t)(kb - kb0);
			ilen[c] = (uint8_t)(ki - ki0);
		}
	}
	bsite[n] = (uint32_t)kb;
	isite[n] = (uint32_t)ki;
	out->alt32 = nullptr;
	if (pl.alt32) {
		uint32_t* a32 = reinterpret_cast<uint32_t*>(take(pl.nalt * 4));
		for (size_t q = 0; q < pl.nalt; q++) a32[q] = CRISPGPU_ALT32(compact->alt_reads[q].pool, compact->alt_reads[q].offset, compact->alt_reads[q].aligned, compact->alt_reads[q].strand);
		out->alt32 = a32;
	}
	out->indcounts = nullptr; out->read_off = nullptr; out->reads = nullptr;
	out->bin_off = nullptr; out->bins = nullptr; out->ic_off = nullptr; out->ic_sparse = nullptr;
	out->bins16 = b16; out->bin_len = blen; out->bin_site = bsite; out->ic16 = i16; out->ic_len = ilen; out->ic_site = isite;
	out->arena_base = arena;
	out->arena_bytes = off;
	return CRISPGPU_OK;
}

int crispgpu_upload(crispgpu_ctx* ctx, const crispgpu_batch* batch)
{
	if (!ctx || !batch || batch->n_sites < 0) return CRISPGPU_ERR_ARG;
	if (ctx->inflight) return CRISPGPU_ERR_TICKET;
	CK(cudaSetDevice(ctx->device));
	int rc = copy_batch(ctx, batch, false);
	if (rc) return rc;
	CK(cudaStreamSynchronize(ctx->stream));
	ctx->resident = true;
	return CRISPGPU_OK;
}

int crispgpu_run_resident(crispgpu_ctx* ctx, int fetch_results, crispgpu_results* out)
{
	if (!ctx) return CRISPGPU_ERR_ARG;
	if (!ctx->resident) {
		snprintf(ctx->err, sizeof ctx->err, "no resident batch: call crispgpu_upload first");
		return CRISPGPU_ERR_ARG;
	}
	CK(cudaSetDevice(ctx->device));
	ctx->pile_timing = false;
	CK(cudaEventRecord(ctx->ev[EV_START], ctx->stream));
	int rc = run_stage1(ctx);
	if (rc) return rc;
	rc = run_stage2(ctx, fetch_results != 0);
	if (rc) return rc;
	CK(cudaEventSynchronize(ctx->ev[EV_D2H]));
	ctx->timing_pending = true;
	if (fetch_results) fill_results(ctx, out);
	return CRISPGPU_OK;
}

// Debugging aid (not part of the documented ABI): when each stage boundary of ctx's last completed batch happened,
// in ms relative to `epoch` (a context's creation).  out[0..6] = START, H2D, EVAL, PERM, FILTER, CALL, D2H.
extern "C" int crispgpu_debug_timeline(crispgpu_ctx* ctx, crispgpu_ctx* epoch, float* out)
{
	if (!ctx || !epoch || !out) return CRISPGPU_ERR_ARG;
	for (int k = 0; k < EV_COUNT && k < 7; k++)
		if (cudaEventElapsedTime(out + k, epoch->ev_epoch, ctx->ev[k]) != cudaSuccess) return CRISPGPU_ERR_CUDA;
	return CRISPGPU_OK;
}

int crispgpu_get_timing(crispgpu_ctx* ctx, crispgpu_timing* out)
{
	if (!ctx || !out) return CRISPGPU_ERR_ARG;
	if (ctx->timing_pending) {  // the event queries are kept off the submit/wait path
		CK(cudaSetDevice(ctx->device));
		int rc = finish_timing(ctx);
		if (rc) return rc;
		ctx->timing_pending = false;
	}
	*out = ctx->timing;
	return CRISPGPU_OK;
}

const char* crispgpu_last_error(crispgpu_ctx* ctx) { return ctx ? ctx->err : "null context"; }

void* crispgpu_host_alloc(size_t bytes)
{
	void* p = nullptr;
	// mapped + portable: device-addressable from every GPU, so the packed reads can be fetched on demand (k_fetch_reads)
	if (cudaHostAlloc(&p, bytes ? bytes : 1, cudaHostAllocMapped | cudaHostAllocPortable) != cudaSuccess) return nullptr;
	return p;
}

void crispgpu_host_free(void* p)
{
	if (p) cudaFreeHost(p);
}

}  // extern "C"
