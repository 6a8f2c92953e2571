// k_permute.cuh — Monte-Carlo contingency-table p-values, one thread per permutation.
//
//   k_permute_stranded (+ _tail, _resolve)   pvalue_contable_iter_stranded  FET/contables.c:261-339  (pool size > 2)
//   k_permute_lowcov   (+ _tail, _resolve)   pvalue_lowcoverage             FET/lowcovFET.c:23-149   (pool size == 2)
//
// Null law: the alt reads of each strand are placed without replacement among that strand's read slots
// (multivariate hypergeometric per strand, strands independent).  The reference shuffles a slot->pool index with
// drand48; here every permutation k owns a Philox4x32-10 stream keyed by (seed, tid, pos, slot) with counter
// (k, block, stratum), so the result does not depend on batching, grid size or GPU count.  A ball is placed by
// drawing a uniform slot, finding its pool in the shared-memory cumulative marginals, and rejecting the draw when it
// lands on one of the pool's already-occupied slots (slots of a pool are exchangeable).  The reference's sequential
// stopping rule is evaluated on the prefix sums of the hit flags in permutation order, so the reported (hits, k) is
// exactly what a serial run of the same streams gives.
//
// Head kernel: one CTA per task, the first 1024 permutations (decides every non-significant table and covers the
// early clauses of the stop rules).  Tail kernel: undecided tasks are cut into <= 256 chunks processed as
// (chunk, task) work items, chunk-major, each recording its hit counts and first ten hit indices; a chunk is skipped
// once the break condition is known to have been met in earlier chunks.  Resolve kernel: walks the records in
// permutation order to the break index.
#pragma once
#include "dmath.cuh"
#include "engine_types.cuh"
#include "k_evaluate.cuh"

namespace crisp {

struct PermShared {
	int task;
	int hits_before;
	int stop_k, stop_hits;
	int warp_hits[32];
	int warp_hits2[32];
	int warp_first[32];
	int tot[3], ones[3];
	int maxN;
	int maxball;   // most balls one pool can receive on one strand / stratum in a permutation of this table
	int occ_bytes; // stranded test: width of this table's occupancy words (1, 2 or 4 bytes: two 4-, 8- or 16-bit ball counters)
	int hash_log;  // stranded test: > 0 = the permutation's occupancy is a per-thread hash table of 2^hash_log entries
	int active;    // threads of the block that run permutations (the occupancy area holds that many tables / columns)
	double clra;
};

constexpr int kPermHashPools = 128;   // from this many pools on the stranded kernels may keep a permutation's occupancy in a hash table

// inclusive prefix sum of a 0/1 flag over the block in thread order; returns this thread's inclusive count and the block total
__device__ __forceinline__ int block_prefix(int flag, int* warp_tot, int& total)
{
	const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5, nw = blockDim.x >> 5;
	const unsigned bal = __ballot_sync(0xffffffffu, flag);
	int inc = __popc(bal & (0xffffffffu >> (31 - lane)));
	if (lane == 0) warp_tot[warp] = __popc(bal);
	__syncthreads();
	int before = 0, t = 0;
	for (int w = 0; w < nw; w++) { int v = warp_tot[w]; if (w < warp) before += v; t += v; }
	total = t;
	__syncthreads();
	return inc + before;
}

// place `ones` balls of one stratum; cum = cumulative slots per pool (cum[P] = M); occ(i) reads and bump(i) increments
// the occupied count of pool i in this stratum
template <typename Occ, typename Bump>
__device__ __forceinline__ void place_balls(Philox& rng, const uint32_t* cum, int P, uint32_t M, int ones, Occ occ, Bump bump)
{
	// pool of slot u = largest i with cum[i] <= u.  Pools of one experiment hold similar numbers of reads, so the
	// proportional guess u*P/M is usually right or off by one; a short walk fixes it (exact for any marginals).
	const float scale = M ? (float)P / (float)M : 0.f;
	for (int d = 0; d < ones; d++) {
		for (;;) {
			const uint32_t u = rng.below(M);
			int lo = min(P - 1, (int)((float)u * scale));
			while (cum[lo] > u) lo--;
			while (cum[lo + 1] <= u) lo++;
			if (u - cum[lo] >= (uint32_t)occ(lo)) { bump(lo); break; }
		}
	}
}

// ---- stranded test: shared-memory view of one task -----------------------------------------------------------
struct StrandedView {
	PermShared* sh;
	uint32_t* cumf;
	uint32_t* cumr;
	int32_t* Ntot;
	int32_t* alt;
	double* L10;
	unsigned char* occ;  // column mode: [P][active] of OccT, low half forward balls, high half reverse balls (32-bit words, or 16-bit
	                     // when no pool holds more than 255 reads per strand: twice the permutations in flight for the same shared
	                     // memory).  Hash mode (many pools, few balls): [2^hash_log][active] 32-bit entries
	                     // (pool + 1) << 20 | reverse balls << 10 | forward balls, open addressing — shared memory and the per-
	                     // permutation zero fill then scale with the balls of the table, not with its pools.
	int l10cap;
	int area;            // bytes of the occupancy area
};

__device__ __forceinline__ StrandedView stranded_view(unsigned char* smem_raw, int P, int l10cap, int area)
{
	StrandedView v;
	v.sh = reinterpret_cast<PermShared*>(smem_raw);
	size_t off = (sizeof(PermShared) + 15) & ~size_t(15);
	v.cumf = reinterpret_cast<uint32_t*>(smem_raw + off); off += sizeof(uint32_t) * (P + 1);
	v.cumr = reinterpret_cast<uint32_t*>(smem_raw + off); off += sizeof(uint32_t) * (P + 1);
	v.Ntot = reinterpret_cast<int32_t*>(smem_raw + off); off += sizeof(int32_t) * P;
	v.alt = reinterpret_cast<int32_t*>(smem_raw + off); off += sizeof(int32_t) * 2 * P;
	off = (off + 15) & ~size_t(15);
	v.L10 = reinterpret_cast<double*>(smem_raw + off); off += sizeof(double) * l10cap;
	v.occ = smem_raw + off;
	v.l10cap = l10cap;
	v.area = area;
	return v;
}

// Block-wide: marginals (ctable_stranded columns 0..3, allelecounts.c:221-227), cumulative slot counts, the observed
// statistic clra = sum_i log10 C(Nf+Nr, af+ar) (contables.c:296-298) and the log10 table.  Ends with a barrier.
__device__ inline void stranded_setup(const EngineParams& prm, const StrandedView& v, int o)
{
	const DevBatch& b = prm.b;
	const int P = prm.c.P, T = blockDim.x, tid = threadIdx.x;
	const int s = o / kSlots, slot = o - s * kSlots;
	int a1, a2, a3;
	slot_alleles(b.maxvec_al + s * 8, slot, a1, a2, a3);
	const int refbase = b.refbase[s];
	const int row = slot_amb_row(b, s, slot, a2, a3);
	const int altal = (a1 == refbase || a2 != refbase) ? a2 : a1;
	for (int i = tid; i < P; i += T) {
		const int32_t* r = b.indcounts + ((size_t)s * P + i) * kNC;
		int dd[2] = {0, 0};
		if (row >= 0) { const int32_t* d = b.amb_dec + ((size_t)row * P + i) * 4; dd[0] = d[0]; dd[1] = d[1]; }
		const int nref = (a1 == refbase) + (a2 == refbase) + (a3 == refbase);
		const int nf = r[a1] + r[a2] + r[a3] - nref * dd[0], nr = r[a1 + 8] + r[a2 + 8] + r[a3 + 8] - nref * dd[1];
		// an inconsistent batch (ambiguity decrements above the counts) must not wrap the unsigned capacities or ask for more
		// balls than slots — the sampler would never finish; for a valid batch the clamps change nothing
		const int nfc = nf < 0 ? 0 : nf, nrc = nr < 0 ? 0 : nr;
		v.cumf[i + 1] = nfc;
		v.cumr[i + 1] = nrc;
		v.Ntot[i] = nfc + nrc;
		v.alt[i] = min(max(r[altal] - (altal == refbase ? dd[0] : 0), 0), nfc);
		v.alt[P + i] = min(max(r[altal + 8] - (altal == refbase ? dd[1] : 0), 0), nrc);
	}
	__syncthreads();
	double* terms = v.L10;  // borrowed before the log table is built
	for (int i = tid; i < P; i += T) terms[i] = d_ncr(v.Ntot[i], v.alt[i] + v.alt[P + i]);
	__syncthreads();
	if (tid == 0) {
		uint32_t cf = 0, cr = 0;
		int of = 0, orr = 0, mx = 0;
		uint32_t deepest = 0;  // most reads of one strand in one pool: decides the counter width of the tail kernel
		double clra = 0.0;
		v.cumf[0] = v.cumr[0] = 0;
		for (int i = 0; i < P; i++) {
			deepest = max(deepest, max(v.cumf[i + 1], v.cumr[i + 1]));
			cf += v.cumf[i + 1]; v.cumf[i + 1] = cf;
			cr += v.cumr[i + 1]; v.cumr[i + 1] = cr;
			of += v.alt[i]; orr += v.alt[P + i];
			if (v.Ntot[i] > mx) mx = v.Ntot[i];
			clra += terms[i];  // pool order, as the reference
		}
		v.sh->tot[0] = (int)cf; v.sh->tot[1] = (int)cr;
		v.sh->ones[0] = of; v.sh->ones[1] = orr;
		v.sh->maxN = mx;
		v.sh->clra = clra;
		atomicMax(prm.q.counters + 11, (int)deepest);
		// most balls one pool can receive on one strand: must fit the 16-bit occupancy fields of the head kernel
		v.sh->maxball = (int)min(deepest, (uint32_t)max(of, orr));
		atomicMax(prm.q.counters + 13, v.sh->maxball);
		// how a permutation keeps its occupancy: a column of P words per thread, or — when the table has many pools and few
		// balls — a hash table of at least twice as many entries as balls; whichever lets more threads run
		// the narrowest words whose two fields hold the most balls a pool can receive on one strand: the narrower the column,
		// the more permutations fit the occupancy area
		const int occ_bytes = v.sh->maxball <= 15 ? 1 : (v.sh->maxball <= 255 ? 2 : 4);
		v.sh->occ_bytes = occ_bytes;
		int logc = 4;
		while ((1 << logc) < 2 * (of + orr) && logc < 20) logc++;
		const int act_hash = v.area / (4 << logc), act_col = v.area / (P * occ_bytes);
		const bool hash = P >= kPermHashPools && v.sh->maxball <= 1023 && act_hash > act_col;
		int act = min(T, hash ? act_hash : act_col);
		if (act >= 32) act &= ~31;
		v.sh->hash_log = hash ? logc : 0;
		v.sh->active = max(act, 1);
	}
	__syncthreads();
	const int maxN = v.sh->maxN;
	if (maxN + 2 <= v.l10cap) for (int x = tid; x <= maxN + 1; x += T) v.L10[x] = x > 0 ? log10((double)x) : 0.0;
	__syncthreads();
}

// One permutation (thread-local): place the forward and reverse balls, accumulate the statistic incrementally
// (contables.c:312-323) and compare with the observed one.
// MANY: the many-pool geometry (hash tables, a per-table number of active threads); otherwise every thread of the block runs
// a permutation on a column of OccT words
template <typename OccT, bool MANY>
__device__ __forceinline__ int stranded_hit(const EngineParams& prm, const StrandedView& v, int s, int slot, uint32_t k)
{
	constexpr int HB = sizeof(OccT) * 4;            // bits per strand counter
	constexpr uint32_t HM = (1u << HB) - 1u;
	const int P = prm.c.P, T = MANY ? v.sh->active : (int)blockDim.x;
	const bool l10_smem = v.sh->maxN + 2 <= v.l10cap;
	const double* L10 = v.L10;
	const int32_t* Ntot = v.Ntot;
	double stat = 0.0;
	Philox rng;
	auto step = [&](int i, int cb) {   // the statistic's increment when pool i, holding cb balls, receives one more (contables.c:316,322)
		const int N = Ntot[i];
		stat += l10_smem ? L10[N - cb] - L10[cb + 1] : log10((double)(N - cb)) - log10((double)(cb + 1));
	};
	if (MANY && v.sh->hash_log) {
		const int logc = v.sh->hash_log;
		const uint32_t mask = (1u << logc) - 1u;
		uint32_t* tab = reinterpret_cast<uint32_t*>(v.occ) + threadIdx.x;
		for (uint32_t e = 0; e <= mask; e++) tab[(size_t)e * T] = 0;
		// slot of pool i: its entry, or the empty slot where it would go
		auto find = [&](int i, uint32_t& at) -> uint32_t {
			uint32_t h = ((uint32_t)(i + 1) * 0x9E3779B1u) >> (32 - logc);
			for (;;) {
				const uint32_t e = tab[(size_t)h * T];
				if (e == 0 || (e >> 20) == (uint32_t)(i + 1)) { at = h; return e; }
				h = (h + 1) & mask;
			}
		};
		uint32_t at;
		rng.init(prm.c.rng_seed, prm.b.tid[s], prm.b.pos[s], slot, k, 0);
		place_balls(rng, v.cumf, P, (uint32_t)v.sh->tot[0], v.sh->ones[0],
			[&](int i) { return find(i, at) & 1023u; },
			[&](int i) {
				const uint32_t e = find(i, at);
				step(i, (int)(e & 1023u) + (int)((e >> 10) & 1023u));
				tab[(size_t)at * T] = (e ? e : ((uint32_t)(i + 1) << 20)) + 1u;
			});
		rng.init(prm.c.rng_seed, prm.b.tid[s], prm.b.pos[s], slot, k, 1);
		place_balls(rng, v.cumr, P, (uint32_t)v.sh->tot[1], v.sh->ones[1],
			[&](int i) { return (find(i, at) >> 10) & 1023u; },
			[&](int i) {
				const uint32_t e = find(i, at);
				step(i, (int)(e & 1023u) + (int)((e >> 10) & 1023u));
				tab[(size_t)at * T] = (e ? e : ((uint32_t)(i + 1) << 20)) + (1u << 10);
			});
		return stat <= v.sh->clra + kCtEps ? 1 : 0;
	}
	OccT* myocc = reinterpret_cast<OccT*>(v.occ) + threadIdx.x;
	for (int i = 0; i < P; i++) myocc[(size_t)i * T] = 0;
	rng.init(prm.c.rng_seed, prm.b.tid[s], prm.b.pos[s], slot, k, 0);
	place_balls(rng, v.cumf, P, (uint32_t)v.sh->tot[0], v.sh->ones[0],
		[&](int i) { return (uint32_t)myocc[(size_t)i * T] & HM; },
		[&](int i) {
			const uint32_t w = myocc[(size_t)i * T];
			step(i, (int)(w & HM) + (int)(w >> HB));
			myocc[(size_t)i * T] = (OccT)(w + 1u);
		});
	rng.init(prm.c.rng_seed, prm.b.tid[s], prm.b.pos[s], slot, k, 1);
	place_balls(rng, v.cumr, P, (uint32_t)v.sh->tot[1], v.sh->ones[1],
		[&](int i) { return (uint32_t)myocc[(size_t)i * T] >> HB; },
		[&](int i) {
			const uint32_t w = myocc[(size_t)i * T];
			step(i, (int)(w & HM) + (int)(w >> HB));
			myocc[(size_t)i * T] = (OccT)(w + (1u << HB));
		});
	return stat <= v.sh->clra + kCtEps ? 1 : 0;
}

__device__ __forceinline__ void stranded_store(const EngineParams& prm, int o, int kk, int hh)
{
	prm.o.ctpval[o] = log10((double)(hh + 1)) - log10((double)(kk + 1));  // contables.c:336
	prm.o.perm_k[o] = kk;
	prm.o.perm_hits[o] = hh;
}

constexpr int kPermHead = 1024;   // permutations run by the head kernel (covers the k <= 1000 clause of the stop rule)
constexpr int kPermRec = 12;      // ints per chunk record: [0] hits in the chunk, [1] skipped flag, [2..11] first ten hit indices

// long-running tasks handed from the head kernel to the chunked tail kernel
struct PermTail {
	int32_t* task;       // [n] site*5+slot
	int32_t* hits_head;  // [n] hits after the head permutations
	int32_t* total;      // [n] hits found so far (head + finished chunks), drives the skip of later chunks
	int32_t* rec;        // [n][nchunks][kPermRec]
	int32_t* counters;   // [0] number of tail tasks, [1] work-item head
	int chunk;           // permutations per chunk (multiple of the block size)
	int nchunks;
};

// Head: one CTA per task, the first kPermHead permutations with the reference's sequential stop rule
// (contables.c:327-328) evaluated on prefix sums in permutation order.  Tasks that neither stopped nor exhausted
// their permutations go to the tail queue.
template <bool MANY>
__global__ void __launch_bounds__(256, MANY ? 3 : 4) k_permute_stranded(EngineParams prm, int l10cap, int occ_area, PermTail tail, const int32_t* __restrict__ exact_done)
{
	extern __shared__ __align__(16) unsigned char smem_raw[];
	const DevCfg& c = prm.c;
	const int T = blockDim.x, tid = threadIdx.x;
	const StrandedView v = stranded_view(smem_raw, c.P, l10cap, occ_area);
	PermShared* sh = v.sh;
	for (;;) {
		__syncthreads();
		if (tid == 0) sh->task = atomicAdd(prm.q.counters + 2, 1);
		__syncthreads();
		const int task = sh->task;
		if (task >= prm.q.counters[0]) break;
		if (exact_done && exact_done[task]) continue;  // small table: the exact test already wrote the p-value
		const int o = prm.q.perm_tasks[task], s = o / kSlots, slot = o - s * kSlots;
		stranded_setup(prm, v, o);
		if (v.sh->maxball > 65535) {  // does not fit the 16-bit occupancy fields: the host reports CRISPGPU_ERR_UNSUPPORTED (counters[13])
			if (tid == 0) { prm.o.ctpval[o] = 0.0; prm.o.perm_k[o] = 0; prm.o.perm_hits[o] = 0; }
			continue;
		}
		const int maxiter = c.max_iter;
		const int head = maxiter < kPermHead ? maxiter : kPermHead;
		int stop_k = 0, stop_hits = 0, hits_before = 0;
		const int act = MANY ? sh->active : T;   // threads that run permutations
		for (int base = 0; base < head; base += act) {
			const int k = base + tid + 1;
			const bool mine = tid < act && k <= head;
			int hit = 0;
			if (mine) {
				if (!MANY) hit = stranded_hit<uint32_t, false>(prm, v, s, slot, (uint32_t)k);
				else hit = sh->occ_bytes == 1 ? stranded_hit<uint8_t, true>(prm, v, s, slot, (uint32_t)k)
				         : sh->occ_bytes == 2 ? stranded_hit<uint16_t, true>(prm, v, s, slot, (uint32_t)k) : stranded_hit<uint32_t, true>(prm, v, s, slot, (uint32_t)k);
			}
			int total;
			const int H = hits_before + block_prefix(hit, sh->warp_hits, total);
			const bool stop = mine && (H >= 10 || (k <= 1000 && H >= 5));
			const unsigned sb = __ballot_sync(0xffffffffu, stop);
			if ((tid & 31) == 0) sh->warp_first[tid >> 5] = sb ? (__ffs(sb) - 1) : -1;
			__syncthreads();
			int fw = -1;
			for (int w = 0; w < (T >> 5); w++) if (sh->warp_first[w] >= 0) { fw = w; break; }
			if (fw >= 0) {
				if ((tid >> 5) == fw && (tid & 31) == sh->warp_first[fw]) { sh->stop_k = k; sh->stop_hits = H; }
				__syncthreads();
				stop_k = sh->stop_k;
				stop_hits = sh->stop_hits;
				break;
			}
			hits_before += total;
			__syncthreads();
		}
		if (tid == 0) {
			if (stop_k > 0) stranded_store(prm, o, stop_k, stop_hits);
			else if (maxiter <= head) stranded_store(prm, o, maxiter + 1, hits_before);  // loop ran out: k = maxiter+1
			else {
				const int h = atomicAdd(tail.counters + 0, 1);
				tail.task[h] = o;
				tail.hits_head[h] = hits_before;
				tail.total[h] = hits_before;
			}
		}
	}
}

// Tail: work items = (chunk, task), chunk-major so that every earlier chunk of a task has started before a later one.
// A chunk records its hit count and the indices of its first ten hits; it is skipped when ten hits are already known
// (they all lie in earlier chunks, so the reference's loop would have broken before this chunk).
template <typename OccT, bool MANY>   // MANY: the width of the occupancy words is chosen per table at run time
__global__ void __launch_bounds__(256, MANY ? 3 : 4) k_permute_stranded_tail(EngineParams prm, int l10cap, int occ_area, PermTail tail)
{
	extern __shared__ __align__(16) unsigned char smem_raw[];
	const DevCfg& c = prm.c;
	const int tid = threadIdx.x;
	const StrandedView v = stranded_view(smem_raw, c.P, l10cap, occ_area);
	PermShared* sh = v.sh;
	const int ntask = tail.counters[0];
	const long long nitems = (long long)ntask * tail.nchunks;
	__shared__ int firsts[10];
	for (;;) {
		__syncthreads();
		if (tid == 0) {
			// one thread decides whether the chunk is skipped: the running total changes under concurrent chunks, and the
			// whole block must take the same branch around the barriers below
			const int it = atomicAdd(tail.counters + 1, 1);
			sh->task = it;
			if (it < nitems) {
				const int ch = it / ntask, hh = it - ch * ntask;
				sh->stop_k = (kPermHead + ch * tail.chunk >= c.max_iter || *((volatile int32_t*)(tail.total + hh)) >= 10) ? 1 : 0;
			}
		}
		__syncthreads();
		const int item = sh->task;
		if (item >= nitems) break;
		const int chunk = item / ntask, h = item - chunk * ntask;
		int32_t* rec = tail.rec + ((size_t)h * tail.nchunks + chunk) * kPermRec;
		const int kbeg = kPermHead + chunk * tail.chunk, kend = min(c.max_iter, kbeg + tail.chunk);  // permutations (kbeg, kend]
		if (sh->stop_k) {
			if (tid == 0) { rec[0] = 0; rec[1] = 1; }
			continue;
		}
		const int o = tail.task[h], s = o / kSlots, slot = o - s * kSlots;
		stranded_setup(prm, v, o);
		int found = 0;
		const int act = MANY ? sh->active : (int)blockDim.x;
		for (int base = kbeg; base < kend; base += act) {
			const int k = base + tid + 1;
			int hit = 0;
			if (tid < act && k <= kend) {
				if (!MANY) hit = stranded_hit<OccT, false>(prm, v, s, slot, (uint32_t)k);
				else hit = sh->occ_bytes == 1 ? stranded_hit<uint8_t, true>(prm, v, s, slot, (uint32_t)k)
				         : sh->occ_bytes == 2 ? stranded_hit<uint16_t, true>(prm, v, s, slot, (uint32_t)k) : stranded_hit<uint32_t, true>(prm, v, s, slot, (uint32_t)k);
			}
			int total;
			const int idx = found + block_prefix(hit, sh->warp_hits, total);  // 1-based rank of this hit within the chunk
			if (hit && idx <= 10) firsts[idx - 1] = k;
			found += total;
			__syncthreads();
		}
		if (tid == 0) {
			rec[0] = found;
			rec[1] = 0;
			for (int q = 0; q < 10 && q < found; q++) rec[2 + q] = firsts[q];
			atomicAdd(tail.total + h, found);
		}
	}
}

// Resolve: walk each tail task's chunk records in permutation order to the tenth hit (the break of contables.c:327).
__global__ void k_permute_stranded_resolve(EngineParams prm, PermTail tail)
{
	const int ntask = tail.counters[0];
	for (int h = blockIdx.x * blockDim.x + threadIdx.x; h < ntask; h += gridDim.x * blockDim.x) {
		int cum = tail.hits_head[h], stop_k = 0;
		for (int ch = 0; ch < tail.nchunks && stop_k == 0; ch++) {
			const int32_t* rec = tail.rec + ((size_t)h * tail.nchunks + ch) * kPermRec;
			if (rec[1]) break;  // skipped: ten hits were already known
			const int cnt = rec[0];
			for (int q = 0; q < cnt && q < 10; q++) {
				if (++cum >= 10) { stop_k = rec[2 + q]; break; }
			}
			if (stop_k == 0 && cnt > 10) cum += cnt - 10;  // cannot happen (ten recorded hits already stop the walk)
		}
		if (stop_k > 0) stranded_store(prm, tail.task[h], stop_k, 10);
		else stranded_store(prm, tail.task[h], prm.c.max_iter + 1, cum);
	}
}

// log10 C(n,r): table lookup (bit-identical to the host's libm evaluation) when n is inside the table
__device__ __forceinline__ double ncr_lookup(const double* tab, int n, int r)
{
	if (r == 0 || n == r) return 0.0;
	if (2 * r > n) r = n - r;
	if (tab && n < kNcrTabN) return __ldg(tab + (size_t)n * (kNcrTabN / 2 + 1) + r);
	return d_ncr(n, r);
}

// ---- low-coverage test (pvalue_lowcoverage, FET/lowcovFET.c:23-149) -----------------------------------------------
// Three strata (forward, reverse, bidirectional); statistic = sum over pools that received balls of log10 C(N_i, r_i),
// summed in pool order; two-sided counters without epsilon (:132-133); stop rule of :135.
struct LowcovView {
	PermShared* sh;
	uint32_t* cum;    // [3][P+1]
	int32_t* Ntot;    // [P]
	int32_t* alt;     // [3][P]
	double* terms;    // [P]
	uint32_t* touched;  // [ceil(P/32)][T] bitmap of pools holding balls in the current permutation
	unsigned char* occ; // [P][T] of OccT: three ball counts (f, r, b), 10 bits each in 32-bit words or 5 bits each in 16-bit words
	                    // (when no pool holds more than 31 reads per stratum); all zero between permutations
	int words;
};

__device__ __forceinline__ LowcovView lowcov_view(unsigned char* smem_raw, int P, int T)
{
	LowcovView v;
	v.sh = reinterpret_cast<PermShared*>(smem_raw);
	size_t off = (sizeof(PermShared) + 15) & ~size_t(15);
	v.cum = reinterpret_cast<uint32_t*>(smem_raw + off); off += sizeof(uint32_t) * 3 * (P + 1);
	v.Ntot = reinterpret_cast<int32_t*>(smem_raw + off); off += sizeof(int32_t) * P;
	v.alt = reinterpret_cast<int32_t*>(smem_raw + off); off += sizeof(int32_t) * 3 * P;
	off = (off + 15) & ~size_t(15);
	v.terms = reinterpret_cast<double*>(smem_raw + off); off += sizeof(double) * P;
	v.words = (P + 31) >> 5;
	v.touched = reinterpret_cast<uint32_t*>(smem_raw + off); off += sizeof(uint32_t) * v.words * T;
	v.occ = smem_raw + off;
	return v;
}

// Block-wide set-up; returns false (after writing the result) when the table has fewer than two alt reads (:42).
__device__ inline bool lowcov_setup(const EngineParams& prm, const LowcovView& v, int o, int occ_bytes)
{
	const DevBatch& b = prm.b;
	const DevCfg& c = prm.c;
	const int P = c.P, T = blockDim.x, tid = threadIdx.x;
	const int s = o / kSlots, slot = o - s * kSlots;
	int a1, a2, a3;
	slot_alleles(b.maxvec_al + s * 8, slot, a1, a2, a3);
	const int refbase = b.refbase[s];
	const int row = slot_amb_row(b, s, slot, a2, a3);
	const int altal = (a1 == refbase || a2 != refbase) ? a2 : a1;
	for (int i = tid; i < P; i += T) {
		const int32_t* r = b.indcounts + ((size_t)s * P + i) * kNC;
		int dd[3] = {0, 0, 0};
		if (row >= 0) { const int32_t* d = b.amb_dec + ((size_t)row * P + i) * 4; dd[0] = d[0]; dd[1] = d[1]; dd[2] = d[2] + d[3]; }
		const int nref3 = (a1 == refbase) + (a2 == refbase) + (a3 == refbase), nref2 = (a1 == refbase) + (a2 == refbase);
		const int n0 = r[a1] + r[a2] + r[a3] - nref3 * dd[0], n1 = r[a1 + 8] + r[a2 + 8] + r[a3 + 8] - nref3 * dd[1];
		const int n2 = r[a1 + 16] + r[a2 + 16] - nref2 * dd[2];  // bidirectional column excludes allele3 (allelecounts.c:228)
		const int nc[3] = {n0 < 0 ? 0 : n0, n1 < 0 ? 0 : n1, n2 < 0 ? 0 : n2};  // see stranded_setup: no-ops on a valid batch
		v.cum[0 * (P + 1) + i + 1] = nc[0]; v.cum[1 * (P + 1) + i + 1] = nc[1]; v.cum[2 * (P + 1) + i + 1] = nc[2];
		v.Ntot[i] = nc[0] + nc[1] + nc[2];
		for (int k = 0; k < 3; k++) v.alt[k * P + i] = min(max(r[altal + 8 * k] - (altal == refbase ? dd[k] : 0), 0), nc[k]);
		const int R = v.alt[i] + v.alt[P + i] + v.alt[2 * P + i];
		v.terms[i] = R > 0 ? ncr_lookup(c.ncr_tab, nc[0] + nc[1] + nc[2], R) : 0.0;
	}
	for (int k = tid; k < (int)((size_t)P * T * occ_bytes / 4); k += T) reinterpret_cast<uint32_t*>(v.occ)[k] = 0;
	for (int k = tid; k < v.words * T; k += T) v.touched[k] = 0;
	__syncthreads();
	if (tid == 0) {
		double clra = 0.0;
		uint32_t deepest = 0;  // most reads of one stratum in one pool: decides the counter width of the tail kernel
		uint32_t maxball = 0;  // most balls one pool can receive in one stratum: must fit the 10-bit fields of the head kernel
		for (int k = 0; k < 3; k++) {
			uint32_t cc = 0, dk = 0; int on = 0;
			v.cum[k * (P + 1)] = 0;
			for (int i = 0; i < P; i++) { dk = max(dk, v.cum[k * (P + 1) + i + 1]); cc += v.cum[k * (P + 1) + i + 1]; v.cum[k * (P + 1) + i + 1] = cc; on += v.alt[k * P + i]; }
			v.sh->tot[k] = (int)cc; v.sh->ones[k] = on;
			deepest = max(deepest, dk);
			maxball = max(maxball, min(dk, (uint32_t)on));
		}
		v.sh->maxball = (int)maxball;
		atomicMax(prm.q.counters + 13, (int)maxball);
		for (int i = 0; i < P; i++) if (v.alt[i] + v.alt[P + i] + v.alt[2 * P + i] > 0) clra += v.terms[i];
		v.sh->clra = clra;
		atomicMax(prm.q.counters + 11, (int)deepest);
	}
	__syncthreads();
	// fewer than two alt reads (:42) — or a table whose ball counts do not fit the 10-bit fields of the head kernel: the host
	// reports CRISPGPU_ERR_UNSUPPORTED for the batch (crispgpu.cu, counters[13]); nothing is sampled from a corrupt layout
	if (v.sh->ones[0] + v.sh->ones[1] + v.sh->ones[2] < 2 || v.sh->maxball > 1023) {
		if (tid == 0) { prm.o.ctpval[o] = 0.0; prm.o.perm_k[o] = 0; prm.o.perm_hits[o] = 0; }
		return false;
	}
	return true;
}

// One permutation: bit 0 = statistic <= observed, bit 1 = statistic >= observed.  Work is proportional to the number of
// balls: pools that received balls are tracked in a bitmap, walked in ascending pool order and cleared on the way.
template <typename OccT>
__device__ __forceinline__ int lowcov_hit(const EngineParams& prm, const LowcovView& v, int s, int slot, uint32_t k)
{
	constexpr int FB = sizeof(OccT) == 4 ? 10 : 5;  // bits per stratum counter
	constexpr uint32_t FM = (1u << FB) - 1u;
	const int P = prm.c.P, T = blockDim.x;
	OccT* myocc = reinterpret_cast<OccT*>(v.occ) + threadIdx.x;
	uint32_t* mybits = v.touched + threadIdx.x;
	for (int st = 0; st < 3; st++) {
		Philox rng;
		rng.init(prm.c.rng_seed, prm.b.tid[s], prm.b.pos[s], slot, k, (uint32_t)st);
		const int sft = FB * st;
		place_balls(rng, v.cum + st * (P + 1), P, (uint32_t)v.sh->tot[st], v.sh->ones[st],
			[&](int i) { return ((uint32_t)myocc[(size_t)i * T] >> sft) & FM; },
			[&](int i) {
				myocc[(size_t)i * T] = (OccT)(myocc[(size_t)i * T] + (1u << sft));
				mybits[(size_t)(i >> 5) * T] |= 1u << (i & 31);
			});
	}
	double stat = 0.0;
	for (int w = 0; w < v.words; w++) {
		uint32_t bits = mybits[(size_t)w * T];
		if (!bits) continue;
		mybits[(size_t)w * T] = 0;
		while (bits) {
			const int i = (w << 5) + __ffs(bits) - 1;
			bits &= bits - 1;
			const uint32_t x = myocc[(size_t)i * T];
			myocc[(size_t)i * T] = 0;
			const int r = (int)(x & FM) + (int)((x >> FB) & FM) + (int)((x >> (2 * FB)) & FM);
			stat += ncr_lookup(prm.c.ncr_tab, v.Ntot[i], r);
		}
	}
	const double clra = v.sh->clra;
	return (stat <= clra ? 1 : 0) | (stat >= clra ? 2 : 0);
}

__device__ __forceinline__ void lowcov_store(const EngineParams& prm, int o, int kk, int lo, int hi)
{
	const int use = hi < lo ? hi : lo;  // lowcovFET.c:139-144
	prm.o.ctpval[o] = log10((double)use + 1) - log10((double)(kk + 1));
	prm.o.perm_k[o] = kk;
	prm.o.perm_hits[o] = use;
}

constexpr int kLowRec = 24;  // ints per chunk record: [0] lo hits, [1] hi hits, [2] skipped, [3] pad, [4..13] first ten lo indices, [14..23] hi

// Head: the first kPermHead permutations of every task with the sequential stop rule (covers its k <= 100 clause).
__global__ void k_permute_lowcov(EngineParams prm, PermTail tail)
{
	extern __shared__ __align__(16) unsigned char smem_raw[];
	const DevCfg& c = prm.c;
	const int T = blockDim.x, tid = threadIdx.x;
	const LowcovView v = lowcov_view(smem_raw, c.P, T);
	PermShared* sh = v.sh;
	for (;;) {
		__syncthreads();
		if (tid == 0) sh->task = atomicAdd(prm.q.counters + 2, 1);
		__syncthreads();
		const int task = sh->task;
		if (task >= prm.q.counters[0]) break;
		const int o = prm.q.perm_tasks[task], s = o / kSlots, slot = o - s * kSlots;
		if (!lowcov_setup(prm, v, o, 4)) continue;
		const int maxiter = c.max_iter;
		const int head = maxiter < kPermHead ? maxiter : kPermHead;
		int lo_before = 0, hi_before = 0, stop_k = 0, stop_lo = 0, stop_hi = 0;
		for (int base = 0; base < head; base += T) {
			const int k = base + tid + 1;
			const int f = k <= head ? lowcov_hit<uint32_t>(prm, v, s, slot, (uint32_t)k) : 0;
			int tl, th;
			const int Hlo = lo_before + block_prefix(f & 1, sh->warp_hits, tl);
			const int Hhi = hi_before + block_prefix((f >> 1) & 1, sh->warp_hits2, th);
			const bool stop = k <= head && ((Hlo >= 10 && Hhi >= 10) || (k <= 100 && Hlo >= 5 && Hhi >= 5));
			const unsigned sb = __ballot_sync(0xffffffffu, stop);
			if ((tid & 31) == 0) sh->warp_first[tid >> 5] = sb ? (__ffs(sb) - 1) : -1;
			__syncthreads();
			int fw = -1;
			for (int w = 0; w < (T >> 5); w++) if (sh->warp_first[w] >= 0) { fw = w; break; }
			if (fw >= 0) {
				if ((tid >> 5) == fw && (tid & 31) == sh->warp_first[fw]) { sh->stop_k = k; sh->stop_hits = Hlo; sh->hits_before = Hhi; }
				__syncthreads();
				stop_k = sh->stop_k; stop_lo = sh->stop_hits; stop_hi = sh->hits_before;
				break;
			}
			lo_before += tl;
			hi_before += th;
			__syncthreads();
		}
		if (tid == 0) {
			if (stop_k > 0) lowcov_store(prm, o, stop_k, stop_lo, stop_hi);
			else if (maxiter <= head) lowcov_store(prm, o, maxiter + 1, lo_before, hi_before);
			else {
				const int h = atomicAdd(tail.counters + 0, 1);
				tail.task[h] = o;
				tail.hits_head[2 * h] = lo_before;
				tail.hits_head[2 * h + 1] = hi_before;
				tail.total[2 * h] = lo_before;
				tail.total[2 * h + 1] = hi_before;
			}
		}
	}
}

// Tail: (chunk, task) items chunk-major; a chunk is skipped once both counters are known to have reached ten.
template <typename OccT>
__global__ void k_permute_lowcov_tail(EngineParams prm, PermTail tail)
{
	extern __shared__ __align__(16) unsigned char smem_raw[];
	const DevCfg& c = prm.c;
	const int T = blockDim.x, tid = threadIdx.x;
	const LowcovView v = lowcov_view(smem_raw, c.P, T);
	PermShared* sh = v.sh;
	const int ntask = tail.counters[0];
	const long long nitems = (long long)ntask * tail.nchunks;
	__shared__ int firsts[20];
	for (;;) {
		__syncthreads();
		if (tid == 0) {
			// one thread decides the skip for the whole block (the totals change under concurrent chunks)
			const int it = atomicAdd(tail.counters + 1, 1);
			sh->task = it;
			if (it < nitems) {
				const int ch = it / ntask, hh = it - ch * ntask;
				const volatile int32_t* tot = tail.total + 2 * hh;
				sh->stop_k = (kPermHead + ch * tail.chunk >= c.max_iter || (tot[0] >= 10 && tot[1] >= 10)) ? 1 : 0;
			}
		}
		__syncthreads();
		const int item = sh->task;
		if (item >= nitems) break;
		const int chunk = item / ntask, h = item - chunk * ntask;
		int32_t* rec = tail.rec + ((size_t)h * tail.nchunks + chunk) * kLowRec;
		const int kbeg = kPermHead + chunk * tail.chunk, kend = min(c.max_iter, kbeg + tail.chunk);
		if (sh->stop_k) {
			if (tid == 0) { rec[0] = 0; rec[1] = 0; rec[2] = 1; }
			continue;
		}
		const int o = tail.task[h], s = o / kSlots, slot = o - s * kSlots;
		lowcov_setup(prm, v, o, (int)sizeof(OccT));  // cannot fail: the head already accepted the table
		int flo = 0, fhi = 0;
		for (int base = kbeg; base < kend; base += T) {
			const int k = base + tid + 1;
			const int f = k <= kend ? lowcov_hit<OccT>(prm, v, s, slot, (uint32_t)k) : 0;
			int tl, th;
			const int ilo = flo + block_prefix(f & 1, sh->warp_hits, tl);
			const int ihi = fhi + block_prefix((f >> 1) & 1, sh->warp_hits2, th);
			if ((f & 1) && ilo <= 10) firsts[ilo - 1] = k;
			if ((f & 2) && ihi <= 10) firsts[10 + ihi - 1] = k;
			flo += tl;
			fhi += th;
			__syncthreads();
		}
		if (tid == 0) {
			rec[0] = flo; rec[1] = fhi; rec[2] = 0;
			for (int q = 0; q < 10 && q < flo; q++) rec[4 + q] = firsts[q];
			for (int q = 0; q < 10 && q < fhi; q++) rec[14 + q] = firsts[10 + q];
			atomicAdd(tail.total + 2 * h, flo);
			atomicAdd(tail.total + 2 * h + 1, fhi);
		}
	}
}

// Resolve: the loop of lowcovFET.c:73-137 breaks at the first k where both counters have reached ten, i.e. at the later of
// the two tenth hits; the smaller counter is then exactly ten (:139-144).
__global__ void k_permute_lowcov_resolve(EngineParams prm, PermTail tail)
{
	const int ntask = tail.counters[0];
	for (int h = blockIdx.x * blockDim.x + threadIdx.x; h < ntask; h += gridDim.x * blockDim.x) {
		int lo = tail.hits_head[2 * h], hi = tail.hits_head[2 * h + 1], klo = lo >= 10 ? kPermHead : 0, khi = hi >= 10 ? kPermHead : 0;
		for (int ch = 0; ch < tail.nchunks && !(klo && khi); ch++) {
			const int32_t* rec = tail.rec + ((size_t)h * tail.nchunks + ch) * kLowRec;
			if (rec[2]) break;
			for (int q = 0; q < rec[0] && q < 10 && !klo; q++) if (++lo >= 10) klo = rec[4 + q];
			if (!klo && rec[0] > 10) lo += rec[0] - 10;
			for (int q = 0; q < rec[1] && q < 10 && !khi; q++) if (++hi >= 10) khi = rec[14 + q];
			if (!khi && rec[1] > 10) hi += rec[1] - 10;
		}
		if (klo && khi) lowcov_store(prm, tail.task[h], klo > khi ? klo : khi, 10, 10);
		else {
			// ran out of permutations: totals over all chunks
			int tlo = tail.hits_head[2 * h], thi = tail.hits_head[2 * h + 1];
			for (int ch = 0; ch < tail.nchunks; ch++) {
				const int32_t* rec = tail.rec + ((size_t)h * tail.nchunks + ch) * kLowRec;
				if (rec[2]) break;
				tlo += rec[0];
				thi += rec[1];
			}
			lowcov_store(prm, tail.task[h], prm.c.max_iter + 1, tlo, thi);
		}
	}
}

// ---- exact 2 x k test for small tables (exact_mode) ------------------------------------------------------------------
// Specification: compute_pvalue_exact / pvalue_contable_iter, FET/contables_exact.c:62-117 (not part of CRISP.binary).
// p = sum over all tables (a_1..a_k), sum a_i = A, a_i <= N_i, whose log10 probability sum_i log10 C(N_i, a_i) does not
// exceed the observed one, of prod_i C(N_i, a_i) / C(N, A), on the unstranded table (N_i, alt_i).  Gate: fewer than 70 alt
// reads, columns below 2001 reads, 2 <= pools < 8.  The acceptance test is evaluated with the reference's own
// operands in its own order: the budget is reduced column by column (prob - cp) and the last column is accepted when
// its term does not exceed what is left; both term tables come from the host's libm (exact ties with the observed
// table decide whether its own probability is counted).  One CTA per task; threads split the values of the first two
// columns and walk the remaining columns depth-first.
constexpr int kExactMaxN = 2001;  // MAXN, FET/contables.c:10
constexpr int kExactOnes = 70;

__global__ void k_exact_2xk(EngineParams prm, const double* __restrict__ ncrT, const double* __restrict__ clT, int32_t* __restrict__ done)
{
	const DevBatch& b = prm.b;
	const DevCfg& c = prm.c;
	const int P = c.P, T = blockDim.x, tid = threadIdx.x;
	__shared__ int N[8], A[8], Rsuf[9], s_ok, s_ones;
	__shared__ double s_cl, s_part[32];
	for (int task = blockIdx.x; task < prm.q.counters[0]; task += gridDim.x) {
		__syncthreads();
		const int o = prm.q.perm_tasks[task], s = o / kSlots, slot = o - s * kSlots;
		if (tid < P) {
			int a1, a2, a3;
			slot_alleles(b.maxvec_al + s * 8, slot, a1, a2, a3);
			const int refbase = b.refbase[s];
			const int row = slot_amb_row(b, s, slot, a2, a3);
			const int altal = (a1 == refbase || a2 != refbase) ? a2 : a1;
			const int32_t* r = b.indcounts + ((size_t)s * P + tid) * kNC;
			int dd = 0;
			if (row >= 0) { const int32_t* d = b.amb_dec + ((size_t)row * P + tid) * 4; dd = d[0] + d[1] + d[2] + d[3]; }
			// variant->ctable[i][0..1] (allelecounts.c:214-220): N = c0 + c1, alt = c1 (or c0 when the reference allele is second)
			const int c0 = r[a1] + r[a1 + 8] + r[a1 + 16] + r[a3] + r[a3 + 8] + r[a3 + 16] - ((a1 == refbase) + (a3 == refbase)) * dd;
			const int c1 = r[a2] + r[a2 + 8] + r[a2 + 16] - (a2 == refbase ? dd : 0);
			N[tid] = c0 + c1;
			A[tid] = altal == a2 ? c1 : c0;
		}
		__syncthreads();
		if (tid == 0) {
			int ones = 0, maxcol = 0, run = 0;
			for (int i = 0; i < P; i++) { ones += A[i]; if (N[i] > maxcol) maxcol = N[i]; }
			Rsuf[P] = 0;
			for (int i = P - 1; i >= 0; i--) { run += N[i]; Rsuf[i] = run; }
			const int ok = ones < kExactOnes && maxcol < kExactMaxN && P >= 2 && P < 8 && ones >= 0;
			double cl = 0.0;
			if (ok) for (int i = 0; i < P; i++) cl += clT[(size_t)N[i] * kExactOnes + A[i]];  // ncr(N_i, A_i), pool order
			s_ok = ok; s_ones = ones; s_cl = cl;
		}
		__syncthreads();
		if (!s_ok) { if (tid == 0) done[task] = 0; continue; }
		const int ones = s_ones;
		const double cl = s_cl;
		double S = 0.0;
		if (ones <= Rsuf[0]) {
			const int m0 = min(N[0], ones) + 1, m1 = P > 2 ? min(N[1], ones) + 1 : 1;
			for (int idx = tid; idx < m0 * m1; idx += T) {
				const int t0 = idx / m1, t1 = idx - t0 * m1;
				int lvl = 1, start = 1;
				int tt[8], Ar[8];
				double pr[8];
				const double cp0 = N[0] > 0 ? ncrT[(size_t)N[0] * kExactOnes + t0] : 0.0;
				Ar[1] = ones - t0;
				pr[1] = cl - cp0;
				if (P > 2) {
					if (Ar[1] > Rsuf[1] || t1 > Ar[1]) continue;
					const double cp1 = N[1] > 0 ? ncrT[(size_t)N[1] * kExactOnes + t1] : 0.0;
					Ar[2] = Ar[1] - t1;
					pr[2] = pr[1] - cp1;
					lvl = start = 2;
				}
				tt[lvl] = 0;
				for (;;) {
					if (lvl == P - 1) {  // last column takes what is left
						const int a = Ar[lvl];
						if (a <= N[lvl]) {
							const double cp = N[lvl] > 0 ? ncrT[(size_t)N[lvl] * kExactOnes + a] : 0.0;
							const double x = cp - pr[lvl];
							if (!(cp > pr[lvl]) && x > -290.0) S += exp10_lse(x);
						}
						lvl--;
						if (lvl < start) break;
						tt[lvl]++;
						continue;
					}
					if (Ar[lvl] > Rsuf[lvl] || tt[lvl] > N[lvl] || tt[lvl] > Ar[lvl]) {
						lvl--;
						if (lvl < start) break;
						tt[lvl]++;
						continue;
					}
					const double cp = N[lvl] > 0 ? ncrT[(size_t)N[lvl] * kExactOnes + tt[lvl]] : 0.0;
					Ar[lvl + 1] = Ar[lvl] - tt[lvl];
					pr[lvl + 1] = pr[lvl] - cp;
					tt[lvl + 1] = 0;
					lvl++;
				}
			}
		}
		S = warp_sum(S);
		if ((tid & 31) == 0) s_part[tid >> 5] = S;
		__syncthreads();
		if (tid == 0) {
			double tot = 0.0;
			for (int w = 0; w < (T >> 5); w++) tot += s_part[w];
			const double p = tot > 0 ? log10(tot) + cl : -1.0;  // compute_pvalue_exact returns -1 when no table qualifies
			prm.o.ctpval[o] = p - d_ncr(Rsuf[0], ones);
			prm.o.perm_k[o] = 0;
			prm.o.perm_hits[o] = -1;
			done[task] = 1;
		}
	}
}

}  // namespace crisp
