// k_permute.cuh — Monte-Carlo contingency-table p-values: one CTA per (site, slot) task, one thread per permutation.
//
//   k_permute_stranded   pvalue_contable_iter_stranded  FET/contables.c:261-339  (pool size > 2)
//   k_permute_lowcov     pvalue_lowcoverage             FET/lowcovFET.c:23-149   (pool size == 2)
//
// Null law: the alt reads of each strand are placed without replacement among that strand's read slots
// (multivariate hypergeometric per strand, strands independent).  The reference shuffles a slot->pool index with
// drand48; here every permutation k owns a Philox4x32-10 stream keyed by (seed, tid, pos, slot) with counter
// (k, block, stratum), so the result does not depend on batching, grid size or GPU count.  A ball is placed by
// drawing a uniform slot, finding its pool in the shared-memory cumulative marginals, and rejecting the draw when it
// lands on one of the pool's already-occupied slots (slots of a pool are exchangeable).  The reference's sequential
// stopping rule (hits >= 10, or >= 5 within the first 1000 permutations) is evaluated on the prefix sums of the
// hit flags in permutation order, so the reported (hits, k) is exactly what a serial run of the same streams gives.
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
	double clra;
};

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
	uint32_t* occ;  // [P][T]: low 16 bits forward balls, high 16 reverse balls
	int l10cap;
};

__device__ __forceinline__ StrandedView stranded_view(unsigned char* smem_raw, int P, int l10cap)
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
	v.occ = reinterpret_cast<uint32_t*>(smem_raw + off);
	v.l10cap = l10cap;
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
		v.cumf[i + 1] = nf;
		v.cumr[i + 1] = nr;
		v.Ntot[i] = nf + nr;
		v.alt[i] = r[altal] - (altal == refbase ? dd[0] : 0);
		v.alt[P + i] = r[altal + 8] - (altal == refbase ? dd[1] : 0);
	}
	__syncthreads();
	double* terms = v.L10;  // borrowed before the log table is built
	for (int i = tid; i < P; i += T) terms[i] = d_ncr(v.Ntot[i], v.alt[i] + v.alt[P + i]);
	__syncthreads();
	if (tid == 0) {
		uint32_t cf = 0, cr = 0;
		int of = 0, orr = 0, mx = 0;
		double clra = 0.0;
		v.cumf[0] = v.cumr[0] = 0;
		for (int i = 0; i < P; i++) {
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
	}
	__syncthreads();
	const int maxN = v.sh->maxN;
	if (maxN + 2 <= v.l10cap) for (int x = tid; x <= maxN + 1; x += T) v.L10[x] = x > 0 ? log10((double)x) : 0.0;
	__syncthreads();
}

// One permutation (thread-local): place the forward and reverse balls, accumulate the statistic incrementally
// (contables.c:312-323) and compare with the observed one.
__device__ __forceinline__ int stranded_hit(const EngineParams& prm, const StrandedView& v, int s, int slot, uint32_t k)
{
	const int P = prm.c.P, T = blockDim.x;
	uint32_t* myocc = v.occ + threadIdx.x;
	const bool l10_smem = v.sh->maxN + 2 <= v.l10cap;
	const double* L10 = v.L10;
	const int32_t* Ntot = v.Ntot;
	for (int i = 0; i < P; i++) myocc[(size_t)i * T] = 0;
	double stat = 0.0;
	Philox rng;
	rng.init(prm.c.rng_seed, prm.b.tid[s], prm.b.pos[s], slot, k, 0);
	place_balls(rng, v.cumf, P, (uint32_t)v.sh->tot[0], v.sh->ones[0],
		[&](int i) { return myocc[(size_t)i * T] & 0xffffu; },
		[&](int i) {
			const uint32_t w = myocc[(size_t)i * T];
			const int cb = (int)(w & 0xffffu) + (int)(w >> 16);
			const int N = Ntot[i];
			stat += l10_smem ? L10[N - cb] - L10[cb + 1] : log10((double)(N - cb)) - log10((double)(cb + 1));
			myocc[(size_t)i * T] = w + 1u;
		});
	rng.init(prm.c.rng_seed, prm.b.tid[s], prm.b.pos[s], slot, k, 1);
	place_balls(rng, v.cumr, P, (uint32_t)v.sh->tot[1], v.sh->ones[1],
		[&](int i) { return myocc[(size_t)i * T] >> 16; },
		[&](int i) {
			const uint32_t w = myocc[(size_t)i * T];
			const int cb = (int)(w & 0xffffu) + (int)(w >> 16);
			const int N = Ntot[i];
			stat += l10_smem ? L10[N - cb] - L10[cb + 1] : log10((double)(N - cb)) - log10((double)(cb + 1));
			myocc[(size_t)i * T] = w + 0x10000u;
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
__global__ void k_permute_stranded(EngineParams prm, int l10cap, PermTail tail)
{
	extern __shared__ __align__(16) unsigned char smem_raw[];
	const DevCfg& c = prm.c;
	const int T = blockDim.x, tid = threadIdx.x;
	const StrandedView v = stranded_view(smem_raw, c.P, l10cap);
	PermShared* sh = v.sh;
	for (;;) {
		__syncthreads();
		if (tid == 0) sh->task = atomicAdd(prm.q.counters + 2, 1);
		__syncthreads();
		const int task = sh->task;
		if (task >= prm.q.counters[0]) break;
		const int o = prm.q.perm_tasks[task], s = o / kSlots, slot = o - s * kSlots;
		stranded_setup(prm, v, o);
		const int maxiter = c.max_iter;
		const int head = maxiter < kPermHead ? maxiter : kPermHead;
		int stop_k = 0, stop_hits = 0, hits_before = 0;
		for (int base = 0; base < head; base += T) {
			const int k = base + tid + 1;
			const int hit = k <= head ? stranded_hit(prm, v, s, slot, (uint32_t)k) : 0;
			int total;
			const int H = hits_before + block_prefix(hit, sh->warp_hits, total);
			const bool stop = k <= head && (H >= 10 || (k <= 1000 && H >= 5));
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
__global__ void k_permute_stranded_tail(EngineParams prm, int l10cap, PermTail tail)
{
	extern __shared__ __align__(16) unsigned char smem_raw[];
	const DevCfg& c = prm.c;
	const int T = blockDim.x, tid = threadIdx.x;
	const StrandedView v = stranded_view(smem_raw, c.P, l10cap);
	PermShared* sh = v.sh;
	const int ntask = tail.counters[0];
	const long long nitems = (long long)ntask * tail.nchunks;
	__shared__ int firsts[10];
	for (;;) {
		__syncthreads();
		if (tid == 0) sh->task = atomicAdd(tail.counters + 1, 1);
		__syncthreads();
		const int item = sh->task;
		if (item >= nitems) break;
		const int chunk = item / ntask, h = item - chunk * ntask;
		int32_t* rec = tail.rec + ((size_t)h * tail.nchunks + chunk) * kPermRec;
		const int kbeg = kPermHead + chunk * tail.chunk, kend = min(c.max_iter, kbeg + tail.chunk);  // permutations (kbeg, kend]
		if (kbeg >= c.max_iter || *((volatile int32_t*)(tail.total + h)) >= 10) {
			if (tid == 0) { rec[0] = 0; rec[1] = 1; }
			continue;
		}
		const int o = tail.task[h], s = o / kSlots, slot = o - s * kSlots;
		stranded_setup(prm, v, o);
		int found = 0;
		for (int base = kbeg; base < kend; base += T) {
			const int k = base + tid + 1;
			const int hit = k <= kend ? stranded_hit(prm, v, s, slot, (uint32_t)k) : 0;
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

// pvalue_lowcoverage: three strata (forward, reverse, bidirectional), statistic = sum over pools with balls of
// log10 C(N_i, r_i) summed in pool order, two-sided counters without epsilon, stop rule of lowcovFET.c:135.
__global__ void k_permute_lowcov(EngineParams prm)
{
	extern __shared__ __align__(16) unsigned char smem_raw[];
	const DevBatch& b = prm.b;
	const DevCfg& c = prm.c;
	const int P = c.P, T = blockDim.x, tid = threadIdx.x;
	PermShared* sh = reinterpret_cast<PermShared*>(smem_raw);
	size_t off = (sizeof(PermShared) + 15) & ~size_t(15);
	uint32_t* cum = reinterpret_cast<uint32_t*>(smem_raw + off); off += sizeof(uint32_t) * 3 * (P + 1);
	int32_t* Ntot = reinterpret_cast<int32_t*>(smem_raw + off); off += sizeof(int32_t) * P;
	int32_t* alt = reinterpret_cast<int32_t*>(smem_raw + off); off += sizeof(int32_t) * 3 * P;
	off = (off + 15) & ~size_t(15);
	double* terms = reinterpret_cast<double*>(smem_raw + off); off += sizeof(double) * P;
	uint32_t* occ = reinterpret_cast<uint32_t*>(smem_raw + off);  // [P][T]: 3 x 10-bit ball counts (f, r, b)

	for (;;) {
		__syncthreads();
		if (tid == 0) sh->task = atomicAdd(prm.q.counters + 2, 1);
		__syncthreads();
		const int task = sh->task;
		if (task >= prm.q.counters[0]) break;
		const int o = prm.q.perm_tasks[task], s = o / kSlots, slot = o - s * kSlots;
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
			cum[0 * (P + 1) + i + 1] = n0; cum[1 * (P + 1) + i + 1] = n1; cum[2 * (P + 1) + i + 1] = n2;
			Ntot[i] = n0 + n1 + n2;
			for (int k = 0; k < 3; k++) alt[k * P + i] = r[altal + 8 * k] - (altal == refbase ? dd[k] : 0);
			const int R = alt[i] + alt[P + i] + alt[2 * P + i];
			terms[i] = R > 0 ? ncr_lookup(c.ncr_tab, n0 + n1 + n2, R) : 0.0;
		}
		__syncthreads();
		if (tid == 0) {
			double clra = 0.0;
			for (int k = 0; k < 3; k++) {
				uint32_t cc = 0; int on = 0;
				cum[k * (P + 1)] = 0;
				for (int i = 0; i < P; i++) { cc += cum[k * (P + 1) + i + 1]; cum[k * (P + 1) + i + 1] = cc; on += alt[k * P + i]; }
				sh->tot[k] = (int)cc; sh->ones[k] = on;
			}
			for (int i = 0; i < P; i++) if (alt[i] + alt[P + i] + alt[2 * P + i] > 0) clra += terms[i];
			sh->clra = clra;
			sh->stop_k = 0;
		}
		__syncthreads();
		const int tot_ones = sh->ones[0] + sh->ones[1] + sh->ones[2];
		if (tot_ones < 2) {  // lowcovFET.c:42
			if (tid == 0) { prm.o.ctpval[o] = 0.0; prm.o.perm_k[o] = 0; prm.o.perm_hits[o] = 0; }
			continue;
		}
		const double clra = sh->clra;
		const int maxiter = c.max_iter;
		uint32_t* myocc = occ + tid;
		int lo_before = 0, hi_before = 0, stop_k = 0, stop_lo = 0, stop_hi = 0;
		for (int base = 0; base < maxiter; base += T) {
			const int k = base + tid + 1;
			int flo = 0, fhi = 0;
			if (k <= maxiter) {
				for (int i = 0; i < P; i++) myocc[(size_t)i * T] = 0;
				for (int st = 0; st < 3; st++) {
					Philox rng;
					rng.init(c.rng_seed, b.tid[s], b.pos[s], slot, (uint32_t)k, (uint32_t)st);
					const int sft = 10 * st;
					place_balls(rng, cum + st * (P + 1), P, (uint32_t)sh->tot[st], sh->ones[st],
						[&](int i) { return (myocc[(size_t)i * T] >> sft) & 0x3ffu; },
						[&](int i) { myocc[(size_t)i * T] += 1u << sft; });
				}
				double stat = 0.0;
				for (int i = 0; i < P; i++) {
					const uint32_t v = myocc[(size_t)i * T];
					const int r = (int)(v & 0x3ffu) + (int)((v >> 10) & 0x3ffu) + (int)((v >> 20) & 0x3ffu);
					if (r > 0) stat += ncr_lookup(c.ncr_tab, Ntot[i], r);
				}
				flo = stat <= clra ? 1 : 0;
				fhi = stat >= clra ? 1 : 0;
			}
			int tl, th;
			const int Hlo = lo_before + block_prefix(flo, sh->warp_hits, tl);
			const int Hhi = hi_before + block_prefix(fhi, sh->warp_hits2, th);
			const bool stop = k <= maxiter && ((Hlo >= 10 && Hhi >= 10) || (k <= 100 && Hlo >= 5 && Hhi >= 5));
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
			int kk = stop_k, l = stop_lo, h = stop_hi;
			if (kk == 0) { kk = maxiter + 1; l = lo_before; h = hi_before; }
			const int use = h < l ? h : l;  // lowcovFET.c:139-144
			prm.o.ctpval[o] = log10((double)use + 1) - log10((double)(kk + 1));
			prm.o.perm_k[o] = kk;
			prm.o.perm_hits[o] = use;
		}
	}
}

}  // namespace crisp
