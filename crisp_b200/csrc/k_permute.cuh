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
	for (int d = 0; d < ones; d++) {
		for (;;) {
			const uint32_t u = rng.below(M);
			int lo = 0, hi = P;
			while (hi - lo > 1) {
				const int mid = (lo + hi) >> 1;
				if (cum[mid] <= u) lo = mid; else hi = mid;
			}
			if (u - cum[lo] >= (uint32_t)occ(lo)) { bump(lo); break; }
		}
	}
}

__global__ void k_permute_stranded(EngineParams prm, int l10cap)
{
	extern __shared__ __align__(16) unsigned char smem_raw[];
	const DevBatch& b = prm.b;
	const DevCfg& c = prm.c;
	const int P = c.P, T = blockDim.x, tid = threadIdx.x;
	PermShared* sh = reinterpret_cast<PermShared*>(smem_raw);
	size_t off = (sizeof(PermShared) + 15) & ~size_t(15);
	uint32_t* cumf = reinterpret_cast<uint32_t*>(smem_raw + off); off += sizeof(uint32_t) * (P + 1);
	uint32_t* cumr = reinterpret_cast<uint32_t*>(smem_raw + off); off += sizeof(uint32_t) * (P + 1);
	int32_t* Ntot = reinterpret_cast<int32_t*>(smem_raw + off); off += sizeof(int32_t) * P;
	int32_t* alt = reinterpret_cast<int32_t*>(smem_raw + off); off += sizeof(int32_t) * 2 * P;
	off = (off + 15) & ~size_t(15);
	double* L10 = reinterpret_cast<double*>(smem_raw + off); off += sizeof(double) * l10cap;
	uint32_t* occ = reinterpret_cast<uint32_t*>(smem_raw + off);  // [P][T]: low 16 bits forward balls, high 16 reverse balls

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
		// ---- marginals: ctable_stranded columns 0..3 (allelecounts.c:221-227) ----
		for (int i = tid; i < P; i += T) {
			const int32_t* r = b.indcounts + ((size_t)s * P + i) * kNC;
			int dd[2] = {0, 0};
			if (row >= 0) { const int32_t* d = b.amb_dec + ((size_t)row * P + i) * 4; dd[0] = d[0]; dd[1] = d[1]; }
			const int nref = (a1 == refbase) + (a2 == refbase) + (a3 == refbase);
			const int nf = r[a1] + r[a2] + r[a3] - nref * dd[0], nr = r[a1 + 8] + r[a2 + 8] + r[a3 + 8] - nref * dd[1];
			cumf[i + 1] = nf;
			cumr[i + 1] = nr;
			Ntot[i] = nf + nr;
			alt[i] = r[altal] - (altal == refbase ? dd[0] : 0);
			alt[P + i] = r[altal + 8] - (altal == refbase ? dd[1] : 0);
		}
		__syncthreads();
		if (tid == 0) {
			// serial prefix sums and the observed statistic clra = sum_i log10 C(Nf+Nr, af+ar)  (contables.c:296-298)
			uint32_t cf = 0, cr = 0;
			int of = 0, orr = 0, mx = 0;
			double clra = 0.0;
			cumf[0] = cumr[0] = 0;
			for (int i = 0; i < P; i++) {
				cf += cumf[i + 1]; cumf[i + 1] = cf;
				cr += cumr[i + 1]; cumr[i + 1] = cr;
				of += alt[i]; orr += alt[P + i];
				if (Ntot[i] > mx) mx = Ntot[i];
			}
			sh->tot[0] = (int)cf; sh->tot[1] = (int)cr;
			sh->ones[0] = of; sh->ones[1] = orr;
			sh->maxN = mx;
			sh->hits_before = 0;
			sh->stop_k = 0;
			sh->stop_hits = 0;
			sh->clra = clra;
		}
		__syncthreads();
		{
			// clra summed in pool order by one thread after a parallel evaluation of the terms
			double* terms = L10;  // borrowed before the log table is built
			for (int i = tid; i < P; i += T) terms[i] = d_ncr(Ntot[i], alt[i] + alt[P + i]);
			__syncthreads();
			if (tid == 0) { double v = 0.0; for (int i = 0; i < P; i++) v += terms[i]; sh->clra = v; }
			__syncthreads();
		}
		const int maxN = sh->maxN;
		const bool l10_smem = maxN + 2 <= l10cap;
		if (l10_smem) for (int x = tid; x <= maxN + 1; x += T) L10[x] = x > 0 ? log10((double)x) : 0.0;
		__syncthreads();
		const int onesf = sh->ones[0], onesr = sh->ones[1];
		const uint32_t Mf = (uint32_t)sh->tot[0], Mr = (uint32_t)sh->tot[1];
		const double thresh = sh->clra + kCtEps;
		const int maxiter = c.max_iter;
		uint32_t* myocc = occ + tid;

		int stop_k = 0, stop_hits = 0, hits_before = 0;
		for (int base = 0; base < maxiter; base += T) {
			const int k = base + tid + 1;
			int hit = 0;
			if (k <= maxiter) {
				for (int i = 0; i < P; i++) myocc[(size_t)i * T] = 0;
				double stat = 0.0;
				Philox rng;
				rng.init(c.rng_seed, b.tid[s], b.pos[s], slot, (uint32_t)k, 0);
				place_balls(rng, cumf, P, Mf, onesf,
					[&](int i) { return myocc[(size_t)i * T] & 0xffffu; },
					[&](int i) {
						const uint32_t v = myocc[(size_t)i * T];
						const int cb = (int)(v & 0xffffu) + (int)(v >> 16);
						const int N = Ntot[i];
						stat += l10_smem ? L10[N - cb] - L10[cb + 1] : log10((double)(N - cb)) - log10((double)(cb + 1));
						myocc[(size_t)i * T] = v + 1u;
					});
				rng.init(c.rng_seed, b.tid[s], b.pos[s], slot, (uint32_t)k, 1);
				place_balls(rng, cumr, P, Mr, onesr,
					[&](int i) { return myocc[(size_t)i * T] >> 16; },
					[&](int i) {
						const uint32_t v = myocc[(size_t)i * T];
						const int cb = (int)(v & 0xffffu) + (int)(v >> 16);
						const int N = Ntot[i];
						stat += l10_smem ? L10[N - cb] - L10[cb + 1] : log10((double)(N - cb)) - log10((double)(cb + 1));
						myocc[(size_t)i * T] = v + 0x10000u;
					});
				hit = stat <= thresh ? 1 : 0;
			}
			int total;
			const int H = hits_before + block_prefix(hit, sh->warp_hits, total);
			// first permutation (in order) at which the reference loop would break (contables.c:327-328)
			const bool stop = k <= maxiter && (H >= 10 || (k <= 1000 && H >= 5));
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
			int kk = stop_k, hh = stop_hits;
			if (kk == 0) { kk = maxiter + 1; hh = hits_before; }  // loop ran out: k = maxiter+1 (contables.c:308,336)
			prm.o.ctpval[o] = log10((double)(hh + 1)) - log10((double)(kk + 1));
			prm.o.perm_k[o] = kk;
			prm.o.perm_hits[o] = hh;
		}
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
