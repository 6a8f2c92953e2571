// k_em.cuh — fused likelihood + EM + genotype + pool-statistics kernel: one CTA per (site, allele slot) task that
// survived k_evaluate, i.e. the work of compute_GLL_pooled (crisp/crispEM.c:356-391):
//   calculate_pooled_likelihoods_snp   crispEM.c:91-138   (quality-value likelihood; histogram x table contraction)
//   calculate_pooled_likelihoods_indel crispEM.c:54-87    (count-based likelihood, redone every EM iteration)
//   EMmethod                           crispEM.c:207-351
//   calculate_genotypes                crispEM.c:141-203
//   calculate_pool_statistics          crisp/poolstats.c:5-60 + calculate_uniquereads parsebam/allelecounts.c:42-81
//
// Data layout: the three genotype-likelihood planes that feed outputs (all reads, forward, reverse) live in shared
// memory as G[plane][j][pool] for the whole task, so the EM never touches HBM; the bidirectional plane is only needed
// at j = 0.  EM work is organised as units (round of 32 pools) x (plane): one lane per pool walks j sequentially, so
// the per-pool log-sum-exp needs no cross-lane traffic and terms more than 20 decades below the pool's maximum are
// skipped warp-uniformly (they cannot change a double-precision sum).
#pragma once
#include <type_traits>

#include "dmath.cuh"
#include "engine_types.cuh"
#include "k_evaluate.cuh"

namespace crisp {

constexpr int kEmMaxWarps = 8;
// cut-off of the EM sums: a stored shift is accepted while it lies within 3 decades of the current log-sum-exp, so every
// term above 10^-18 of the sum is kept — dropped terms total less than a fiftieth of an ulp of the sum
constexpr double kExpCut2 = 21.0;
constexpr uint32_t kCutHi = 0xC0350000u;  // high word of -21.0
constexpr int kMaxRounds = 64;  // pools <= 2048

struct EmShared {
	// iteration state (written by warp 0, read by all)
	double lp[3][2];      // log10 p, log10(1-p) per plane
	double p[3];
	double LL[3];
	double E01[3], E10[3];
	double LLnull0;
	int part_w[kEmMaxWarps];
	int exitflag;
	int iter;
	int next_task;
	double delta, deltaf, hwe;
};

// The EM term GENLL + NC + log10(p)*j + log10(1-p)*(n-j) (crispEM.c:249) is evaluated as
//   H[j] + j*(log10 p - log10(1-p)) + n*log10(1-p),   H[j] = GENLL[j] + NC[j]  (stored once per likelihood update),
// so the inner loops cost one shared-memory load and one FMA per genotype; the pool-constant n*log10(1-p) is added
// after the maximum is known.

// tasks[0..n_tasks): the (site, slot) list of this launch; head: its work counter; row0: first AC/QV row of the list
// NOBIDIR: the batch has no bidirectional (overlapping-mate) reads and one pool size: the all-reads plane is not stored
// (it equals forward + reverse - NC), two planes instead of three -> one more CTA per SM (or planes that fit at all)
// LIK: 0 = the likelihood path (dense / sparse, bins / reads) is chosen at run time; 1 = compiled for the dense path on a
// compact batch alone (binned-quality data in the (value, count) format: the benchmark's case) — a third of the code,
// fewer registers
template <bool SMEM_G, bool ASSOC, bool COUNTLIK, bool NOBIDIR, int LIK = 0>
__global__ void __maxnreg__(NOBIDIR ? 80 : 96) k_em(EngineParams prm, const int32_t* __restrict__ tasks, int n_tasks, int32_t* head, int row0)
{
	extern __shared__ __align__(16) unsigned char smem_raw[];
	const DevBatch& b = prm.b;
	const DevCfg& c = prm.c;
	const int P = c.P, nmax = c.nmax, J = nmax + 1;
	const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5, nwarps = blockDim.x >> 5;
	const int rounds = (P + 31) >> 5;
	const size_t plane = (size_t)J * P;

	// ---- shared memory carve-up ----
	EmShared* sh = reinterpret_cast<EmShared*>(smem_raw);
	size_t off = (sizeof(EmShared) + 15) & ~size_t(15);
	double* Gb0 = reinterpret_cast<double*>(smem_raw + off); off += sizeof(double) * P;
	double* gcnt = reinterpret_cast<double*>(smem_raw + off); off += sizeof(double) * J;          // genotype_counts / HWE terms
	int32_t* cnt6 = reinterpret_cast<int32_t*>(smem_raw + off); off += sizeof(int32_t) * 6 * P;   // [6][P]: a1 f,r,b ; a2 f,r,b
	off = (off + 15) & ~size_t(15);
	const int nq = c.nq;
	const bool dense = LIK == 1 || c.dense_lik != 0;
	const bool has_bins = LIK == 1 || b.bins != nullptr;
	// sparse path: one 16-bit histogram [6][128] per warp; dense path: one 32-bit histogram [allele][quality][class][pool]
	uint16_t* hist = reinterpret_cast<uint16_t*>(smem_raw + off);
	uint32_t* dh = reinterpret_cast<uint32_t*>(smem_raw + off);
	off += dense ? sizeof(uint32_t) * 2 * nq * 3 * P : sizeof(uint16_t) * 6 * kNQ * nwarps;
	off = (off + 15) & ~size_t(15);
	double* Ts = reinterpret_cast<double*>(smem_raw + off); off += dense ? sizeof(double) * 2 * nq * J : 0;  // staged table rows
	uint8_t* qd = reinterpret_cast<uint8_t*>(smem_raw + off); off += dense ? kNQ : 0;                         // quality -> dense index
	off = (off + 15) & ~size_t(15);
	// per-round partial sums of one iteration: [0..2] LL per plane, [3..5] pnew per plane, [6] LLnull0, [7..18] T sums
	double* part_base = reinterpret_cast<double*>(smem_raw + off); off += sizeof(double) * 2 * 19 * rounds;  // double-buffered
	int32_t* part_i = reinterpret_cast<int32_t*>(smem_raw + off); off += sizeof(int32_t) * 2 * rounds;
	off = (off + 15) & ~size_t(15);
	double* mprev = reinterpret_cast<double*>(smem_raw + off); off += sizeof(double) * 3 * P;  // log-sum-exp shift per (plane, pool)
	double* NCs = reinterpret_cast<double*>(smem_raw + off); off += NOBIDIR ? sizeof(double) * J : 0;  // log10 C(n,j), one pool size
	off = (off + 15) & ~size_t(15);
	double* G = SMEM_G ? reinterpret_cast<double*>(smem_raw + off) : prm.gws + (size_t)blockIdx.x * prm.gws_stride;
	double* Gt = NOBIDIR ? nullptr : G;
	double* Gf = NOBIDIR ? G : G + plane;
	double* Gr = NOBIDIR ? G + plane : G + 2 * plane;

	if (NOBIDIR) for (int j = tid; j < J; j += blockDim.x) NCs[j] = c.NC[j];
	if (dense) {
		// batch-constant: dense index of each quality character and the table rows of the present qualities
		for (int q = tid; q < kNQ; q += blockDim.x) {
			int d = 0;
			for (int k = 0; k < nq; k++) if (c.qlist[k] < q) d++;
			qd[q] = (uint8_t)d;
		}
		for (int idx = tid; idx < 2 * nq * J; idx += blockDim.x) {
			const int j = idx % J, wd = idx / J, w = wd / nq, d = wd - w * nq;
			Ts[idx] = c.T[((size_t)w * kNQ + c.qlist[d]) * J + j];
		}
	}

	// the work counter is read one task ahead: the round trip of the atomic overlaps the task instead of starting it
	if (tid == 0) sh->next_task = atomicAdd(head, 1);
	for (;;) {
		__syncthreads();
		const int cur_task = sh->next_task;
		__syncthreads();
		if (cur_task >= n_tasks) break;
		if (tid == 0) sh->next_task = atomicAdd(head, 1);
		// measurement switch: phase timestamps of the task (thread 0, straight to the trace buffer)
		auto stamp = [&](int k) {
			if (prm.em_trace && tid == 0) {
				long long t;
				asm volatile("mov.u64 %0, %%globaltimer;" : "=l"(t));
				prm.em_trace[(size_t)(row0 + cur_task) * 8 + k] = t;
			}
		};
		stamp(0);
		const int o = tasks[cur_task], s = o / kSlots, slot = o - s * kSlots;
		const int task = row0 + cur_task;  // AC/QV row
		const uint8_t* mal = b.maxvec_al + s * 8;
		int a1, a2, a3;
		slot_alleles(mal, slot, a1, a2, a3);
		const int refbase = b.refbase[s];
		const int row = slot_amb_row(b, s, slot, a2, a3);
		constexpr bool countlik = COUNTLIK;  // (a2 >= 4 || --usebq 0), decided when the task was queued
		// e(f) is linear in f and f linear in j unless a reference bias bends it: log-likelihood planes concave in j
		const bool unimodal = countlik || c.ref_bias == 0.5;
		int ilen = 0, is_ins = 0;
		if (a2 >= 4 && b.indel_len) { ilen = abs((int)b.indel_len[s * 8 + a2]); is_ins = b.indel_len[s * 8 + a2] > 0; }
		// indel_bias (crispEM.c:36-49)
		double het_bias = 0;
		if (is_ins && ilen >= 1) het_bias = (5.0 + ilen / 2) / (double)(100.0 - 5.0 - ilen / 2);
		else if (ilen >= 1 && !is_ins) het_bias = 5.0 / (double)(100.0 - 5.0);

		// ---- per-pool counts of the two alleles (ambiguity-adjusted) ----
		for (int i = tid; i < P; i += blockDim.x) {
			const int32_t* r = b.indcounts + ((size_t)s * P + i) * kNC;
			int dd[3] = {0, 0, 0};
			if (row >= 0) {
				const int32_t* d = b.amb_dec + ((size_t)row * P + i) * 4;
				dd[0] = d[0]; dd[1] = d[1]; dd[2] = d[2] + d[3];
			}
#pragma unroll
			for (int k = 0; k < 3; k++) {
				cnt6[k * P + i] = r[a1 + 8 * k] - (a1 == refbase ? dd[k] : 0);
				cnt6[(3 + k) * P + i] = r[a2 + 8 * k] - (a2 == refbase ? dd[k] : 0);
			}
		}
		// site totals for the initial indel error rates (crispEM.c:366-374)
		if (tid == 0) {
			double E[3] = {0.001, 0.001, 0.001};
			if constexpr (COUNTLIK)
			for (int k = 0; k < 3; k++) {
				int t1 = b.counts[s * kNC + a1 + 8 * k], t2 = b.counts[s * kNC + a2 + 8 * k];
				if (row >= 0) {
					// counts[] is reduced by the same reads as indcounts[] (process_indel_variant.c:42,46)
					int dec = 0;
					for (int i = 0; i < P; i++) {
						const int32_t* d = b.amb_dec + ((size_t)row * P + i) * 4;
						dec += k == 0 ? d[0] : k == 1 ? d[1] : d[2] + d[3];
					}
					if (a1 == refbase) t1 -= dec;
					if (a2 == refbase) t2 -= dec;
				}
				E[k] = (double)(t2 + 1.0) / (t1 + t2 + 1.0);
			}
			if (COUNTLIK && (E[0] >= 0.005 || E[1] >= 0.005)) E[0] = E[1] = E[2] = 0.005;
			for (int k = 0; k < 3; k++) { sh->E01[k] = E[k]; sh->E10[k] = E[k]; }
			// starting allele frequency: p[0]/K with the clamp of crispEM.c:360
			double p0 = prm.o.paf[o] / c.K;
			if (1.0 - p0 < kMinP) p0 = 1.0 - kMinP;
			sh->p[0] = sh->p[1] = sh->p[2] = p0;
			sh->LL[0] = sh->LL[1] = sh->LL[2] = 0.0;
			sh->exitflag = 0;
			sh->iter = 0;
		}
		// dense path on a compact batch: the offsets of the site's P cells (the bins are then walked as one flat range)
		// and the zeroed histogram are staged under the same barrier.  boff aliases the log-sum-exp shifts, which are
		// first written in EM iteration 0.
		uint32_t* boff = reinterpret_cast<uint32_t*>(mprev);
		if (!COUNTLIK && dense) {
			for (int k = tid; k < 2 * nq * 3 * P; k += blockDim.x) dh[k] = 0;
			if (has_bins) for (int i = tid; i <= P; i += blockDim.x) boff[i] = b.bin_off[(size_t)s * P + i];
		}
		__syncthreads();
		stamp(1);

		// =====================================================================================================
		// likelihood planes
		// =====================================================================================================
		// error rates of the count-based likelihood, identical in every thread (registers)
		double E01[3], E10[3], EL01[3] = {0.001, 0.001, 0.001}, EL10[3] = {0.001, 0.001, 0.001};
		for (int k = 0; k < 3; k++) { E01[k] = sh->E01[k]; E10[k] = sh->E10[k]; }
		auto count_likelihood = [&]() {
			// calculate_pooled_likelihoods_indel (crispEM.c:54-87) with the current E01/E10; the clamps and the
			// product rule for the bidirectional class are applied in place (:65-69)
			double* e01 = E01;
			double* e10 = E10;
			for (int k = 0; k < 3; k++) {
				if (e01[k] < kMinP) e01[k] = kMinP; else if (e01[k] > 1.0 - kMinP) e01[k] = 1.0 - kMinP;
				if (e10[k] < kMinP) e10[k] = kMinP; else if (e10[k] > 1.0 - kMinP) e10[k] = 1.0 - kMinP;
			}
			e01[2] = e01[0] * e01[1];
			e10[2] = e10[0] * e10[1];
			for (int idx = tid; idx < J * P; idx += blockDim.x) {
				const int j = idx / P, i = idx - j * P;
				const int n = c.ploidy[i];
				if (j > n) continue;
				double f = (double)j / n;
				f /= (1 + het_bias);
				double ef = (1.0 - e01[0]) * (1.0 - f) + e10[0] * f, er = (1.0 - e01[1]) * (1.0 - f) + e10[1] * f, eb = (1.0 - e01[2]) * (1.0 - f) + e10[2] * f;
				double gf = cnt6[0 * P + i] * log10(ef) + cnt6[3 * P + i] * log10(1.0 - ef);
				double gr = cnt6[1 * P + i] * log10(er) + cnt6[4 * P + i] * log10(1.0 - er);
				double gb = cnt6[2 * P + i] * log10(eb) + cnt6[5 * P + i] * log10(1.0 - eb);
				const double nc = __ldg(c.NC + (size_t)c.pclass[i] * J + j);
				Gf[idx] = gf + nc;
				Gr[idx] = gr + nc;
				if (!NOBIDIR) Gt[idx] = (gf + gr + gb) + nc;
				if (j == 0) Gb0[i] = gb;
			}
			__syncthreads();
		};

		if constexpr (COUNTLIK) {
			count_likelihood();
			// EMmethod starts from its own E01 = E10 = 0.001 (crispEM.c:211); the count-based values above only
			// shape the first likelihood (crispEM.c:362-375) and, later, the association statistic (:385)
			for (int k = 0; k < 3; k++) { EL01[k] = E01[k]; EL10[k] = E10[k]; E01[k] = 0.001; E10[k] = 0.001; }
		} else if (dense) {
			// calculate_pooled_likelihoods_snp, dense form (few distinct qualities, one pool size).
			// (1) per pool (one warp) bin the packed reads into dh[allele][quality][class][pool];
			// (2) contract with the staged table rows on the FP64 tensor cores (see below).
			if (has_bins) {
				// compact input: the (value, count) bins are the histogram already — one shared-memory add per bin.  The bins of
				// a site are contiguous: every thread takes bins of the flat range (all loads independent, one memory latency
				// for the site) and finds the pool of a bin in the staged offsets.
				const uint32_t b0 = boff[0], b1 = boff[P];
				// four bins per thread and trip, loaded before any is used: a site has ~3.6 bins per thread at the benchmark
				// shape, and one at a time the loop paid a memory latency for each
				for (uint32_t bq = b0 + tid; bq < b1; bq += 4 * blockDim.x) {
					uint32_t ev[4];
#pragma unroll
					for (int u = 0; u < 4; u++) { const uint32_t bb = bq + u * blockDim.x; ev[u] = bb < b1 ? __ldg(b.bins + bb) : 0u; }
#pragma unroll
					for (int u = 0; u < 4; u++) {
						const uint32_t bb = bq + u * blockDim.x, e = ev[u];
						const int base = (e >> 8) & 3;
						const int which = base == a1 ? 0 : (base == a2 ? 1 : -1);
						if (bb >= b1 || which < 0) continue;
						int lo = 0, hi = P;  // boff[lo] <= bb < boff[hi]: the last cell starting at or before bb
						while (hi - lo > 1) { const int mid = (lo + hi) >> 1; if (boff[mid] <= bb) lo = mid; else hi = mid; }
						const int cls = (e >> 11) & 1 ? 2 : ((e >> 10) & 1);
						atomicAdd(dh + ((which * nq + qd[e & 0x7f]) * 3 + cls) * P + lo, e >> 16);
					}
				}
			} else
			for (int i = warp; i < P; i += nwarps) {
				const uint32_t r0 = b.read_off[(size_t)s * P + i], r1 = b.read_off[(size_t)s * P + i + 1];
				for (uint32_t cb = r0; cb < r1; cb += 256) {
					uint32_t wv[8];
#pragma unroll
					for (int k = 0; k < 8; k++) {
						const uint32_t r = cb + k * 32 + lane;
						wv[k] = r < r1 ? (uint32_t)__ldg(b.reads + r) : 0xffffffffu;
					}
#pragma unroll
					for (int k = 0; k < 8; k++) {
						if (cb + k * 32 >= r1) break;
						const uint32_t w = wv[k];
						const int base = (w >> 8) & 3;
						const int which = base == a1 ? 0 : (base == a2 ? 1 : -1);
						const bool use = w != 0xffffffffu && which >= 0;
						const int cls = (w >> 11) & 1 ? 2 : ((w >> 10) & 1);
						const int key = use ? (which * nq + qd[w & 0x7f]) * 3 + cls : -1 - lane;
						const unsigned peers = __match_any_sync(0xffffffffu, key);
						if (use && lane == __ffs(peers) - 1) atomicAdd(dh + key * P + i, (uint32_t)__popc(peers));  // fire-and-forget; one warp per column
					}
				}
			}
			__syncthreads();
			stamp(2);
			// (2) on the FP64 tensor cores: G[class][pool][j] = sum_bin dh[bin][class][pool] * Ts[bin][j] is a small GEMM
			// (M = pools, N = genotypes, K = 2 nq bins).  One DMMA.8x8x4 does the work of eight DFMA warp instructions in one
			// issue slot — the FP64 rate is the same, but this kernel is bound by issue slots, not by the FP64 pipe.
			// Fragments (mma.sync.m8n8k4.row.col.f64): A[row = lane/4][k = lane%4], B[k = lane%4][col = lane/4],
			// D[row = lane/4][col = 2 (lane%4) + {0,1}].
			const int ptiles = (P + 7) >> 3, jtiles = (J + 7) >> 3, ksteps = (2 * nq + 3) >> 2;
			constexpr int NCLS = NOBIDIR ? 2 : 3;  // no bidirectional reads: that class is empty
			// Rows / columns beyond the table are clamped (their results are not stored) and a bin index beyond 2 nq gets a
			// zero B element, so the inner loop carries no bounds predicates.  A warp takes eight pools and TWO genotype tiles
			// per trip: the pool fragment (histogram load + conversion) serves both, and four accumulator chains (2 classes x
			// 2 tiles) are in flight instead of two — the phase ran at the latency of the DMMA chain (6 us of a 38 us task).
			const int kk = lane & 3, g8 = lane >> 2;
			const int jpairs = (jtiles + 1) >> 1;
			for (int item = warp; item < ptiles * jpairs; item += nwarps) {
				const int pt = item / jpairs, jp = item - pt * jpairs;
				const int prow = pt * 8 + g8, jcol0 = jp * 16 + g8, jcol1 = jcol0 + 8;
				const uint32_t* dhp = dh + (prow < P ? prow : P - 1);
				const double* tsp0 = Ts + (jcol0 < J ? jcol0 : J - 1);
				const double* tsp1 = Ts + (jcol1 < J ? jcol1 : J - 1);
				double acc[2][3][2] = {{{0.0, 0.0}, {0.0, 0.0}, {0.0, 0.0}}, {{0.0, 0.0}, {0.0, 0.0}, {0.0, 0.0}}};
				for (int ks = 0; ks < ksteps; ks++) {
					const int wd = ks * 4 + kk;
					const bool kin = wd < 2 * nq;
					const int wdc = kin ? wd : 0;
					const double tv0 = tsp0[wdc * J], tv1 = tsp1[wdc * J];
					const double bf0 = kin ? tv0 : 0.0, bf1 = kin ? tv1 : 0.0;
#pragma unroll
					for (int cls = 0; cls < NCLS; cls++) {
						const double af = u32_to_double(dhp[(wdc * 3 + cls) * P]);
						asm volatile("mma.sync.aligned.m8n8k4.row.col.f64.f64.f64.f64 {%0,%1}, {%2}, {%3}, {%0,%1};"
						             : "+d"(acc[0][cls][0]), "+d"(acc[0][cls][1]) : "d"(af), "d"(bf0));
						asm volatile("mma.sync.aligned.m8n8k4.row.col.f64.f64.f64.f64 {%0,%1}, {%2}, {%3}, {%0,%1};"
						             : "+d"(acc[1][cls][0]), "+d"(acc[1][cls][1]) : "d"(af), "d"(bf1));
					}
				}
				const int i = prow;
				if (i < P) {
#pragma unroll
					for (int h = 0; h < 2; h++) {
#pragma unroll
						for (int t = 0; t < 2; t++) {
							const int j = jp * 16 + h * 8 + kk * 2 + t;
							if (j < J) {
								const double nc = NOBIDIR ? NCs[j] : c.NC[j];
								const size_t idx = (size_t)j * P + i;
								Gf[idx] = acc[h][0][t] + nc;
								Gr[idx] = acc[h][1][t] + nc;
								if (!NOBIDIR) Gt[idx] = (acc[h][0][t] + acc[h][1][t] + acc[h][2][t]) + nc;
								if (j == 0) Gb0[i] = acc[h][2][t];
							}
						}
					}
				}
			}
			__syncthreads();
		} else {
			// calculate_pooled_likelihoods_snp, sparse form (any number of qualities / pool sizes): per pool, histogram
			// the packed reads over (strand class, allele, quality) and contract with the table T[allele][quality][j].
			// One warp per pool; each lane owns genotypes j and j+32 (and j+64, j+96 for pool sizes above 63).
			uint16_t* h = hist + (size_t)warp * 6 * kNQ;
			for (int i = warp; i < P; i += nwarps) {
				const uint32_t r0 = b.read_off[(size_t)s * P + i], r1 = b.read_off[(size_t)s * P + i + 1];
				const int n = c.ploidy[i];
				const double* Tp = c.T + (size_t)c.pclass[i] * 2 * kNQ * J;
				double acc[3][4];
#pragma unroll
				for (int cls = 0; cls < 3; cls++)
#pragma unroll
					for (int jc = 0; jc < 4; jc++) acc[cls][jc] = 0.0;
				if (has_bins) {
					// compact input: every (value, count) bin of the cell is one term count * T[allele][quality][j] — no histogram,
					// no folding; the bin is broadcast and each lane updates its genotypes
					const uint32_t b0 = b.bin_off[(size_t)s * P + i], b1 = b.bin_off[(size_t)s * P + i + 1];
					for (uint32_t bb = b0; bb < b1; bb += 32) {
						const uint32_t mine = bb + lane < b1 ? __ldg(b.bins + bb + lane) : 0u;
						const int m = (int)min(32u, b1 - bb);
						for (int k = 0; k < m; k++) {
							const uint32_t e = __shfl_sync(0xffffffffu, mine, k);
							const int base = (e >> 8) & 3;
							const int which = base == a1 ? 0 : (base == a2 ? 1 : -1);
							if (which < 0) continue;
							const int cls = (e >> 11) & 1 ? 2 : ((e >> 10) & 1);
							const double cntv = u32_to_double(e >> 16);
							const double* Tr = Tp + ((size_t)which * kNQ + (e & 0x7f)) * J + lane;
							double t[4];
#pragma unroll
							for (int jc = 0; jc < 4; jc++) t[jc] = (jc * 32 <= nmax && jc * 32 + lane <= n) ? __ldg(Tr + jc * 32) : 0.0;
							if (cls == 0) {
#pragma unroll
								for (int jc = 0; jc < 4; jc++) acc[0][jc] = fma(cntv, t[jc], acc[0][jc]);
							} else if (cls == 1) {
#pragma unroll
								for (int jc = 0; jc < 4; jc++) acc[1][jc] = fma(cntv, t[jc], acc[1][jc]);
							} else {
#pragma unroll
								for (int jc = 0; jc < 4; jc++) acc[2][jc] = fma(cntv, t[jc], acc[2][jc]);
							}
						}
					}
					// pool sizes above 127: the remaining genotypes, one chunk of 32 at a time over the same bins
					for (int jc = 4; jc * 32 <= n; jc++) {
						const int j = jc * 32 + lane;
						double a3[3] = {0.0, 0.0, 0.0};
						for (uint32_t bb = b0; bb < b1; bb += 32) {
							const uint32_t mine = bb + lane < b1 ? __ldg(b.bins + bb + lane) : 0u;
							const int m = (int)min(32u, b1 - bb);
							for (int k = 0; k < m; k++) {
								const uint32_t e = __shfl_sync(0xffffffffu, mine, k);
								const int base = (e >> 8) & 3;
								const int which = base == a1 ? 0 : (base == a2 ? 1 : -1);
								if (which < 0 || j > n) continue;
								const int cls = (e >> 11) & 1 ? 2 : ((e >> 10) & 1);
								const double t = u32_to_double(e >> 16) * __ldg(Tp + ((size_t)which * kNQ + (e & 0x7f)) * J + j);
								a3[0] += cls == 0 ? t : 0.0;
								a3[1] += cls == 1 ? t : 0.0;
								a3[2] += cls == 2 ? t : 0.0;
							}
						}
						if (j <= n) {
							const size_t idx = (size_t)j * P + i;
							const double nc = __ldg(c.NC + (size_t)c.pclass[i] * J + j);
							Gf[idx] = nc + a3[0];
							Gr[idx] = nc + a3[1];
							if (!NOBIDIR) Gt[idx] = nc + (a3[0] + a3[1] + a3[2]);
						}
					}
				} else
				// 16-bit bins: a pool deeper than 65504 reads is folded block by block
				for (uint32_t blk = r0; blk < r1 || blk == r0; blk += 65504u) {
					const uint32_t bend = min(r1, blk + 65504u);
					{
						uint4* hz = reinterpret_cast<uint4*>(h);
						for (int k = lane; k < 6 * kNQ * 2 / 16; k += 32) hz[k] = make_uint4(0u, 0u, 0u, 0u);
					}
					__syncwarp();
					unsigned present = 0;  // which 32-wide quality chunks hold any read of interest
					for (uint32_t cb = blk; cb < bend; cb += 256) {
						// 8 independent coalesced loads per lane first (memory-level parallelism), then the warp-wide binning
						uint32_t wv[8];
#pragma unroll
						for (int k = 0; k < 8; k++) {
							const uint32_t r = cb + k * 32 + lane;
							wv[k] = r < bend ? (uint32_t)__ldg(b.reads + r) : 0xffffffffu;
						}
#pragma unroll
						for (int k = 0; k < 8; k++) {
							if (cb + k * 32 >= bend) break;
							const uint32_t w = wv[k];
							const bool in = w != 0xffffffffu;
							const int base = (w >> 8) & 3;
							const int which = base == a1 ? 0 : (base == a2 ? 1 : -1);
							const bool use = in && which >= 0;
							const int cls = (w >> 11) & 1 ? 2 : ((w >> 10) & 1);
							const int key = use ? (cls * 2 + which) * kNQ + (w & 0x7f) : -1 - lane;
							const unsigned peers = __match_any_sync(0xffffffffu, key);
							if (use && lane == __ffs(peers) - 1) h[key] += (uint16_t)__popc(peers);
							if (use) present |= 1u << ((w & 0x7f) >> 5);
						}
					}
					present |= __shfl_xor_sync(0xffffffffu, present, 16);
					present |= __shfl_xor_sync(0xffffffffu, present, 8);
					present |= __shfl_xor_sync(0xffffffffu, present, 4);
					present |= __shfl_xor_sync(0xffffffffu, present, 2);
					present |= __shfl_xor_sync(0xffffffffu, present, 1);
					__syncwarp();
#pragma unroll
					for (int cls = 0; cls < 3; cls++) {
#pragma unroll
						for (int which = 0; which < 2; which++) {
							const uint16_t* hh = h + (cls * 2 + which) * kNQ;
							const double* Tw = Tp + (size_t)which * kNQ * J + lane;
							for (int qc = 0; qc < kNQ; qc += 32) {
								if (!((present >> (qc >> 5)) & 1u)) continue;
								const uint32_t hv = hh[qc + lane];
								unsigned m = __ballot_sync(0xffffffffu, hv != 0);
								while (m) {
									const int bq = __ffs(m) - 1;
									m &= m - 1;
									const double cntv = u32_to_double(__shfl_sync(0xffffffffu, hv, bq));
									const double* Tr = Tw + (size_t)(qc + bq) * J;
#pragma unroll
									for (int jc = 0; jc < 4; jc++)
										if (jc * 32 <= nmax && jc * 32 + lane <= n) acc[cls][jc] = fma(cntv, __ldg(Tr + jc * 32), acc[cls][jc]);
								}
							}
						}
					}
					// pool sizes above 127: remaining genotypes are accumulated in place, one chunk at a time
					for (int jc = 4; jc * 32 <= n; jc++) {
						const int j = jc * 32 + lane;
						double a3[3] = {0.0, 0.0, 0.0};
						for (int cw = 0; cw < 6; cw++) {
							const uint16_t* hh = h + cw * kNQ;
							const double* Tw = Tp + (size_t)(cw & 1) * kNQ * J;
							for (int q = 0; q < kNQ; q++) {
								const uint32_t hv = hh[q];
								if (hv && j <= n) a3[cw >> 1] = fma((double)hv, __ldg(Tw + (size_t)q * J + j), a3[cw >> 1]);
							}
						}
						if (j <= n) {
							const size_t idx = (size_t)j * P + i;
							const bool first = blk == r0;
							const double nc = first ? __ldg(c.NC + (size_t)c.pclass[i] * J + j) : 0.0;
							Gf[idx] = (first ? nc : Gf[idx]) + a3[0];
							Gr[idx] = (first ? nc : Gr[idx]) + a3[1];
							if (!NOBIDIR) Gt[idx] = (first ? nc : Gt[idx]) + (a3[0] + a3[1] + a3[2]);
						}
					}
					__syncwarp();
					if (bend >= r1) break;
				}
				const double* ncp = c.NC + (size_t)c.pclass[i] * J;
#pragma unroll
				for (int jc = 0; jc < 4; jc++) {
					const int j = jc * 32 + lane;
					if (jc * 32 <= nmax && j <= n) {
						const size_t idx = (size_t)j * P + i;
						const double nc = __ldg(ncp + j);
						Gf[idx] = acc[0][jc] + nc;
						Gr[idx] = acc[1][jc] + nc;
						if (!NOBIDIR) Gt[idx] = (acc[0][jc] + acc[1][jc] + acc[2][jc]) + nc;
						if (j == 0) Gb0[i] = acc[2][jc];
					}
				}
				__syncwarp();
			}
			__syncthreads();
		}

		stamp(3);
		// =====================================================================================================
		// EM iterations (crispEM.c:233-314)
		// =====================================================================================================
		// Iteration state lives in registers, replicated in every warp (lane pl holds plane pl): ONE barrier per iteration,
		// after which every warp reduces the per-round partial sums in the same fixed order and derives the next frequencies
		// itself (~120 instructions).  A version that left this to one warp and published the result through shared memory
		// saved those instructions in the other five, but the five then waited at a second barrier for a dependent chain
		// that was 45 % of the iteration's duration (task timeline, scripts/em_trace.py).  The partial sums are double
		// buffered, so a warp that runs ahead into the next iteration does not overwrite what a slower one still reads.
		double my_p = sh->p[0], my_LL = 0.0;
		// log10 p (even lanes) and log10(1-p) (odd lanes) of plane lane/2; the planes start from the same frequency
		double cur_lg = log10_fast((lane & 1) ? 1.0 - my_p : my_p);
		int it = 0;
		for (;;) {
			double* part = part_base + (it & 1) * 19 * rounds;
			const int nunits = rounds * 3;
			for (int u = warp; u < nunits; u += nwarps) {
				const int pl = u / rounds, rd = u - pl * rounds;
				const int i = rd * 32 + lane;
				const bool act = i < P;
				const int ii = act ? i : P - 1;
				const int n = NOBIDIR ? nmax : c.ploidy[ii];   // NOBIDIR: one pool size
				// plane pl as a column of H; with NOBIDIR plane 0 is formed as forward + reverse - NC on the fly
				const bool sum2 = NOBIDIR && pl == 0;
				const double* Hp = (NOBIDIR ? (pl == 2 ? Gr : Gf) : G + (size_t)pl * plane) + ii;
				const double* Hq = Gr + ii;
				const double l0 = __shfl_sync(0xffffffffu, cur_lg, 2 * pl), l1 = __shfl_sync(0xffffffffu, cur_lg, 2 * pl + 1);
				const double dl = l0 - l1, base = (double)n * l1;
				// value of genotype j: H[j] + j*dl  (tag: compile-time "plane 0 of the two-plane layout")
				auto term = [&](auto tag, const double* hp, const double* hq, int j, double jd) -> double {
					if constexpr (decltype(tag)::value) return fma(jd, dl, (*hp + *hq) - NCs[j]);
					else return fma(jd, dl, *hp);
				};
				// Shift of the log-sum-exp: the exact maximum on the first iteration, afterwards the maximum seen in the
				// previous iteration (any shift within a few decades of the maximum gives the same sum; the fallback below
				// repeats the unit when the assumed shift turns out more than 3 decades too high).
				double m;
				auto exact_max = [&](auto tag) {
					m = -1e300;
					const double* hp = Hp;
					const double* hq = Hq;
					double jd = 0.0;
					// two genotypes per trip (independent loads and FMAs; one vote per pair)
					for (int j = 0; j <= nmax; j += 2, hp += 2 * P, hq += 2 * P, jd += 2.0) {
						const bool in0 = NOBIDIR || j <= n, in1 = j < nmax && (NOBIDIR || j + 1 <= n);
						const double v0 = in0 ? term(tag, hp, hq, j, jd) : -1e300;
						const double v1 = in1 ? term(tag, hp + P, hq + P, j + 1, jd + 1.0) : -1e300;
						m = v0 > m ? v0 : m;
						m = v1 > m ? v1 : m;
						// H[j] + j*dl is concave in j (unimodal): past the mode nothing larger follows
						if (unimodal && __all_sync(0xffffffffu, !in1 || v1 < v0)) break;
					}
				};
				if (it == 0 || countlik) {  // count-based planes are rebuilt every iteration: always take the exact maximum
					if (sum2) exact_max(std::true_type{}); else exact_max(std::false_type{});
				} else m = mprev[pl * P + ii];
				double S = 0.0, Wj = 0.0;
				double T00[3] = {0, 0, 0}, T01[3] = {0, 0, 0}, T10[3] = {0, 0, 0}, T11[3] = {0, 0, 0};
				auto sum_loop = [&](auto tag) {
					const Exp10K ek = exp10_constants();
					const double* hp = Hp;
					const double* hq = Hq;
					double jd = 0.0;
					auto add_term = [&](double e, double jv) {
						S += e;
						Wj = fma(jv, e, Wj);
						if (countlik && pl == 0) {
							const double f = jv / n;
#pragma unroll
							for (int k = 0; k < 3; k++) {
								const double den1 = (1.0 - f) * (1.0 - E01[k]) + f * E10[k], den2 = (1.0 - f) * E01[k] + f * (1.0 - E10[k]);
								T00[k] += e * (1.0 - f) * (1.0 - E01[k]) / den1;
								T10[k] += e * f * E10[k] / den1;
								T11[k] += e * f * (1.0 - E10[k]) / den2;
								T01[k] += e * (1.0 - f) * E01[k] / den2;
							}
						}
					};
					// Two genotypes per trip: a warp walks its unit alone, one dependent chain of ~12 FP64 operations per 10^d, so
					// the loop ran at the latency of that chain; the two chains of a pair are independent and interleave.  Terms are
					// added in the order of j, a term below the cut-off adds an exact zero: the sums are those of the one-by-one walk.
					for (int j = 0; j <= nmax; j += 2, hp += 2 * P, hq += 2 * P, jd += 2.0) {
						const bool in0 = NOBIDIR || j <= n, in1 = j < nmax && (NOBIDIR || j + 1 <= n);
						const double v0 = in0 ? term(tag, hp, hq, j, jd) : -1e300;
						const double v1 = in1 ? term(tag, hp + P, hq + P, j + 1, jd + 1.0) : -1e300;
						const double d0 = v0 - m, d1 = v1 - m;
						// d > -cut on the integer pipe: the high word of a double, as unsigned, orders like -value for negatives and
						// sits below 0x80000000 for positives (the low word decides nothing that matters for a cut-off)
						const bool live0 = (uint32_t)__double2hiint(d0) < kCutHi, live1 = (uint32_t)__double2hiint(d1) < kCutHi;
						if (live0 || live1) {
							const double e0 = exp10_lse(live0 ? d0 : 0.0, ek), e1 = exp10_lse(live1 ? d1 : 0.0, ek);
							add_term(live0 ? e0 : 0.0, jd);
							add_term(live1 ? e1 : 0.0, jd + 1.0);
						}
						// descending and below the cut-off for every lane at the second of the pair (then also at whatever follows)
						if (unimodal && __all_sync(0xffffffffu, !in1 || (v1 < v0 && !live1))) break;
					}
				};
				// The shift kept for the next iteration is this iteration's log-sum-exp itself (within log10(n+1) decades above
				// the maximum).  A stored shift is good while the sum it produces stays within three decades of 1; otherwise (the
				// frequency moved a lot, or the sum underflowed to zero) the unit is redone with the corrected shift.
				double lS = 0.0;
				for (int attempt = 0; attempt < 3; attempt++) {
					if (sum2) sum_loop(std::true_type{}); else sum_loop(std::false_type{});
					lS = log10_fast(S);
					if (__all_sync(0xffffffffu, fabs(lS) <= 3.0)) break;   // false for S = 0 (-inf) and NaN as well
					if (__any_sync(0xffffffffu, !(S > 0.0) || !(fabs(lS) < 1e300))) {
						if (sum2) exact_max(std::true_type{}); else exact_max(std::false_type{});
					} else m += lS;
					S = 0.0; Wj = 0.0;
#pragma unroll
					for (int k = 0; k < 3; k++) { T00[k] = T01[k] = T10[k] = T11[k] = 0.0; }
				}
				if (act) mprev[pl * P + ii] = m + lS;
				m += base;
				double L = m + lS;
				double llc = 0.0, pn = 0.0, nul = 0.0;
				if (act) {
					llc = L;
					pn = Wj / S;
					if (pl == 0) nul = NOBIDIR ? Gf[ii] + Gr[ii] : Gt[ii];  // GENLL[i][0] (NC[0] = 0)
					else if (pl == 1) llc += Gr[ii] + Gb0[ii];   // LLtotalf += GENLLr[i][0] + GENLLb[i][0]  (:262)
					else llc += Gf[ii] + Gb0[ii];                // LLtotalr += GENLLf[i][0] + GENLLb[i][0]  (:261)
				}
				llc = warp_sum(llc);
				pn = warp_sum(pn);
				if (pl == 0) nul = warp_sum(nul);
				if (countlik && pl == 0) {
#pragma unroll
					for (int k = 0; k < 3; k++) {
						const double c1 = act ? (double)cnt6[k * P + ii] / S : 0.0, c2 = act ? (double)cnt6[(3 + k) * P + ii] / S : 0.0;
						double t00 = warp_sum(c1 * T00[k]), t10 = warp_sum(c1 * T10[k]), t11 = warp_sum(c2 * T11[k]), t01 = warp_sum(c2 * T01[k]);
						if (lane == 0) { part[(7 + k) * rounds + rd] = t00; part[(10 + k) * rounds + rd] = t01; part[(13 + k) * rounds + rd] = t10; part[(16 + k) * rounds + rd] = t11; }
					}
				}
				if (lane == 0) {
					part[pl * rounds + rd] = llc;
					part[(3 + pl) * rounds + rd] = pn;
					if (pl == 0) part[6 * rounds + rd] = nul;
				}
			}
			__syncthreads();
			// ---- finish the iteration (every warp, lane pl handles plane pl): partial sums, clamps, convergence test ----
			double LLn = 0.0, pn = 0.0;
			if (lane < 3) {
				// (up to 64 pools spelled out: the general loop, unrolled by the compiler for any `rounds`, was a sixth of the
				// instructions of this stretch)
				if (rounds == 2) { LLn = part[lane * 2] + part[lane * 2 + 1]; pn = part[(3 + lane) * 2] + part[(3 + lane) * 2 + 1]; }
				else if (rounds == 1) { LLn = part[lane]; pn = part[3 + lane]; }
				else
				for (int r = 0; r < rounds; r++) { LLn += part[lane * rounds + r]; pn += part[(3 + lane) * rounds + r]; }
				pn /= c.K;
				if (pn < kMinP) pn = kMinP;
				if (1.0 - pn <= kMinP) pn = 1.0 - kMinP;
			}
			// convergence: fabsf() of the double differences against double thresholds (crispEM.c:304)
			const double dconv = (double)fabsf((float)(LLn - my_LL));
			const double d0 = __shfl_sync(0xffffffffu, dconv, 0), d1 = __shfl_sync(0xffffffffu, dconv, 1), d2 = __shfl_sync(0xffffffffu, dconv, 2);
			my_LL = LLn;
			my_p = lane < 3 ? pn : 0.5;
			if (countlik) {
				double e01 = 0.0, e10 = 0.0;
				if (lane < 3) {
					const int k = lane;
					double t00 = 0, t01 = 0, t10 = 0, t11 = 0;
					for (int r = 0; r < rounds; r++) { t00 += part[(7 + k) * rounds + r]; t01 += part[(10 + k) * rounds + r]; t10 += part[(13 + k) * rounds + r]; t11 += part[(16 + k) * rounds + r]; }
					e01 = (t01 + 1.5 - 1) / (t01 + t00 + 1.5 + 50 - 1);
					e10 = (t10 + 1.5 - 1) / (t10 + t11 + 50 + 1.5 - 1);
					if (e10 > 0.05) e10 = 0.05;
				}
				for (int k = 0; k < 3; k++) { E01[k] = __shfl_sync(0xffffffffu, e01, k); E10[k] = __shfl_sync(0xffffffffu, e10, k); }
			}
			const bool stop = (it + 1 >= 2 && d0 <= 0.0001 && d1 <= 0.01 && d2 <= 0.01) || it + 1 >= 1000;
			it++;
			if (it == 1) stamp(4);
			if (stop) {
				if (warp == 0) {
					// statistics of the fit (crispEM.c:321-349)
					double LLnull0 = 0.0;
					for (int r = 0; r < rounds; r++) LLnull0 += part[6 * rounds + r];
					const double LLt = __shfl_sync(0xffffffffu, my_LL, 0), LLf = __shfl_sync(0xffffffffu, my_LL, 1), LLr = __shfl_sync(0xffffffffu, my_LL, 2);
					const double pf = __shfl_sync(0xffffffffu, my_p, 0);
					const double df = (LLf > LLr) ? LLf - LLt - log10(2.0) : LLr - LLt - log10(2.0);
					const double ALPHA = c.alpha_prior;
					const double LLrefadd = LLnull0 + log10_fast(1.0 - ALPHA);
					double LL = LLt + log10_fast(ALPHA);
					if (LLrefadd < LL) LL += log10_fast(1.0 + exp10(LLrefadd - LL));
					else LL = LLrefadd + log10_fast(1.0 + exp10(LL - LLrefadd));
					if (lane == 0) { sh->delta = LL - LLrefadd; sh->deltaf = df; sh->p[0] = pf; sh->hwe = 0.0; }
				}
				if (countlik) count_likelihood();  // crispEM.c:298, also on the exit iteration
				break;
			}
			{
				const double pp = __shfl_sync(0xffffffffu, my_p, (lane >> 1) < 3 ? (lane >> 1) : 0);
				cur_lg = log10_fast((lane & 1) ? 1.0 - pp : pp);   // lane 2 pl: log10 p, lane 2 pl + 1: log10(1-p)
			}
			if (countlik) count_likelihood();  // crispEM.c:298 (ends on a barrier: the planes are rebuilt before the next units read them)
		}
		__syncthreads();  // the statistics of the fit, written by warp 0
		stamp(5);
		// (the planes were last read before the barriers that closed the final iteration: the genotype phase may overwrite them)
		const double delta = sh->delta, deltaf = sh->deltaf, pfin = sh->p[0];
		const bool called = delta >= 2;
		int varpools = 0;
		if (called) {
			// ---- calculate_genotypes: argmax / runner-up per pool, posterior genotype counts, HWE ----
			const double l0 = log10_fast(pfin), l1 = log10_fast(1.0 - pfin);
			double* post = Gf;  // forward plane is no longer needed: reuse as posterior[j][pool]
			// with more warps than rounds of pools the genotype pass occupies `rounds` warps; the others run the pool statistics
			// (which read the counts and the alt-read list only) at the same time instead of waiting at the barrier
			const int gw = rounds < nwarps ? rounds : 0;
			if (gw == 0 || warp < gw)
			for (int rd = warp; rd < rounds; rd += nwarps) {
				const int i = rd * 32 + lane;
				const bool act = i < P;
				const int ii = act ? i : P - 1;
				const int n = c.ploidy[ii];
				const double* Hp = (NOBIDIR ? Gf : Gt) + ii;
				const double* Hq = Gr + ii;
				const double dl = l0 - l1, base = (double)n * l1;
				double best = -100000, second = -100000;
				int bg = 0;
				// The posterior terms t[j] are concave in j (see `unimodal`): once every lane of the warp is past its mode
				// nothing larger follows, and once it is also below the cut-off nothing that counts follows — the three
				// passes stop there instead of walking all n+1 genotypes (the argmax, the runner-up and every term of the
				// sums that is kept are the ones a full walk finds).
				{
					const double* hp = Hp;
					const double* hq = Hq;
					double jd = 0.0, tprev = -1e300;
					for (int j = 0; j <= nmax; j++, hp += P, hq += P, jd += 1.0) {
						const bool in = j <= n;
						const double t = in ? fma(jd, dl, NOBIDIR ? (*hp + *hq) - NCs[j] : *hp) + base : -1e300;
						if (in) {
							if (t > best) { second = best; bg = j; best = t; }
							else if (t > second) second = t;
						}
						if (unimodal && __all_sync(0xffffffffu, !in || t < tprev)) break;
						tprev = t;
					}
				}
				// the running log-sum-exp of :168-169 equals max + log10(sum 10^(t-max)) up to rounding
				double S = 0.0;
				{
					const Exp10K ek = exp10_constants();
					const double* hp = Hp;
					const double* hq = Hq;
					double jd = 0.0, tprev = -1e300;
					for (int j = 0; j <= nmax; j++, hp += P, hq += P, jd += 1.0) {
						const bool in = j <= n;
						const double t = in ? fma(jd, dl, NOBIDIR ? (*hp + *hq) - NCs[j] : *hp) + base : -1e300;
						const double d = t - best;
						if (in && d > -kExpCut) S += exp10_lse(d, ek);
						if (unimodal && __all_sync(0xffffffffu, !in || (t < tprev && d <= -kExpCut))) break;
						tprev = t;
					}
				}
				const double Lb = best + log10_fast(S);
				{
					const Exp10K ek = exp10_constants();
					const double* hp = Hp;
					const double* hq = Hq;
					double jd = 0.0, tprev = -1e300;
					bool tail = false;  // every lane is past its mode and below the cut-off: the rest of the column is zero
					for (int j = 0; j <= nmax; j++, hp += P, hq += P, jd += 1.0) {
						double v = 0.0;
						if (!tail) {
							const bool in = j <= n;
							// (NOBIDIR: the forward value is read here before this lane overwrites it with the posterior)
							const double t = in ? fma(jd, dl, NOBIDIR ? (*hp + *hq) - NCs[j] : *hp) + base : -1e300;
							const double d = t - Lb;
							if (act && in && d > -kExpCut - 3) v = exp10_lse(d, ek);
							if (unimodal && __all_sync(0xffffffffu, !in || (t < tprev && d <= -kExpCut - 3))) tail = true;
							tprev = t;
						}
						if (act) post[(size_t)j * P + ii] = v;
					}
				}
				if (act) {
					prm.o.ac[(size_t)task * P + i] = bg;
					prm.o.qv[(size_t)task * P + i] = 10 * (best - second);
				}
				const int vp = warp_sum(act && bg > 0 ? 1 : 0), acs = warp_sum(act ? bg : 0);
				if (lane == 0) { part_i[rd] = vp; part_i[rounds + rd] = acs; }
			}
			if (gw == 0) __syncthreads();
			// ---- calculate_pool_statistics: read-evidence filter per pool ----
			const uint32_t alo = b.alt_off ? b.alt_off[s * kSlots + slot] : 0, ahi = b.alt_off ? b.alt_off[s * kSlots + slot + 1] : 0;
			const int MR = c.min_reads;
			int vpl = 0;
			if (gw == 0 || warp >= gw) {
			const int w0 = gw, nw = nwarps - gw, gt = tid - gw * 32, gn = nw * 32;  // the warps of this phase, as a group
			auto group_sync = [&]() {
				if (gw == 0) __syncthreads();
				else asm volatile("bar.sync 1, %0;" ::"r"(gn) : "memory");
			};
			// first entry of every pool in the (pool-sorted) alt-read list of this slot, built by one pass over the list;
			// the array aliases the log-sum-exp shifts, which the EM no longer needs
			int32_t* astart = reinterpret_cast<int32_t*>(mprev);
			if (ahi > alo) {
				for (uint32_t e = alo + gt; e < ahi; e += gn) {
					const int pe = b.alt_reads[e].pool, pp = e > alo ? (int)b.alt_reads[e - 1].pool : -1;
					for (int q = pp + 1; q <= pe; q++) astart[q] = (int32_t)e;
					if (e == ahi - 1) for (int q = pe + 1; q <= P; q++) astart[q] = (int32_t)ahi;
				}
				group_sync();
			}
			for (int i = warp - w0; i < P; i += nw) {
				const int f = cnt6[3 * P + i], r = cnt6[4 * P + i], bd = cnt6[5 * P + i];
				if (f + r + bd < 2) continue;
				const uint32_t e0 = ahi > alo ? (uint32_t)astart[i] : alo;
				int m = ahi > alo ? astart[i + 1] - astart[i] : 0;
				if (m > f + r + bd) m = f + r + bd;  // `i < reads` cap, allelecounts.c:53
				// Reads with a (strand, start) pair not seen earlier in the pool's list (calculate_uniquereads).  The list was
				// re-read from memory once per earlier read (m^2/2 loads per pool: the alleles at 10-20 % frequency made the
				// slowest tasks of a launch, 110-190 us against 50); now 32 reads per trip: one __match_any finds the repeats
				// inside the trip, the keys of the first four trips stay in registers and are compared through shuffles, and
				// only a list beyond 128 reads goes back to memory for the part past them.
				int uq = 0, mid = 0;
				uint32_t kept[4] = {0, 0, 0, 0};
				for (int c0 = 0, ch = 0; c0 < m; c0 += 32, ch++) {
					const int e = c0 + lane;
					const bool in = e < m;
					crispgpu_altread x{};
					if (in) x = b.alt_reads[e0 + e];
					const uint32_t key = in ? ((uint32_t)x.strand << 16) | x.offset : 0x80000000u | (uint32_t)lane;
					const unsigned peers = __match_any_sync(0xffffffffu, key);
					bool first = in && __ffs(peers) - 1 == lane;
#pragma unroll
					for (int pc = 0; pc < 4; pc++) {
						if (pc < ch) {
							bool seen = false;
							for (int j = 0; j < 32; j++) seen |= __shfl_sync(0xffffffffu, kept[pc], j) == key;
							first = first && !seen;
						}
					}
					if (ch >= 4) {
						for (int e2 = 128; e2 < c0 && first; e2++) {
							const crispgpu_altread y = b.alt_reads[e0 + e2];
							if (y.strand == x.strand && y.offset == x.offset) first = false;
						}
					}
#pragma unroll
					for (int pc = 0; pc < 4; pc++) if (pc == ch) kept[pc] = key;
					uq += first ? 1 : 0;
					if (in && x.offset >= 5 && (int)x.aligned - (int)x.offset >= 5) mid++;
				}
				uq = warp_sum(uq);
				mid = warp_sum(mid);
				if (m == 0) uq = 1;  // *unique starts at 1 (allelecounts.c:45)
				int varpool = 0;
				if (f + r + bd >= MR && mid > 0 && uq >= MR - 1) varpool = 1;
				else if (f + r + bd >= MR - 1 && f + bd > 0 && r + bd > 0 && mid > 0 && uq >= MR - 1) varpool = 1;
				vpl += varpool;
			}
			}
			__syncthreads();
			if (c.varpoolsize == 0) {
				// HWE chi-square over genotype classes with expected count >= 2 (crispEM.c:187-199)
				const int n0 = c.ploidy[0];
				// expected class counts, one thread per class; then one warp per class sums the posterior over pools
				for (int j = tid; j <= n0; j += blockDim.x) gcnt[j] = exp10(c.NC[j] + j * l0 + (n0 - j) * l1 + log10((double)P));
				__syncthreads();
				for (int j = warp; j <= n0; j += nwarps) {
					const double ec = gcnt[j];
					if (ec >= 2) {
						double gc = 0.0;
						for (int i = lane; i < P; i += 32) gc += post[(size_t)j * P + i];
						gc = warp_sum(gc);
						__syncwarp();
						if (lane == 0) gcnt[j] = (gc - ec) * (gc - ec) / ec;
					} else if (lane == 0) gcnt[j] = -1.0;
				}
				__syncthreads();
				if (tid == 0) {
					double stat = 0.0, dof = 0.0;
					for (int j = 0; j <= n0; j++) if (gcnt[j] >= 0.0) { stat += gcnt[j]; dof += 1; }
					sh->hwe = (stat > 0 && dof >= 2) ? d_gammaq((dof - 1) / 2, stat / 2) / log(10.0) : 0.0;
				}
			}
			// ---- calculate_LRstatistic (crisp/LRstatistic.c:48-69), --phenotypes only: five EMs restricted to pool subsets ----
			if (ASSOC && !NOBIDIR && c.association && pfin >= 0.0001 && delta >= 3) {
				__syncthreads();  // HWE above still reads the posterior plane
				for (int k = 0; k < 3; k++) { E01[k] = EL01[k]; E10[k] = EL10[k]; }
				const int flags[5] = {7, 1, 6, 4, 5};
				double pout[5], llout[5];
				for (int q = 0; q < 5; q++) {
					const int flag = flags[q];
					int Ksub = 0;
					for (int i = 0; i < P; i++) if (c.phenotypes[i] & flag) Ksub += c.ploidy[i];
					double p = pfin, LLprev = 0.0, LL = 0.0;
					for (int it2 = 1; it2 <= 1000; it2++) {
						double* pb = part_base + (it2 & 1) * 19 * rounds;
						const double l0 = log10(p), l1 = log10(1.0 - p), dl = l0 - l1;
						for (int rd = warp; rd < rounds; rd += nwarps) {
							const int i = rd * 32 + lane;
							const bool act = i < P && (c.phenotypes[i < P ? i : 0] & flag) != 0;
							const int ii = i < P ? i : P - 1;
							const int n = c.ploidy[ii];
							const double* hp0 = Gt + ii;
							double m = -1e300, S = 0.0, Wj = 0.0;
							{
								const double* hp = hp0; double jd = 0.0;
								for (int j = 0; j <= n; j++, hp += P, jd += 1.0) m = fmax(m, fma(jd, dl, *hp));
							}
							{
								const double* hp = hp0; double jd = 0.0;
								for (int j = 0; j <= n; j++, hp += P, jd += 1.0) {
									const double d = fma(jd, dl, *hp) - m;
									if (d > -kExpCut) { const double e = exp10_lse(d); S += e; Wj = fma(jd, e, Wj); }
								}
							}
							const double L = m + (double)n * l1 + log10(S);
							const double lls = warp_sum(act ? L : 0.0), pns = warp_sum(act ? Wj / S : 0.0);
							if (lane == 0) { pb[rd] = lls; pb[rounds + rd] = pns; }
						}
						__syncthreads();
						LL = 0.0;
						double pn = 0.0;
						for (int r = 0; r < rounds; r++) { LL += pb[r]; pn += pb[rounds + r]; }
						pn /= Ksub;
						if (countlik) count_likelihood();  // LRstatistic.c:40 (with compute_GLL_pooled's error rates)
						if (pn < kMinP) pn = kMinP;
						if (1.0 - pn <= kMinP) pn = 1.0 - kMinP;
						const bool done = it2 >= 2 && (double)fabsf((float)(LL - LLprev)) <= 0.01;
						p = pn;
						LLprev = LL;
						if (done) break;
					}
					pout[q] = p;
					llout[q] = LL;
					__syncthreads();  // the next subset reuses the partial-sum buffers
				}
				if (tid == 0) {
					const double pctl = pout[1], pcase = pout[2], pcase1 = pout[3];
					const double LLR = llout[2] + llout[1] - llout[0], LLR1 = llout[3] + llout[1] - llout[4];
					prm.o.assoc_llr[o] = LLR;
					prm.o.assoc_or[o] = pcase * (1.0 - pctl) / ((1.0 - pcase) * pctl);
					prm.o.assoc_p[o] = d_gammaq(0.5, LLR * log(10.0)) / log(10.0);
					prm.o.assoc_llr1[o] = LLR1;
					prm.o.assoc_or1[o] = pcase1 * (1.0 - pctl) / ((1.0 - pcase1) * pctl);
					prm.o.assoc_p1[o] = d_gammaq(0.5, LLR1 * log(10.0)) / log(10.0);
				}
				__syncthreads();
			}
			// combine per-warp pool counts
			__syncthreads();
			if (lane == 0) sh->part_w[warp] = vpl;
			__syncthreads();
			if (tid == 0) for (int w = 0; w < nwarps; w++) varpools += sh->part_w[w];
		}
		__syncthreads();
		if (tid == 0) {
			prm.o.af_ml[o] = pfin;
			prm.o.delta[o] = delta;
			prm.o.deltaf[o] = deltaf;
			prm.o.hwe[o] = sh->hwe;
			prm.o.em_iters[o] = it;
			if (prm.em_trace) {
				unsigned sm;
				asm volatile("mov.u32 %0, %%smid;" : "=r"(sm));
				stamp(6);
				prm.em_trace[(size_t)task * 8 + 7] = (long long)sm | ((long long)o << 16) | ((long long)(it * 2 + (called ? 1 : 0)) << 48);
			}
			prm.o.call_row[o] = called ? task : -1;
			if (called) {
				int vp = 0, acs = 0;
				for (int r = 0; r < rounds; r++) { vp += part_i[r]; acs += part_i[rounds + r]; }
				prm.o.variantpools[o] = vp;
				prm.o.allelecount[o] = acs;
				prm.o.varpools[o] = varpools;
				prm.o.status[o] = CRISPGPU_SLOT_CALLED;
			}
		}
	}
}

// likelihood table T[pclass][which][q][j] (crispEM.c:110,117-118): which=0 -> log10 e, which=1 -> log10(1-e)
__global__ void k_build_T(double* T, const int32_t* class_ploidy, int n_pclass, int J, int qv_offset, double beta)
{
	const size_t total = (size_t)n_pclass * kNQ * J;
	for (size_t idx = blockIdx.x * (size_t)blockDim.x + threadIdx.x; idx < total; idx += (size_t)gridDim.x * blockDim.x) {
		const int j = (int)(idx % J);
		const int q = (int)((idx / J) % kNQ);
		const int pc = (int)(idx / ((size_t)J * kNQ));
		const int n = class_ploidy[pc];
		double v0 = 0.0, v1 = 0.0;
		if (j <= n) {
			const double ep = pow(0.1, ((double)q - qv_offset) / 10);
			double f = (double)j * (1.0 - beta);
			f /= f + (double)(n - j) * beta;
			const double e = (1.0 - ep) * (1.0 - f) + ep * f;
			v0 = log10(e);
			v1 = log10(1.0 - e);
		}
		T[((size_t)(pc * 2 + 0) * kNQ + q) * J + j] = v0;
		T[((size_t)(pc * 2 + 1) * kNQ + q) * J + j] = v1;
	}
}

}  // namespace crisp
