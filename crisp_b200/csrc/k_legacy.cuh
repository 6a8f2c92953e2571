// k_legacy.cuh — the --EM 0 caller: pooled_variantcaller + call_genotypes (crisp/oldcrispmethod.c:131-204, 4-110).
// One warp per surviving (site, slot); lanes run over pools.  Uses the Chernoff quality-value p-values
// (FET/contables.c:342-364), the 2x2 Fisher strand test (:131-171) and the lower binomial tails (:197-209).
#pragma once
#include "dmath.cuh"
#include "engine_types.cuh"
#include "k_evaluate.cuh"

namespace crisp {

// distinct (strand, offset) start sites and "middle" reads among the first m alt-list entries of one pool
// (calculate_uniquereads, parsebam/allelecounts.c:42-81); serial version for one lane
__device__ inline void unique_middle_serial(const crispgpu_altread* v, int m, int& uq, int& mid)
{
	uq = 1;
	mid = 0;
	if (m <= 0) return;
	uq = 0;
	for (int e = 0; e < m; e++) {
		const crispgpu_altread x = v[e];
		bool first = true;
		for (int e2 = 0; e2 < e; e2++) {
			const crispgpu_altread y = v[e2];
			if (y.strand == x.strand && y.offset == x.offset) { first = false; break; }
		}
		uq += first ? 1 : 0;
		if (x.offset >= 5 && (int)x.aligned - (int)x.offset >= 5) mid++;
	}
}

template <int WARPS>
__global__ void __launch_bounds__(WARPS * 32) k_legacy(EngineParams prm)
{
	extern __shared__ __align__(16) unsigned char smem_raw[];
	const DevBatch& b = prm.b;
	const DevCfg& c = prm.c;
	const int P = c.P, lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
	double* pvstrand = reinterpret_cast<double*>(smem_raw) + (size_t)warp * P;
	const int ntask = prm.q.counters[1];
	for (int task = blockIdx.x * WARPS + warp; task < ntask; task += gridDim.x * WARPS) {
		const int o = prm.q.call_tasks[task], s = o / kSlots, slot = o - s * kSlots;
		int a1, a2, a3;
		slot_alleles(b.maxvec_al + s * 8, slot, a1, a2, a3);
		const int refbase = b.refbase[s];
		const int row = slot_amb_row(b, s, slot, a2, a3);
		const double ctpval = prm.o.ctpval[o];
		const double ser = (a1 >= 4 || a2 >= 4) ? 0.05 : c.ser;
		// ---- site totals (ambiguity-adjusted) ----
		int cnts[16];
		double tst[16];
#pragma unroll
		for (int q = 0; q < 16; q++) { cnts[q] = b.counts[s * kNC + q]; tst[q] = b.tstats ? b.tstats[(size_t)s * 16 + q] : 0.0; }
		if (row >= 0) {
			int d0 = 0, d1 = 0;
			double e0 = 0, e1 = 0;
			for (int i = lane; i < P; i += 32) {
				const int32_t* d = b.amb_dec + ((size_t)row * P + i) * 4;
				d0 += d[0]; d1 += d[1];
				if (b.amb_ep) { e0 += b.amb_ep[((size_t)row * P + i) * 2]; e1 += b.amb_ep[((size_t)row * P + i) * 2 + 1]; }
			}
			d0 = warp_sum(d0); d1 = warp_sum(d1); e0 = warp_sum(e0); e1 = warp_sum(e1);
			cnts[refbase] -= d0; cnts[refbase + 8] -= d1; tst[refbase] -= e0; tst[refbase + 8] -= e1;
		}
		// ---- quality-value p-values: per pool and overall ----
		double pall0 = 0, pall1 = 0;
		int qpvars = 0;
		if (a1 < 4 && a2 < 4) {
			for (int i = lane; i < P; i += 32) {
				int ci[16];
				double si[16];
				const int32_t* r = b.indcounts + ((size_t)s * P + i) * kNC;
#pragma unroll
				for (int q = 0; q < 16; q++) { ci[q] = r[q]; si[q] = b.stats ? b.stats[((size_t)s * P + i) * 16 + q] : 0.0; }
				if (row >= 0) {
					const int32_t* d = b.amb_dec + ((size_t)row * P + i) * 4;
					ci[refbase] -= d[0]; ci[refbase + 8] -= d[1];
					if (b.amb_ep) { si[refbase] -= b.amb_ep[((size_t)row * P + i) * 2]; si[refbase + 8] -= b.amb_ep[((size_t)row * P + i) * 2 + 1]; }
				}
				double q0, q1;
				d_chernoff(ci, si, a1, a2, ser, &q0, &q1);
				if (q0 <= c.min_qpv && q1 <= c.min_qpv) qpvars++;
			}
			qpvars = warp_sum(qpvars);
			d_chernoff(cnts, tst, a1, a2, pow(0.1, 0.1 * c.qmin), &pall0, &pall1);
		}
		const int C1 = cnts[a2], N1 = C1 + cnts[a1], C2 = cnts[a2 + 8], N2 = C2 + cnts[a1 + 8];
		const double strandpv = d_fet(C1, N1, C2, N2);
		int type = 0;
		if (ctpval <= c.thresh1) type = 1;
		else if (ctpval <= -3 && pall0 <= -5 && pall1 <= -5) type = 2;
		else if (ctpval <= -2 && pall0 + pall1 <= -10 && strandpv >= -1) type = 3;
		else if (strandpv >= -1 && pall0 + pall1 <= -20 && cnts[a2] >= 5 && cnts[a2 + 8] >= 5) type = 4;
		else if (qpvars >= 2) type = 5;
		int variants = 0, strandpass = 0;
		if (type > 0) {
			// ---- call_genotypes: per-pool read-evidence tests ----
			const uint32_t alo = b.alt_off ? b.alt_off[s * kSlots + slot] : 0, ahi = b.alt_off ? b.alt_off[s * kSlots + slot + 1] : 0;
			const int altal = (a1 == refbase || a2 != refbase) ? a2 : a1;
			int spools = 0;
			for (int i = lane; i < P; i += 32) {
				pvstrand[i] = 0.0;
				const int32_t* r = b.indcounts + ((size_t)s * P + i) * kNC;
				int dd[3] = {0, 0, 0};
				if (row >= 0) { const int32_t* d = b.amb_dec + ((size_t)row * P + i) * 4; dd[0] = d[0]; dd[1] = d[1]; dd[2] = d[2] + d[3]; }
				int x1[3], x2[3], x3[3];
#pragma unroll
				for (int k = 0; k < 3; k++) {
					x1[k] = r[a1 + 8 * k] - (a1 == refbase ? dd[k] : 0);
					x2[k] = r[a2 + 8 * k] - (a2 == refbase ? dd[k] : 0);
					x3[k] = r[a3 + 8 * k] - (a3 == refbase ? dd[k] : 0);
				}
				const int altf = x2[0], altr = x2[1], altb = x2[2];
				if (altf + altr + altb < 2) continue;
				const int n = c.ploidy[i];
				// contingency-table columns (allelecounts.c:214-229)
				const int c0 = x1[0] + x1[1] + x1[2] + x3[0] + x3[1] + x3[2], c1 = x2[0] + x2[1] + x2[2];
				const int ctN = c0 + c1, ctA = (altal == a2) ? c1 : c0;
				const int t0 = x1[0] + x2[0] + x3[0], t2 = x1[1] + x2[1] + x3[1];
				const int t1 = (altal == a2) ? x2[0] : x1[0], t3 = (altal == a2) ? x2[1] : x1[1];
				uint32_t lo = alo, hi = ahi;
				while (lo < hi) { uint32_t mid = (lo + hi) >> 1; if (b.alt_reads[mid].pool < i) lo = mid + 1; else hi = mid; }
				const uint32_t e0 = lo;
				hi = ahi;
				while (lo < hi) { uint32_t mid = (lo + hi) >> 1; if (b.alt_reads[mid].pool <= i) lo = mid + 1; else hi = mid; }
				int m = (int)(lo - e0);
				if (m > altf + altr + altb) m = altf + altr + altb;
				int uq, mid;
				unique_middle_serial(b.alt_reads + e0, m, uq, mid);
				const double pvallow = d_minreads_pvalue(ctN, ctA, 1.0 / n);
				double pvs = d_fet(altf, altf + x1[0], altr, altr + x1[1]);
				const double minpvlog = d_minreads_pvalue(t0, t1, 1.0 / n) + d_minreads_pvalue(t2, t3, 1.0 / n);
				const double minpvlog1 = d_minreads_pvalue(t0, t1, c.relax / n) + d_minreads_pvalue(t2, t3, c.relax / n);
				if (n > 2) {
					int pass = 0;
					if (minpvlog >= c.thresh3 || (minpvlog1 >= c.thresh3 && ctpval - minpvlog <= c.thresh1 && n >= 10) || (a2 >= 4 && minpvlog1 >= c.thresh3)) pass++;
					else if (minpvlog < c.thresh3 && (pvallow + pvs >= -3)) pass++;
					if ((t1 > 0 && t3 > 0 && uq >= 3) || uq >= 4) pass++;
					if (ctA >= c.min_reads) pass++;
					if (mid > 0) pass++;
					if (pass == 4) variants += 1;
				} else {
					if (minpvlog >= c.thresh3 && t1 > 0 && t3 > 0 && uq >= 2) variants += 1;
					else if (minpvlog >= c.thresh3 && ctA >= c.min_reads && uq >= 2) variants += 1;
					else if (minpvlog1 >= c.thresh3 && a2 >= 4 && ctA >= c.min_reads && uq >= 2) variants += 1;
					else if (minpvlog1 >= c.thresh3 && a2 < 4 && ctA >= 5 && uq >= 3) variants += 1;
				}
				if (minpvlog1 >= c.thresh3) spools++;
				else pvs = 0;
				pvstrand[i] = pvs;
			}
			variants = warp_sum(variants);
			spools = warp_sum(spools);
			__syncwarp();
			const double lim = log10(0.01) - log10(1.0 + spools);
			for (int i = lane; i < P; i += 32) if (pvstrand[i] < lim) strandpass++;
			strandpass = warp_sum(strandpass);
			__syncwarp();
		}
		if (lane == 0) {
			prm.o.qvpf[o] = pall0;
			prm.o.qvpr[o] = pall1;
			if (type > 0 && variants > 0) {
				prm.o.status[o] = CRISPGPU_SLOT_CALLED;
				prm.o.varpools[o] = variants;
				prm.o.variantpools[o] = variants;
				prm.o.allelecount[o] = 0;
			}
			// strandpass of this slot (+1), consumed by k_finalize: the site keeps the value of the last slot that ran call_genotypes
			prm.o.em_iters[o] = type > 0 ? (strandpass + 1) : 0;
		}
	}
}

// call_rank / varalleles per site (newcrispcaller.c:227-230), and the site's strandpass = value left by the last
// slot that ran call_genotypes
__global__ void k_finalize(EngineParams prm)
{
	const int n = prm.b.n_sites;
	for (int s = blockIdx.x * blockDim.x + threadIdx.x; s < n; s += gridDim.x * blockDim.x) {
		int rank = 0, sp = 0;
		for (int k = 0; k < kSlots; k++) {
			const int o = s * kSlots + k;
			if (prm.c.em_flag == 0) {
				const int v = prm.o.em_iters[o];
				if (v > 0) sp = v - 1;
				prm.o.em_iters[o] = 0;
			}
			if (prm.o.status[o] == CRISPGPU_SLOT_CALLED) prm.o.call_rank[o] = (int8_t)rank++;
		}
		prm.o.varalleles[s] = rank;
		prm.o.strandpass[s] = sp;
	}
}

}  // namespace crisp
