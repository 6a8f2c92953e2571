// k_evaluate.cuh — per-site screening kernel: the work of evaluate_variant (crisp/newcrispcaller.c:34-132) for
// every tested allele slot of every candidate site: initial allele-frequency guess, populate_contable
// (parsebam/allelecounts.c:208-261), the asymptotic contingency-table p-values chi2pvalue_stratified
// (FET/chisquare-stratified.c:12-203) / chi2pvalue (FET/chisquare.c:16-136) and the FAST_FILTER rejects.
//
// One warp per site, lanes over pools; the [P][24] count block (96 B per pool) is read straight from global memory.  Slots that survive are appended to the caller queue (or, with --chisquare 0, to the
// permutation queue; k_filter applies the rejects once the Monte-Carlo p-value is known).
#pragma once
#include "dmath.cuh"
#include "engine_types.cuh"

namespace crisp {

// resident threads per SM k_evaluate is compiled for: 1024 (64 registers) without a pivot sample, 512 (128 registers) with one
constexpr int kEvalThreadsPerSM = 1024;

__device__ __forceinline__ int slot_alleles(const uint8_t* mal, int slot, int& a1, int& a2, int& a3)
{
	a1 = mal[0];
	a2 = mal[slot + 1];
	a3 = mal[slot == 0 ? 2 : 1];  // newcrispcaller.c:208
	return 0;
}

// ambiguity row that remove_ambiguous_bases(+1) applies for this slot (newcrispcaller.c:212-215) or -1
__device__ __forceinline__ int slot_amb_row(const DevBatch& b, int s, int slot, int a2, int a3)
{
	if (!b.amb_idx || !b.hplen) return -1;
	int hp2 = b.hplen[s * 8 + a2], hp3 = b.hplen[s * 8 + a3];
	if ((a2 >= 4 && hp2 > 0) || (a3 >= 4 && hp3 >= 100)) return b.amb_idx[s * 5 + slot];
	return -1;
}

// the reject rules after the contingency-table p-value is known (newcrispcaller.c:108-127); true = keep
__device__ __forceinline__ bool pass_fast_filter(const DevCfg& c, const DevBatch& b, int s, int a2, double p0, double ctpval, double pv1, double pv2)
{
	if (a2 >= 4) {
		int ilen = b.indel_len ? abs((int)b.indel_len[s * 8 + a2]) : 1;
		if (ilen < 1) ilen = 1;
		int hpl = (b.hplen ? (int)b.hplen[s * 8 + a2] : 0) / ilen;
		if (c.fast_filter == 1 && ctpval >= -3 && c.ploidy0 > 2 && hpl >= 4 && (p0 / c.K) <= 0.05) return false;
		else if (((c.fast_filter == 1) & (ctpval >= -3)) && c.ploidy0 > 2 && hpl >= 8 && (p0 / c.K) <= 0.1) return false;
	}
	if (c.fast_filter == 1) {
		if (ctpval >= -3 && c.ploidy0 > 2 && p0 <= c.P / 4) return false;
		else if (ctpval >= -2 && c.ploidy0 > 2 && p0 <= c.P / 2) return false;
		else if (pv1 >= -2 && pv2 >= -2 && c.ploidy0 == 2 && p0 <= 5 && p0 <= c.P / 4) return false;
	}
	return true;
}

__device__ __forceinline__ size_t cell_row(int row, int P, size_t cell) { return (size_t)row * P + cell % P; }

__device__ __forceinline__ void queue_push(int32_t* list, int32_t* counter, int32_t v)
{
	int idx = atomicAdd(counter, 1);
	list[idx] = v;
}

// a slot that reached the caller: EM tasks are split by likelihood kind so that each gets a specialised kernel
__device__ __forceinline__ void push_call_task(const EngineParams& prm, int o, int a2)
{
	if (prm.c.em_flag >= 1 && (a2 >= 4 || prm.c.use_base_qvs == 0)) queue_push(prm.q.call_tasks2, prm.q.counters + 8, o);
	else {
		queue_push(prm.q.call_tasks, prm.q.counters + 1, o);
		if (prm.q.need_reads) prm.q.need_reads[o / kSlots] = 1;
	}
}

// p-value of the asymptotic test from the per-group joint statistics
__device__ inline double chi2_from_groups(int P, int pivot, double sj0, double sj1, bool stratified)
{
	if (pivot == 0) {
		// chi2pvalue_stratified: DOF = P-1 (:26); chi2pvalue: DOF = P-2+1 (:33-34); both use (DOF-1)/2
		double SF = sj0 + sj1;
		return SF > 0 ? d_chi2_log10p((double)(P - 2), SF) : 0.0;
	}
	int DOF = P - 2, DOF1 = pivot - 1, DOF2 = P - pivot - 1;
	double SF = sj0 + sj1;
	double pv2 = SF > 0 ? d_chi2_log10p((double)(DOF - 1), SF) : 0.0;
	double pv0 = sj0 > DOF1 - 1 ? d_chi2_log10p((double)(DOF1 - 1), sj0) : 0.0;
	double pv1 = sj1 > DOF2 - 1 ? d_chi2_log10p((double)(DOF2 - 1), sj1) : 0.0;
	if (pv0 < pv2 && pv0 < pv1) return pv0;
	if (pv1 < pv2) return pv1;
	return pv2;
}

// counts of one allele in the three strand classes for pool row r, ambiguity-adjusted
__device__ __forceinline__ void load3(const int32_t* __restrict__ r, int a, bool isref, const int* dd, int* x)
{
	x[0] = __ldg(r + a); x[1] = __ldg(r + a + 8); x[2] = __ldg(r + a + 16);
	if (isref) { x[0] -= dd[0]; x[1] -= dd[1]; x[2] -= dd[2]; }
}

// Qhighcounts-style strand totals (forward, reverse) of one allele for the weighted table
__device__ __forceinline__ void loadq(const DevBatch& b, const DevCfg& c, const int32_t* __restrict__ r, size_t cell, int a, bool isref, int row, double* q)
{
#pragma unroll
	for (int k = 0; k < 2; k++) {
		q[k] = b.qhigh ? __ldg(b.qhigh + cell * 16 + 8 * k + a) : (double)__ldg(r + a + 8 * k);
		if (isref && row >= 0) {
			const int32_t* d = b.amb_dec + cell_row(row, c.P, cell) * 4;
			double dec = d[k] + d[2 + k];
			if (c.use_qv == 1 && b.amb_ep) dec -= b.amb_ep[cell_row(row, c.P, cell) * 2 + k];
			q[k] -= dec;
		}
	}
}

// PERM: --chisquare 0 (the weighted table and the permutation queue instead of the stratified statistic); PIVOT: a pivot
// sample splits the pools into two groups.  Compile-time so that the common instantiation carries neither's registers.
template <int WARPS, bool PERM, bool PIVOT>
__global__ void __launch_bounds__(WARPS * 32, (PIVOT ? 512 : kEvalThreadsPerSM) / (WARPS * 32)) k_evaluate(EngineParams prm)
{
	const DevBatch& b = prm.b;
	const DevCfg& c = prm.c;
	const int P = c.P;
	const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
	const int pivot = PIVOT ? c.pivot_sample : 0;
	constexpr bool permmode = PERM;

	for (int s = blockIdx.x * WARPS + warp; s < b.n_sites; s += gridDim.x * WARPS) {
		// lanes run over pools and read their 96-byte count rows straight from global memory (each row is touched
		// by nine scalar loads per tested slot; the three 32-byte sectors of a row stay in L1 between them)
		const int32_t* __restrict__ ic = b.indcounts + (size_t)s * P * kNC;
		const uint8_t* mal = b.maxvec_al + s * 8;
		const int32_t* mct = b.maxvec_ct + s * 8;
		const int refbase = b.refbase[s];
		double my_p0 = 0, my_sj0 = 0, my_sj1 = 0;
		int my_tested = 0;
		for (int slot = 0; slot < kSlots; slot++) {
			if (mct[slot + 1] < 3) continue;
			int a1, a2, a3;
			slot_alleles(mal, slot, a1, a2, a3);
			const int row = slot_amb_row(b, s, slot, a2, a3);
			const int alt = (a1 == refbase || a2 != refbase) ? a2 : a1;
			// ---- pass 1: allele-frequency guess (newcrispcaller.c:48-68) and per-group strand totals ----
			double p0 = 0;
			int g0[6] = {0, 0, 0, 0, 0, 0}, g1[6] = {0, 0, 0, 0, 0, 0};
			double w0[4] = {0, 0, 0, 0}, w1[4] = {0, 0, 0, 0};
			for (int i = lane; i < P; i += 32) {
				const int32_t* r = ic + i * kNC;
				int dd[3] = {0, 0, 0};
				if (row >= 0) { const int32_t* d = b.amb_dec + ((size_t)row * P + i) * 4; dd[0] = d[0]; dd[1] = d[1]; dd[2] = d[2] + d[3]; }
				int x1[3], x2[3], x3[3];
				load3(r, a1, a1 == refbase, dd, x1);
				load3(r, a2, a2 == refbase, dd, x2);
				load3(r, a3, a3 == refbase, dd, x3);
				const int n = c.ploidy[i];
				const double C1 = x1[0] + x1[1] + x1[2], C2 = x2[0] + x2[1] + x2[2], C3 = x3[0] + x3[1] + x3[2];
				double p1 = C2 / (C1 + C2 + C3 + 0.001);
				p1 *= n;
				if (p1 >= 0.2 && C2 >= 4) p0 += p1;
				else if (p1 >= 0.4 && C2 >= 3) p0 += p1;
				else if (n == 2 && p1 >= 0.25 && C2 >= 1) p0 += p1;
				if (!permmode) {
					int* g = (pivot > 0 && i >= pivot) ? g1 : g0;
#pragma unroll
					for (int k = 0; k < 3; k++) { g[2 * k] += x1[k] + x2[k] + x3[k]; g[2 * k + 1] += x2[k]; }
				} else {
					// ctable_weighted columns (allelecounts.c:231-236): Qhighcounts by strand
					double* w = (i < pivot) ? w0 : w1;
					double q1[2], q2[2], q3[2], qa[2];
					const size_t cell = (size_t)s * P + i;
					loadq(b, c, r, cell, a1, a1 == refbase, row, q1);
					loadq(b, c, r, cell, a2, a2 == refbase, row, q2);
					loadq(b, c, r, cell, a3, a3 == refbase, row, q3);
					loadq(b, c, r, cell, alt, alt == refbase, row, qa);
					w[0] += q1[0] + q2[0] + q3[0]; w[1] += qa[0]; w[2] += q1[1] + q2[1] + q3[1]; w[3] += qa[1];
				}
			}
			p0 = warp_sum(p0);
			const bool af_ok = !(p0 < 0.25 || (c.ploidy0 == 2 && p0 <= 0.4));
			double sj0 = 0, sj1 = 0;
			if (af_ok && !permmode) {
				// ---- pass 2: chi2pvalue_stratified joint statistic (chisquare-stratified.c:44-61) ----
#pragma unroll
				for (int k = 0; k < 6; k++) g0[k] = warp_sum(g0[k]);
				if (pivot > 0) {
#pragma unroll
					for (int k = 0; k < 6; k++) g1[k] = warp_sum(g1[k]);
				}
				double Ef[2], Er[2], Eb[2], rvf[2], rvr[2], rvb[2];
				{
					const int* g = g0;
					for (int t = 0; t < 2; t++, g = g1) {
						Ef[t] = (g[1] + 0.25) / (g[0] + 0.25); Er[t] = (g[3] + 0.25) / (g[2] + 0.25); Eb[t] = (g[5] + 0.25) / (g[4] + 0.25);
						if (Ef[t] >= 0.999) Ef[t] = 0.999;
						if (Er[t] >= 0.999) Er[t] = 0.999;
						if (Eb[t] >= 0.999) Eb[t] = 0.999;
						// per-group reciprocal standard deviations per read; the bidirectional term pairs (1-Eb) with Er (:57)
						rvf[t] = rsqrt((1.0 - Ef[t]) * Ef[t]); rvr[t] = rsqrt((1.0 - Er[t]) * Er[t]); rvb[t] = rsqrt((1.0 - Eb[t]) * Er[t]);
						if (pivot == 0) break;
					}
				}
				const double tot = (double)g0[0] + g0[2] + g1[0] + g1[2] + g0[4] + g1[4];
				if (tot >= 25) {
					for (int i = lane; i < P; i += 32) {
						const int32_t* r = ic + i * kNC;
						int dd[3] = {0, 0, 0};
						if (row >= 0) { const int32_t* d = b.amb_dec + ((size_t)row * P + i) * 4; dd[0] = d[0]; dd[1] = d[1]; dd[2] = d[2] + d[3]; }
						int x1[3], x2[3], x3[3];
						load3(r, a1, a1 == refbase, dd, x1);
						load3(r, a2, a2 == refbase, dd, x2);
						load3(r, a3, a3 == refbase, dd, x3);
						const int t = (pivot > 0 && i >= pivot) ? 1 : 0;
						const int a = x1[0] + x2[0] + x3[0], bb = x1[1] + x2[1] + x3[1], cc = x1[2] + x2[2] + x3[2];
						if (a + bb + cc < 1) continue;
						// z = (a df + b dr + c db)/|(a,b,c)| with d* = (alt - E n)/sqrt(var n): n d* = sqrt(n) (alt - E n) / sqrt(var)
						double acc = 0.0;
						if (a > 0) acc += sqrt((double)a) * (x2[0] - Ef[t] * a) * rvf[t];
						if (bb > 0) acc += sqrt((double)bb) * (x2[1] - Er[t] * bb) * rvr[t];
						if (cc > 0) acc += sqrt((double)cc) * (x2[2] - Eb[t] * cc) * rvb[t];
						// the reference forms a*a+b*b+c*c in 32-bit int (chisquare-stratified.c:53): beyond ~33 000 reads per strand it
						// wraps, the square root turns NaN and the test reports p = 0 — reproduced with defined unsigned wrap-around
						const int ss = (int)((unsigned)a * (unsigned)a + (unsigned)bb * (unsigned)bb + (unsigned)cc * (unsigned)cc);
						const double z = acc * rsqrt((double)ss);
						if (t == 0) sj0 += z * z; else sj1 += z * z;
					}
					sj0 = warp_sum(sj0);
					if (pivot > 0) sj1 = warp_sum(sj1);
				}
			} else if (af_ok && permmode && c.ploidy0 > 2) {
				// ---- chi2pvalue on the weighted two-strand table (reported as chi2pval; FET/chisquare.c:62-94) ----
#pragma unroll
				for (int k = 0; k < 4; k++) { w0[k] = warp_sum(w0[k]); w1[k] = warp_sum(w1[k]); }
				double Ef[2], Er[2];
				Ef[0] = (w0[1] + 0.25) / (w0[0] + 0.25); Er[0] = (w0[3] + 0.25) / (w0[2] + 0.25);
				Ef[1] = (w1[1] + 0.25) / (w1[0] + 0.25); Er[1] = (w1[3] + 0.25) / (w1[2] + 0.25);
				for (int t = 0; t < 2; t++) { if (Ef[t] >= 0.999) Ef[t] = 0.999; if (Er[t] >= 0.999) Er[t] = 0.999; }
				const double wtot = w0[0] + w0[2] + w1[0] + w1[2];
				if (wtot >= 25) {
					for (int i = lane; i < P; i += 32) {
						const int32_t* r = ic + i * kNC;
						double q1[2], q2[2], q3[2], qa[2];
						const size_t cell = (size_t)s * P + i;
						loadq(b, c, r, cell, a1, a1 == refbase, row, q1);
						loadq(b, c, r, cell, a2, a2 == refbase, row, q2);
						loadq(b, c, r, cell, a3, a3 == refbase, row, q3);
						loadq(b, c, r, cell, alt, alt == refbase, row, qa);
						const double t0 = q1[0] + q2[0] + q3[0], t1 = qa[0], t2 = q1[1] + q2[1] + q3[1], t3 = qa[1];
						if (!(t0 + t2 >= 1)) continue;
						const int t = i < pivot ? 0 : 1;
						double df = 0, dr = 0;
						const double ww1 = t0 / sqrt(t0 * t0 + t2 * t2);
						const double ww2 = sqrt(1.0 - ww1 * ww1);
						if (t0 > 0) df = (t1 - Ef[t] * t0) / sqrt((1.0 - Ef[t]) * Ef[t] * t0);
						if (t2 > 0) dr = (t3 - Er[t] * t2) / sqrt((1.0 - Er[t]) * Er[t] * t2);
						const double z = ww1 * df + ww2 * dr;
						if (t == 0) sj0 += z * z; else sj1 += z * z;
					}
					sj0 = warp_sum(sj0);
					sj1 = warp_sum(sj1);
				}
			}
			if (lane == slot) {
				my_p0 = p0; my_sj0 = sj0; my_sj1 = sj1;
				my_tested = af_ok ? 2 : 1;
			}
		}
		// ---- lanes 0..4 leave their slot's statistics for k_eval_finish (p-value, filters, outputs: one thread per slot there) ----
		if (lane < kSlots) {
			const int o = s * kSlots + lane;
			prm.o.status[o] = (uint8_t)my_tested;   // 0 untested, 1 rejected by the allele-frequency guess, 2 statistics follow
			prm.o.paf[o] = my_p0;
			prm.o.ctpval[o] = my_sj0;
			prm.o.chi2pval[o] = my_sj1;
		}
	}
}

// appends v to a device queue for the lanes with pred set: one atomic per warp (all lanes of the warp must call it)
__device__ __forceinline__ void queue_push_warp(int32_t* list, int32_t* counter, int32_t v, bool pred)
{
	const unsigned m = __ballot_sync(0xffffffffu, pred);
	if (!m) return;
	const int lane = threadIdx.x & 31, leader = __ffs(m) - 1;
	int base = 0;
	if (lane == leader) base = atomicAdd(counter, __popc(m));
	base = __shfl_sync(0xffffffffu, base, leader);
	if (pred) list[base + __popc(m & ((1u << lane) - 1u))] = v;
}

// Second half of the screening: one thread per (site, slot).  The incomplete-gamma tail (a series or a continued fraction of
// up to 100 dependent steps, FET/kfunc.c:73-116) is the long pole of the screening; here 32 slots run it side by side in a
// warp instead of five lanes of a warp per site.  Also the FAST_FILTER rejects, the task queues and the zero fill of every
// per-slot result field.
template <bool PERM, bool PIVOT>
__global__ void __launch_bounds__(256) k_eval_finish(EngineParams prm)
{
	const DevBatch& b = prm.b;
	const DevCfg& c = prm.c;
	const int P = c.P, pivot = PIVOT ? c.pivot_sample : 0;
	const int total = b.n_sites * kSlots, padded = (total + 31) & ~31;
	for (int o = blockIdx.x * blockDim.x + threadIdx.x; o < padded; o += gridDim.x * blockDim.x) {
		const bool live = o < total;
		bool to_perm = false, to_call = false, to_call2 = false;
		if (live) {
			const int s = o / kSlots, slot = o - s * kSlots;
			const uint8_t* mal = b.maxvec_al + s * 8;
			const int a2 = mal[slot + 1];
			const int tested = prm.o.status[o];
			const double p0 = prm.o.paf[o], sj0 = prm.o.ctpval[o], sj1 = prm.o.chi2pval[o];
			uint8_t status = CRISPGPU_SLOT_UNTESTED;
			double ctp = 0, chi = 0;
			if (tested == 1) status = CRISPGPU_SLOT_REJECT_AF;
			else if (tested == 2) {
				if (!PERM) {
					ctp = chi = chi2_from_groups(P, pivot, sj0, sj1, true);
					status = pass_fast_filter(c, b, s, a2, p0, ctp, ctp, 0.0) ? CRISPGPU_SLOT_NOCALL : CRISPGPU_SLOT_REJECT_CT;
					if (status == CRISPGPU_SLOT_NOCALL) {
						if (c.em_flag >= 1 && (a2 >= 4 || c.use_base_qvs == 0)) to_call2 = true;
						else { to_call = true; if (prm.q.need_reads) prm.q.need_reads[s] = 1; }
					}
				} else {
					if (c.ploidy0 > 2) chi = chi2_from_groups(P, pivot, sj0, sj1, false);
					status = CRISPGPU_SLOT_REJECT_CT;  // provisional: k_filter decides once the permutation p-value exists
					to_perm = true;
				}
			}
			prm.o.status[o] = status;
			prm.o.allele[o] = (uint8_t)a2;
			prm.o.paf[o] = tested ? p0 : 0.0;
			prm.o.ctpval[o] = ctp;
			prm.o.chi2pval[o] = chi;
			prm.o.call_rank[o] = -1;
			prm.o.call_row[o] = -1;
			prm.o.af_ml[o] = 0; prm.o.delta[o] = 0; prm.o.deltaf[o] = 0; prm.o.hwe[o] = 0; prm.o.qvpf[o] = 0; prm.o.qvpr[o] = 0;
			prm.o.varpools[o] = 0; prm.o.allelecount[o] = 0; prm.o.variantpools[o] = 0;
			prm.o.em_iters[o] = 0; prm.o.perm_k[o] = 0; prm.o.perm_hits[o] = 0;
			prm.o.assoc_p[o] = 0; prm.o.assoc_llr[o] = 0; prm.o.assoc_or[o] = 0; prm.o.assoc_p1[o] = 0; prm.o.assoc_llr1[o] = 0; prm.o.assoc_or1[o] = 0;
			if (slot == 0) { prm.o.varalleles[s] = 0; prm.o.strandpass[s] = 0; }
		}
		if (PERM) queue_push_warp(prm.q.perm_tasks, prm.q.counters + 0, o, to_perm);
		else {
			queue_push_warp(prm.q.call_tasks, prm.q.counters + 1, o, to_call);
			queue_push_warp(prm.q.call_tasks2, prm.q.counters + 8, o, to_call2);
		}
	}
}

// Order of the EM task list.  k_em's CTAs take tasks from a shared counter; a task runs for 15-100 us (14 + 3.8 per EM
// iteration + 24 when the allele is called; CRISPGPU_EM_TRACE, scripts/em_trace.py) and the launch for a few hundred, so the
// launch ends one task after the last one was taken: in site order the SMs idle for 15 % of it at the benchmark shape.
// What can be known beforehand is the starting frequency paf (expected copies of the allele): below ~3 copies the allele is
// rarely called and the few slow fits (up to 22 iterations) are among the lowest; above, it is called and the read-evidence
// pass grows with the number of its reads.  The list is therefore put as: uncertain alleles in ascending paf, then the
// clear ones in descending paf (a counting sort on exponent + 3 mantissa bits of paf, one CTA, keys kept in registers) —
// the launch then ends on the lightest called alleles, which are all alike (simulated on the trace: 392 -> 364 us; 345 for
// an order by the true durations).  The order within a bucket — like the order of the list before — depends on timing;
// nothing downstream depends on it.
__global__ void __launch_bounds__(1024) k_order_tasks(const int32_t* __restrict__ in, int n, const double* __restrict__ paf, int32_t* __restrict__ out)
{
	__shared__ int32_t cur[256];
	constexpr int R = 8;             // tasks per thread held in registers (8192 per launch); the rest is read again
	constexpr int kClear = 93;       // key of paf = 3: (10 + log2 3) * 8
	const int tid = threadIdx.x, lane = tid & 31;
	if (tid < 256) cur[tid] = 0;
	__syncthreads();
	auto key_of = [&](int o) {
		int k = (__double2hiint(paf[o]) >> 17) - ((1023 - 10) << 3);   // 2^-10 .. 2^22, eight steps per octave
		k = k < 0 ? 0 : (k > 255 ? 255 : k);
		return k < kClear ? k : kClear + 255 - k;
	};
	int ov[R], kv[R];
#pragma unroll
	for (int r = 0; r < R; r++) { const int i = tid + r * 1024; ov[r] = i < n ? in[i] : -1; }
#pragma unroll
	for (int r = 0; r < R; r++) kv[r] = ov[r] >= 0 ? key_of(ov[r]) : -1;
#pragma unroll
	for (int r = 0; r < R; r++) if (kv[r] >= 0) atomicAdd(&cur[kv[r]], 1);
	for (int i = tid + R * 1024; i < n; i += 1024) atomicAdd(&cur[key_of(in[i])], 1);
	__syncthreads();
	if (tid < 32) {
		int v[8], sum = 0;
#pragma unroll
		for (int k = 0; k < 8; k++) { v[k] = cur[lane * 8 + k]; sum += v[k]; }
		int incl = sum;
		for (int o = 1; o < 32; o <<= 1) { const int t = __shfl_up_sync(0xffffffffu, incl, o); if (lane >= o) incl += t; }
		int run = incl - sum;
#pragma unroll
		for (int k = 0; k < 8; k++) { cur[lane * 8 + k] = run; run += v[k]; }
	}
	__syncthreads();
#pragma unroll
	for (int r = 0; r < R; r++) if (kv[r] >= 0) out[atomicAdd(&cur[kv[r]], 1)] = ov[r];
	for (int i = tid + R * 1024; i < n; i += 1024) { const int o = in[i]; out[atomicAdd(&cur[key_of(o)], 1)] = o; }
}

// Which quality characters occur in the batch's packed read stream: a 128-bit presence bitmap (counters[4..7]).
// HBM-streaming pass (2 B per read, 16-byte loads); lets k_em pick the dense likelihood path.
__global__ void k_qualmap(const uint16_t* __restrict__ reads, size_t n_reads, int32_t* counters)
{
	__shared__ unsigned char seen[kNQ];  // one flag per quality character; racing stores all write 1
	int bidir = 0;
	for (int q = threadIdx.x; q < kNQ; q += blockDim.x) seen[q] = 0;
	__syncthreads();
	const size_t nvec = n_reads / 8;
	const uint4* v = reinterpret_cast<const uint4*>(reads);
	for (size_t k = blockIdx.x * (size_t)blockDim.x + threadIdx.x; k < nvec; k += (size_t)gridDim.x * blockDim.x) {
		const uint4 x = __ldg(v + k);
		if ((x.x | x.y | x.z | x.w) & 0x08000800u) bidir = 1;  // bit 11 of either half: a bidirectional read
		seen[x.x & 0x7f] = 1; seen[(x.x >> 16) & 0x7f] = 1;
		seen[x.y & 0x7f] = 1; seen[(x.y >> 16) & 0x7f] = 1;
		seen[x.z & 0x7f] = 1; seen[(x.z >> 16) & 0x7f] = 1;
		seen[x.w & 0x7f] = 1; seen[(x.w >> 16) & 0x7f] = 1;
	}
	if (blockIdx.x == 0 && threadIdx.x < (int)(n_reads - nvec * 8)) {
		const uint32_t w = reads[nvec * 8 + threadIdx.x];
		seen[w & 0x7f] = 1;
		if (w & 0x800u) bidir = 1;
	}
	if (__any_sync(0xffffffffu, bidir) && (threadIdx.x & 31) == 0) atomicOr(reinterpret_cast<unsigned int*>(counters) + 10, 1u);
	__syncthreads();
	for (int q = threadIdx.x; q < kNQ; q += blockDim.x)
		if (seen[q]) atomicOr(reinterpret_cast<unsigned int*>(counters) + 4 + (q >> 5), 1u << (q & 31));
}

// On-demand read fetch: the packed reads stay in pinned host memory and only the sites that reached the
// quality-value likelihood are pulled over PCIe (zero-copy loads, 16 B per thread, four in flight) into the device
// buffer at their own offsets, so read_off needs no remapping.  Ranges are widened to 16-byte boundaries — the extra
// elements are the neighbouring sites' own reads, written with their own values.  The quality-presence bitmap and the
// bidirectional flag (k_qualmap's outputs) are taken from the fetched reads on the way.
__global__ void k_fetch_reads(const uint16_t* __restrict__ src, uint16_t* __restrict__ dst, const uint32_t* __restrict__ read_off,
                              const uint8_t* __restrict__ need, int n_sites, int P, size_t n_reads, int32_t* counters)
{
	__shared__ unsigned char seen[kNQ];
	int bidir = 0;
	unsigned long long moved = 0;
	for (int q = threadIdx.x; q < kNQ; q += blockDim.x) seen[q] = 0;
	__syncthreads();
	const uint4* s4 = reinterpret_cast<const uint4*>(src);
	uint4* d4 = reinterpret_cast<uint4*>(dst);
	const size_t nvec = n_reads / 8;
	const int T = blockDim.x, tid = threadIdx.x;
	for (int s = blockIdx.x; s < n_sites; s += gridDim.x) {
		if (!need[s]) continue;
		const size_t r0 = read_off[(size_t)s * P], r1 = read_off[(size_t)(s + 1) * P];
		if (r1 <= r0) continue;
		const size_t v0 = r0 / 8;
		size_t v1 = (r1 + 7) / 8;
		if (v1 > nvec) {  // the array ends inside the last vector: scalar tail
			v1 = nvec;
			for (size_t r = nvec * 8 + tid; r < r1; r += T) {
				const uint16_t w = src[r];
				dst[r] = w;
				seen[w & 0x7f] = 1;
				if (w & 0x800u) bidir = 1;
			}
		}
		for (size_t k = v0 + tid; k < v1; k += 4 * (size_t)T) {
			uint4 x[4];
#pragma unroll
			for (int u = 0; u < 4; u++)
				if (k + (size_t)u * T < v1) x[u] = __ldcs(s4 + k + (size_t)u * T);
#pragma unroll
			for (int u = 0; u < 4; u++) {
				if (k + (size_t)u * T >= v1) break;
				d4[k + (size_t)u * T] = x[u];
				if ((x[u].x | x[u].y | x[u].z | x[u].w) & 0x08000800u) bidir = 1;
				seen[x[u].x & 0x7f] = 1; seen[(x[u].x >> 16) & 0x7f] = 1;
				seen[x[u].y & 0x7f] = 1; seen[(x[u].y >> 16) & 0x7f] = 1;
				seen[x[u].z & 0x7f] = 1; seen[(x[u].z >> 16) & 0x7f] = 1;
				seen[x[u].w & 0x7f] = 1; seen[(x[u].w >> 16) & 0x7f] = 1;
			}
		}
		if (tid == 0) moved += r1 - r0;
	}
	if (moved) atomicAdd(reinterpret_cast<unsigned long long*>(counters + 14), moved);
	if (__any_sync(0xffffffffu, bidir) && (threadIdx.x & 31) == 0) atomicOr(reinterpret_cast<unsigned int*>(counters) + 10, 1u);
	__syncthreads();
	for (int q = threadIdx.x; q < kNQ; q += blockDim.x)
		if (seen[q]) atomicOr(reinterpret_cast<unsigned int*>(counters) + 4 + (q >> 5), 1u << (q & 31));
}

// Sparse per-pool allele counts -> the dense [n][P][24] table the kernels index: one thread per cell zeroes its 96-byte
// row (six 16-byte stores) and then drops the cell's non-zero entries into it.
__global__ void k_expand_counts(const uint32_t* __restrict__ ic_off, const uint32_t* __restrict__ ic_sparse, int32_t* __restrict__ indcounts, size_t n_cells)
{
	for (size_t c = blockIdx.x * (size_t)blockDim.x + threadIdx.x; c < n_cells; c += (size_t)gridDim.x * blockDim.x) {
		const uint32_t e0 = ic_off[c], e1 = ic_off[c + 1];
		int32_t* row = indcounts + c * kNC;
		int4* row4 = reinterpret_cast<int4*>(row);
#pragma unroll
		for (int k = 0; k < kNC / 4; k++) row4[k] = make_int4(0, 0, 0, 0);
		for (uint32_t e = e0; e < e1; e++) {
			const uint32_t x = __ldg(ic_sparse + e);
			const uint32_t idx = x >> 27;
			if (idx < (uint32_t)kNC) row[idx] = (int32_t)(x & 0x7ffffffu);
		}
	}
}

// ---- narrow transport form (crispgpu_batch.bins16 / ic16): widened on the device into the compact arrays ----

// One warp per site: the cells' entry counts (one byte each) become 32-bit offsets into the site's run of entries.
__global__ void k_narrow_offsets(const uint8_t* __restrict__ len, const uint32_t* __restrict__ site, uint32_t* __restrict__ off, int n_sites, int P)
{
	const int lane = threadIdx.x & 31, wpb = blockDim.x >> 5;
	for (int s = blockIdx.x * wpb + (threadIdx.x >> 5); s < n_sites; s += gridDim.x * wpb) {
		uint32_t run = site[s];
		for (int i0 = 0; i0 < P; i0 += 32) {
			const int i = i0 + lane;
			const uint32_t v = i < P ? len[(size_t)s * P + i] : 0u;
			uint32_t inc = v;
#pragma unroll
			for (int d = 1; d < 32; d <<= 1) { const uint32_t u = __shfl_up_sync(0xffffffffu, inc, d); if (lane >= d) inc += u; }
			if (i < P) off[(size_t)s * P + i] = run + inc - v;
			run += __shfl_sync(0xffffffffu, inc, 31);
		}
		if (s == n_sites - 1 && lane == 0) off[(size_t)n_sites * P] = run;
	}
}

// 16-bit bins -> the (value, count) words k_em reads, eight per thread and trip (one 16-byte load, two 16-byte stores); the
// quality-presence bitmap and the bidirectional flag (k_scan_bins' outputs) are taken on the way.
__global__ void k_unpack_bins16(const uint16_t* __restrict__ bins16, uint32_t* __restrict__ bins, size_t n_bins, int32_t* counters)
{
	__shared__ unsigned char seen[kNQ];
	int bidir = 0;
	for (int q = threadIdx.x; q < kNQ; q += blockDim.x) seen[q] = 0;
	__syncthreads();
	auto widen = [&](uint32_t e) -> uint32_t {
		seen[e & 0x7f] = 1;
		if (e & 0x400u) bidir = 1;
		return (e & 0x7fu) | (((e >> 7) & 0xfu) << 8) | ((((e >> 11) & 0x1fu) + 1u) << 16);
	};
	const size_t nvec = n_bins / 8;
	const uint4* v = reinterpret_cast<const uint4*>(bins16);
	uint4* o = reinterpret_cast<uint4*>(bins);
	for (size_t k = blockIdx.x * (size_t)blockDim.x + threadIdx.x; k < nvec; k += (size_t)gridDim.x * blockDim.x) {
		const uint4 x = __ldg(v + k);
		o[2 * k] = make_uint4(widen(x.x & 0xffffu), widen(x.x >> 16), widen(x.y & 0xffffu), widen(x.y >> 16));
		o[2 * k + 1] = make_uint4(widen(x.z & 0xffffu), widen(x.z >> 16), widen(x.w & 0xffffu), widen(x.w >> 16));
	}
	if (blockIdx.x == 0 && threadIdx.x < (int)(n_bins - nvec * 8)) bins[nvec * 8 + threadIdx.x] = widen(bins16[nvec * 8 + threadIdx.x]);
	if (__any_sync(0xffffffffu, bidir) && (threadIdx.x & 31) == 0) atomicOr(reinterpret_cast<unsigned int*>(counters) + 10, 1u);
	__syncthreads();
	for (int q = threadIdx.x; q < kNQ; q += blockDim.x)
		if (seen[q]) atomicOr(reinterpret_cast<unsigned int*>(counters) + 4 + (q >> 5), 1u << (q & 31));
}

// Narrow sparse per-pool allele counts -> the dense [n][P][24] table: one warp per site finds its cells' entries from the
// length bytes, every lane zeroes the 96-byte rows of its cells and adds their entries (entries of one index add up).
// counts != null: the site totals (variant->counts = the column sums of the per-pool rows) are written as well.
__global__ void k_expand_counts16(const uint8_t* __restrict__ len, const uint32_t* __restrict__ site, const uint16_t* __restrict__ ic16,
                                  int32_t* __restrict__ indcounts, int32_t* __restrict__ counts, int n_sites, int P)
{
	__shared__ int tot[8][kNC];
	const int lane = threadIdx.x & 31, wpb = blockDim.x >> 5, wib = threadIdx.x >> 5;
	for (int s = blockIdx.x * wpb + (threadIdx.x >> 5); s < n_sites; s += gridDim.x * wpb) {
		if (counts && lane < kNC) tot[wib][lane] = 0;
		__syncwarp();
		uint32_t run = site[s];
		for (int i0 = 0; i0 < P; i0 += 32) {
			const int i = i0 + lane;
			const uint32_t v = i < P ? len[(size_t)s * P + i] : 0u;
			uint32_t inc = v;
#pragma unroll
			for (int d = 1; d < 32; d <<= 1) { const uint32_t u = __shfl_up_sync(0xffffffffu, inc, d); if (lane >= d) inc += u; }
			if (i < P) {
				int32_t* row = indcounts + ((size_t)s * P + i) * kNC;
				int4* row4 = reinterpret_cast<int4*>(row);
#pragma unroll
				for (int k = 0; k < kNC / 4; k++) row4[k] = make_int4(0, 0, 0, 0);
				// entries of one index follow one another (a count above 2047 is split in place): their sum is kept in a register
				// and stored once, so the freshly zeroed row is never read back
				const uint32_t e0 = run + inc - v;
				uint32_t prev = 0xffffffffu;
				int acc = 0;
				for (uint32_t e = e0; e < e0 + v; e++) {
					const uint32_t x = __ldg(ic16 + e);
					const uint32_t idx = x >> 11;
					if (idx >= (uint32_t)kNC) continue;
					if (idx != prev) {
						if (prev != 0xffffffffu) { row[prev] = acc; if (counts) atomicAdd(&tot[wib][prev], acc); }
						prev = idx;
						acc = 0;
					}
					acc += (int)(x & 0x7ffu);
				}
				if (prev != 0xffffffffu) { row[prev] = acc; if (counts) atomicAdd(&tot[wib][prev], acc); }
			}
			run += __shfl_sync(0xffffffffu, inc, 31);
		}
		__syncwarp();
		if (counts && lane < kNC) counts[(size_t)s * kNC + lane] = tot[wib][lane];
		__syncwarp();
	}
}

// 4-byte alt-read records (crispgpu_batch.alt32) -> the 8-byte records the pool statistics read
__global__ void k_unpack_alt32(const uint32_t* __restrict__ alt32, crispgpu_altread* __restrict__ out, size_t n)
{
	for (size_t k = blockIdx.x * (size_t)blockDim.x + threadIdx.x; k < n; k += (size_t)gridDim.x * blockDim.x) {
		const uint32_t x = __ldg(alt32 + k);
		crispgpu_altread e;
		e.pool = (uint16_t)(x & 0x7ffu); e.offset = (uint16_t)((x >> 11) & 0x3ffu); e.aligned = (uint16_t)((x >> 21) & 0x3ffu); e.strand = (uint8_t)(x >> 31); e.pad = 0;
		out[k] = e;
	}
}

// Candidate screen (variantcalls.c:62-92): one warp per position.  Lanes sum the pools' contributions to the
// potential-variant score; lane 0 builds maxvec exactly as sort_allelecounts does (allelecounts.c:134-183, IUPAC off):
// read counts per allele, reference allele swapped to the front, then the exchange sort `if (ct[i] < ct[j]) swap` over
// i = 1..6, j = i+1..7 — the same comparisons in the same order, so ties fall the same way.
__global__ void k_candidates(const uint8_t* __restrict__ refbase, const int32_t* __restrict__ counts, const int32_t* __restrict__ indcounts,
                             const int32_t* __restrict__ ploidy, int n_sites, int P, uint8_t* __restrict__ maxvec_al, int32_t* __restrict__ maxvec_ct,
                             int32_t* __restrict__ potential)
{
	const int lane = threadIdx.x & 31, wpb = blockDim.x >> 5;
	for (int s = blockIdx.x * wpb + (threadIdx.x >> 5); s < n_sites; s += gridDim.x * wpb) {
		const int ref = refbase[s];
		if (potential) {
			int pot = 0;
			for (int i = lane; i < P; i += 32) {
				const int32_t* r = indcounts + ((size_t)s * P + i) * kNC;
				const bool diploid = ploidy[i] == 2;
#pragma unroll
				for (int a = 0; a < 8; a++) {
					if (a == ref) continue;
					const int t = r[a] + r[a + 8] + r[a + 16];
					if (t >= 3) pot += 2;
					else if (t >= 1 && diploid) pot += t;
				}
			}
			pot = warp_sum(pot);
			if (lane == 0) potential[s] = pot;
		}
		if (lane == 0) {
			int ct[8], al[8];
#pragma unroll
			for (int a = 0; a < 8; a++) { ct[a] = counts[s * kNC + a] + counts[s * kNC + a + 8] + counts[s * kNC + a + 16]; al[a] = a; }
			if (ref >= 1 && ref < 8) {
				// swap entry 0 with the reference allele's entry (static indices only: the arrays live in registers)
#pragma unroll
				for (int a = 1; a < 8; a++)
					if (a == ref) { const int tc = ct[0]; ct[0] = ct[a]; ct[a] = tc; const int ta = al[0]; al[0] = al[a]; al[a] = ta; }
			}
#pragma unroll
			for (int i = 1; i < 7; i++)
#pragma unroll
				for (int j = i + 1; j < 8; j++)
					if (ct[i] < ct[j]) { const int tc = ct[i]; ct[i] = ct[j]; ct[j] = tc; const int ta = al[i]; al[i] = al[j]; al[j] = ta; }
#pragma unroll
			for (int a = 0; a < 8; a++) { maxvec_al[s * 8 + a] = (uint8_t)al[a]; maxvec_ct[s * 8 + a] = ct[a]; }
		}
	}
}

// Which qualities occur (and whether any read is bidirectional) in a compact batch: k_qualmap's outputs read from the bins
// themselves (a tenth of the read stream), flat 16-byte loads over the whole bin array.
__global__ void k_scan_bins(const uint32_t* __restrict__ bins, size_t n_bins, int32_t* counters)
{
	__shared__ unsigned char seen[kNQ];
	int bidir = 0;
	for (int q = threadIdx.x; q < kNQ; q += blockDim.x) seen[q] = 0;
	__syncthreads();
	const size_t nvec = n_bins / 4;
	const uint4* v = reinterpret_cast<const uint4*>(bins);
	for (size_t k = blockIdx.x * (size_t)blockDim.x + threadIdx.x; k < nvec; k += (size_t)gridDim.x * blockDim.x) {
		const uint4 x = __ldg(v + k);
		if ((x.x | x.y | x.z | x.w) & 0x800u) bidir = 1;
		seen[x.x & 0x7f] = 1; seen[x.y & 0x7f] = 1; seen[x.z & 0x7f] = 1; seen[x.w & 0x7f] = 1;
	}
	if (blockIdx.x == 0 && threadIdx.x < (int)(n_bins - nvec * 4)) {
		const uint32_t e = bins[nvec * 4 + threadIdx.x];
		seen[e & 0x7f] = 1;
		if (e & 0x800u) bidir = 1;
	}
	if (__any_sync(0xffffffffu, bidir) && (threadIdx.x & 31) == 0) atomicOr(reinterpret_cast<unsigned int*>(counters) + 10, 1u);
	__syncthreads();
	for (int q = threadIdx.x; q < kNQ; q += blockDim.x)
		if (seen[q]) atomicOr(reinterpret_cast<unsigned int*>(counters) + 4 + (q >> 5), 1u << (q & 31));
}

// After the permutation kernel: apply the FAST_FILTER rejects with the Monte-Carlo p-value and queue the survivors.
__global__ void k_filter(EngineParams prm)
{
	const int n = prm.q.counters[0];
	for (int t = blockIdx.x * blockDim.x + threadIdx.x; t < n; t += gridDim.x * blockDim.x) {
		const int o = prm.q.perm_tasks[t], s = o / kSlots, slot = o - s * kSlots;
		const int a2 = prm.b.maxvec_al[s * 8 + slot + 1];
		const double ctp = prm.o.ctpval[o];
		// diploid special filter uses pv1 = permutation p and pv2 = 0 (newcrispcaller.c:92-93,126)
		bool keep = pass_fast_filter(prm.c, prm.b, s, a2, prm.o.paf[o], ctp, ctp, prm.c.ploidy0 > 2 ? prm.o.chi2pval[o] : 0.0);
		if (keep) {
			prm.o.status[o] = CRISPGPU_SLOT_NOCALL;
			push_call_task(prm, o, a2);
		}
	}
}

}  // namespace crisp
