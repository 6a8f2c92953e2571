// dmath.cuh — scalar device functions of the CRISP statistics engine (sm_100a, FP64 CUDA cores).
//
// Each function states the reference lines whose *results* it reproduces (vibansal/crisp).  They are written for
// the device: no recursion, no libm state, branch structure chosen for warp execution.
#pragma once
#include <cstdint>

namespace crisp {

constexpr double kCtEps = 1e-6;  // `epsilon`, FET/contables.h:9
constexpr double kMinP = 1e-6;   // MINP, crisp/crispEM.c:6

// log10 C(n,r), r-term sum of log10 ratios in the reference's order (FET/contables.c:80-87)
__device__ __forceinline__ double d_ncr(int n, int r)
{
	if (r == 0 || n == r) return 0.0;
	if (2 * r > n) r = n - r;
	double ll = 0.0;
	for (int i = 0; i < r; i++) ll += log10((double)(n - i) / (double)(i + 1));
	return ll;
}

// ln Gamma(z), AS245 (FET/kfunc.c:9-22)
__device__ __forceinline__ double d_lgamma(double z)
{
	double x = 0.0;
	x += 0.1659470187408462e-06 / (z + 7);
	x += 0.9934937113930748e-05 / (z + 6);
	x -= 0.1385710331296526 / (z + 5);
	x += 12.50734324009056 / (z + 4);
	x -= 176.6150291498386 / (z + 3);
	x += 771.3234287757674 / (z + 2);
	x -= 1259.139216722289 / (z + 1);
	x += 676.5203681218835 / z;
	x += 0.9999999999995183;
	return log(x) - 5.58106146679532777 - z + (z - 0.5) * log(z + 6.5);
}

// ln of the regularised upper incomplete gamma Q(s,z) with the reference's branch rule, series length and
// continued-fraction tolerances (FET/kfunc.c:73-116).  Callers divide by ln 10.
__device__ inline double d_gammaq(double s, double z)
{
	if (z <= 1. || z < s) {
		double sum = 1., x = 1.;
		for (int k = 1; k < 100; ++k) {
			sum += (x *= z / (s + k));
			if (x < 1e-14 * sum) break;   // the reference's x / sum < 1e-14 without the division (sum >= 1): the same term count but for
			                              // quotients within an ulp of the threshold, where one more term of 1e-14 relative size is added
		}
		double lp = s * log(z) - z - d_lgamma(s + 1.) + log(sum);
		return log(1.0 - exp(lp));
	}
	double f = 1. + z - s, C = f, D = 0.;
	for (int j = 1; j < 100; ++j) {
		double a = j * (s - j), b = (j << 1) + 1 + z - s, d;
		D = b + a * D;
		if (D < 1e-290) D = 1e-290;
		C = b + a / C;
		if (C < 1e-290) C = 1e-290;
		D = 1. / D;
		d = C * D;
		f *= d;
		if (fabs(d - 1.) < 1e-14) break;
	}
	return s * log(z) - z - d_lgamma(s) - log(f);
}

// log10 p of a chi-square tail as the reference forms it: kf_gammaq(dof1/2, stat/2)/ln 10
__device__ __forceinline__ double d_chi2_log10p(double dof1, double stat) { return d_gammaq(dof1 / 2, stat / 2) / log(10.0); }

// lower binomial tail log10 P(X <= A) (FET/contables.c:197-209)
__device__ inline double d_minreads_pvalue(int R, int A, double e)
{
	double e1 = log10(e), e2 = log10(1.0 - e);
	double pvlog = R * e2, sum = pvlog;
	for (int r = 1; r <= A; r++) {
		pvlog += log10((double)(R - r + 1)) - log10((double)r) + e1 - e2;
		if (pvlog < sum) sum += log10(1.0 + exp10(pvlog - sum));
		else sum = pvlog + log10(1.0 + exp10(sum - pvlog));
	}
	return sum;
}

// upper binomial tail log10 P(X >= A) (FET/contables.c:232-242)
__device__ inline double d_binomial_pvalue(int R, int A, double e)
{
	double e1 = log10(e), e2 = log10(1.0 - e);
	double pvlog = d_ncr(R, A) + A * e1 + (R - A) * e2, sum = pvlog;
	for (int r = A + 1; r < R + 1; r++) {
		pvlog += log10((double)(R - r + 1)) - log10((double)r) + e1 - e2;
		sum += log10(1.0 + exp10(pvlog - sum));
	}
	return sum;
}

// two-sided 2x2 Fisher exact test, log10 p (FET/contables.c:131-171; the one-sided sums there feed nothing)
__device__ inline double d_fet(int c1, int r1, int c2, int r2)
{
	int minc = 0, maxc = c1 + c2;
	if (r2 < c1 + c2) minc = c1 + c2 - r2;
	if (r1 < c1 + c2) maxc = r1;
	const double ptable = d_ncr(r1, c1) + d_ncr(r2, c2);
	int c11 = minc, c21 = c1 + c2 - minc;
	double p1 = d_ncr(r1, c11), p2 = d_ncr(r2, c21), pvalue = -10000000;
	bool any = false;
	while (c11 <= maxc) {
		if (p1 + p2 <= ptable + kCtEps) {
			if (p1 + p2 <= pvalue) pvalue += log10(1.0 + exp10(p1 + p2 - pvalue));
			else pvalue = p1 + p2 + log10(1.0 + exp10(pvalue - p1 - p2));
			any = true;
		}
		if (r1 - c11 > 0) p1 += log10((double)(r1 - c11)) - log10((double)(c11 + 1));
		if (c21 > 0) p2 += log10((double)c21) - log10((double)(r2 - c21 + 1));
		c11 += 1;
		c21 -= 1;
	}
	if (!any) pvalue = ptable;
	return pvalue - d_ncr(r1 + r2, c1 + c2);
}

// Chernoff-bound quality-value p-values for one strand pair (FET/contables.c:342-364)
__device__ inline void d_chernoff(const int* counts, const double* stats, int refbase, int altbase, double ser, double* P0, double* P1)
{
	double p0, p1;
	if (ser == -1) {
		double mu0 = stats[altbase] + stats[refbase], mu1 = stats[altbase + 8] + stats[refbase + 8];
		double d0 = (counts[altbase] == 0 || mu0 == 0) ? 1 : counts[altbase] / mu0 - 1;
		double d1 = (counts[altbase + 8] == 0 || mu1 == 0) ? 1 : counts[altbase + 8] / mu1 - 1;
		p0 = d0 <= 0 ? 0 : mu0 * (d0 * log10(2.718281) - (1 + d0) * log10(1 + d0));
		p1 = d1 <= 0 ? 0 : mu1 * (d1 * log10(2.718281) - (1 + d1) * log10(1 + d1));
	} else {
		p0 = d_binomial_pvalue(counts[refbase] + counts[altbase], counts[altbase], ser);
		p1 = d_binomial_pvalue(counts[refbase + 8] + counts[altbase + 8], counts[altbase + 8], ser);
	}
	*P0 = p0;
	*P1 = p1;
}

// 10^x for the log-sum-exp terms (x in about [-30, 4]): k = round(x log2 10), r = x - k log10 2 (|r| <= 0.1506),
// 10^r by its Taylor polynomial to degree 12 (truncation 2e-16 relative), scaled by 2^k through the exponent field.
// The coefficients live in constant memory: as immediates every 64-bit constant costs two extra UMOV issue slots per
// DFMA (28 of the ~50 instructions of the first version).  Evaluated as two interleaved Horner chains (even / odd powers): 12 DFMA + 1 DMUL, dependent
// depth 8 instead of 13 — the callers are bound by issue slots and latency, not by the FP64 pipe.
// Max relative error ~4e-16.
__constant__ double kExp10C[16] = {
	1.0, 2.302585092994046, 2.650949055239199, 2.034678592293476, 1.171255148912267, 0.5393829291955814, 0.2069958486968681,
	0.06808936507443707, 0.019597694626478524, 0.00501392883377544, 0.0011544997789984348, 0.00024166672554424694,
	4.6371516642572196e-05,
	3.321928094887362, 0.3010299956639812, -2.8037281277851704e-18};
// The constants as values: a caller that evaluates 10^x in a loop loads them once (they are warp-uniform, so they can
// stay in uniform registers across the loop) and passes them in.
struct Exp10K { double c[16]; };
__device__ __forceinline__ Exp10K exp10_constants()
{
	Exp10K k;
#pragma unroll
	for (int i = 0; i < 16; i++) k.c[i] = kExp10C[i];
	return k;
}
__device__ __forceinline__ double exp10_lse(double x, const Exp10K& K)
{
	// k = round(x log2 10) without the FP64<->integer conversion instructions (they run on the quarter-rate XU pipe, which
	// the first version kept 84 % busy): adding 1.5 * 2^52 leaves the rounded integer in the low mantissa bits
	const double t = fma(x, K.c[13], 6755399441055744.0);
	const int k = __double2loint(t);
	const double kd = t - 6755399441055744.0;
	double r = fma(-kd, K.c[14], x);
	r = fma(-kd, K.c[15], r);
	// even and odd powers as two independent Horner chains in r^2: one constant operand per DFMA, dependent depth 8
	// instead of 13
	const double r2 = r * r;
	double pe = K.c[12], po = K.c[11];
	pe = fma(pe, r2, K.c[10]); po = fma(po, r2, K.c[9]);
	pe = fma(pe, r2, K.c[8]);  po = fma(po, r2, K.c[7]);
	pe = fma(pe, r2, K.c[6]);  po = fma(po, r2, K.c[5]);
	pe = fma(pe, r2, K.c[4]);  po = fma(po, r2, K.c[3]);
	pe = fma(pe, r2, K.c[2]);  po = fma(po, r2, K.c[1]);
	pe = fma(pe, r2, K.c[0]);
	const double p = fma(po, r, pe);
	return p * __hiloint2double((k + 1023) << 20, 0);
}
__device__ __forceinline__ double exp10_lse(double x) { return exp10_lse(x, exp10_constants()); }

// ---- Philox4x32-10 counter-based generator (one stream per (site, slot, permutation, stratum)) ----
struct Philox {
	uint32_t k0, k1;
	uint32_t c0, c1, c2;
	uint32_t out[4];
	int have;
	__device__ __forceinline__ void init(uint64_t seed, int tid, int pos, int slot, uint32_t perm, uint32_t stratum)
	{
		k0 = (uint32_t)seed ^ (uint32_t)pos;
		k1 = (uint32_t)(seed >> 32) ^ ((uint32_t)tid * 8u + (uint32_t)slot);
		c0 = perm;
		c1 = 0;
		c2 = stratum;
		have = 0;
	}
	__device__ __forceinline__ void refill()
	{
		uint32_t a = c0, b = c1, c = c2, d = 0, x = k0, y = k1;
#pragma unroll
		for (int r = 0; r < 10; r++) {
			uint32_t h0 = __umulhi(0xD2511F53u, a), l0 = 0xD2511F53u * a;
			uint32_t h1 = __umulhi(0xCD9E8D57u, c), l1 = 0xCD9E8D57u * c;
			a = h1 ^ b ^ x;
			b = l1;
			c = h0 ^ d ^ y;
			d = l0;
			x += 0x9E3779B9u;
			y += 0xBB67AE85u;
		}
		out[0] = a; out[1] = b; out[2] = c; out[3] = d;
		c1++;
		have = 4;
	}
	__device__ __forceinline__ uint32_t next()
	{
		if (have == 0) refill();
		uint32_t v = have == 4 ? out[0] : have == 3 ? out[1] : have == 2 ? out[2] : out[3];
		have--;
		return v;
	}
	// exact uniform integer in [0, M): multiply-shift with rejection
	__device__ __forceinline__ uint32_t below(uint32_t M)
	{
		uint64_t m = (uint64_t)next() * M;
		uint32_t l = (uint32_t)m;
		if (l < M) {
			uint32_t t = (0u - M) % M;
			while (l < t) {
				m = (uint64_t)next() * M;
				l = (uint32_t)m;
			}
		}
		return (uint32_t)(m >> 32);
	}
};

// exact uint32 -> double through the exponent field (2^52 + v, minus 2^52): integer pack + one DADD, no conversion
// instruction on the XU pipe
__device__ __forceinline__ double u32_to_double(uint32_t v) { return __hiloint2double(0x43300000, (int)v) - 4503599627370496.0; }

__device__ __forceinline__ double warp_sum(double v)
{
#pragma unroll
	for (int o = 16; o > 0; o >>= 1) v += __shfl_xor_sync(0xffffffffu, v, o);
	return v;
}
__device__ __forceinline__ int warp_sum(int v) { return __reduce_add_sync(0xffffffffu, v); }  // redux.sync: one instruction

}  // namespace crisp
