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

// log10 for the EM's own logarithms (the log-sum-exp of every (plane, pool) unit, log10 p and log10(1-p) of every iteration):
// they sit on the serial part of an iteration, where libdevice's log10 (~95 dependent instructions) was a tenth of the
// iteration's latency.  x = 2^e * m; the top seven mantissa bits pick c ~ m from a table holding rc = fl(1/c) and
// -log10(rc) (m >= 1.4140625 is halved first, so e = 0 and c = 1 around x = 1 and nothing cancels there); r = m*rc - 1
// exactly (one FMA), |r| < 2^-7; log10(1+r) by its Taylor polynomial to degree 9 (Estrin, dependent depth 6);
// e*log10(2) in two pieces, the high one exact.  Max error 2.3 ulp (3e-16 relative), 2e-18 absolute for |log10 x| < 0.01
// (tests/test_dmath.py runs these operations on the CPU against the 80-bit log10l; scripts/make_log10_table.py made the
// table).  Zero, negative, subnormal, infinite and NaN arguments take libdevice's log10.
__device__ const double kLog10Tab[256] = {
	1.0, 0.0,
	0.9884169884169884, 0.005059798769402255,
	0.9808429118773946, 0.008400542026431423,
	0.973384030418251, 0.01171578317790831,
	0.9660377358490566, 0.015005908624958285,
	0.9588014981273408, 0.018271296052725657,
	0.9516728624535316, 0.021512314690558438,
	0.9446494464944649, 0.024729325562556182,
	0.9377289377289377, 0.027922681728906475,
	0.9309090909090909, 0.03109272851841309,
	0.924187725631769, 0.034239803752598975,
	0.9175627240143369, 0.037364237961747995,
	0.9110320284697508, 0.04046635459323036,
	0.9045936395759717, 0.04354647021244068,
	0.8982456140350877, 0.04660489469666063,
	0.89198606271777, 0.049641931422142765,
	0.8858131487889274, 0.052657877444698284,
	0.8797250859106529, 0.055653023674057736,
	0.8737201365187713, 0.05862765504225989,
	0.8677966101694915, 0.06158205066631345,
	0.8619528619528619, 0.0645164840053628,
	0.8561872909698997, 0.06743122301258009,
	0.8504983388704319, 0.07032653028199379,
	0.8448844884488449, 0.07320266319045544,
	0.839344262295082, 0.07605987403493628,
	0.8338762214983714, 0.0788984101653369,
	0.8284789644012945, 0.08171851411298506,
	0.8231511254019293, 0.08452042371498796,
	0.8178913738019169, 0.08730437223459894,
	0.8126984126984127, 0.09007058847775094,
	0.807570977917981, 0.09281929690590195,
	0.8025078369905956, 0.09555071774533158,
	0.7975077881619937, 0.09826506709302253,
	0.7925696594427245, 0.1009625570192533,
	0.7876923076923077, 0.10364339566702482,
	0.7828746177370031, 0.1063077873484365,
	0.7781155015197568, 0.10895593263812474,
	0.7734138972809668, 0.11158802846386917,
	0.7687687687687688, 0.11420426819447031,
	0.764179104477612, 0.11680484172499567,
	0.7596439169139466, 0.11938993555948905,
	0.7551622418879056, 0.1219597328912326,
	0.750733137829912, 0.12451441368064815,
	0.7463556851311953, 0.12705415473092094,
	0.7420289855072464, 0.12957912976142455,
	0.7377521613832853, 0.13208950947902418,
	0.7335243553008596, 0.13458546164733035,
	0.7293447293447294, 0.1370671511539745,
	0.7252124645892352, 0.139534740075973,
	0.7211267605633803, 0.14198838774324452,
	0.7170868347338936, 0.14442825080034363,
	0.713091922005571, 0.14685448326646958,
	0.7091412742382271, 0.14926723659380836,
	1.4104683195592287, -0.14936333593971826,
	1.4027397260273973, -0.14697709651935606,
	1.3950953678474114, -0.1446038967237414,
	1.3875338753387534, -0.14224359481677037,
	1.3800539083557952, -0.1398960513607849,
	1.3726541554959786, -0.13756112916714316,
	1.3653333333333333, -0.13523869324811189,
	1.3580901856763925, -0.13292861077003787,
	1.3509234828496042, -0.1306307510077584,
	1.3438320209973753, -0.12834498530021146,
	1.3368146214099217, -0.126071187007208,
	1.3298701298701299, -0.12380923146733008,
	1.322997416020672, -0.12155899595691938,
	1.3161953727506426, -0.11932035965012298,
	1.3094629156010231, -0.11709320357996399,
	1.3027989821882953, -0.11487741060040409,
	1.2962025316455696, -0.11267286534937053,
	1.2896725440806045, -0.11047945421271568,
	1.2832080200501252, -0.1082970652890825,
	1.2768079800498753, -0.10612558835564843,
	1.2704714640198511, -0.10396491483472131,
	1.2641975308641975, -0.1018149377611622,
	1.257985257985258, -0.09967555175061071,
	1.2518337408312958, -0.09754665296848895,
	1.245742092457421, -0.09542813909976157,
	1.2397094430992737, -0.09331990931942975,
	1.2337349397590363, -0.09122186426373809,
	1.2278177458033572, -0.08913390600207322,
	1.2219570405727924, -0.08705593800953547,
	1.2161520190023754, -0.08498786514016247,
	1.210401891252955, -0.0829295936007884,
	1.204705882352941, -0.08088103092551918,
	1.199063231850117, -0.0788420859508069,
	1.1934731934731935, -0.07681266879110651,
	1.1879350348027842, -0.07479269081509914,
	1.1824480369515011, -0.07278206462246531,
	1.1770114942528735, -0.07078070402119342,
	1.17162471395881, -0.06878852400540891,
	1.1662870159453302, -0.06680544073370936,
	1.1609977324263039, -0.06483137150799223,
	1.1557562076749435, -0.06286623475276117,
	1.150561797752809, -0.0609099499948992,
	1.145413870246085, -0.05896243784389427,
	1.1403118040089086, -0.057023619972507565,
	1.1352549889135255, -0.05509341909787022,
	1.130242825607064, -0.05317175896299888,
	1.1252747252747253, -0.05125856431871835,
	1.1203501094091903, -0.049353760905980516,
	1.1154684095860568, -0.047457275438569556,
	1.1106290672451193, -0.045569035586182624,
	1.1058315334773219, -0.043688969957877646,
	1.1010752688172043, -0.041817008085876836,
	1.0963597430406853, -0.039953080409718615,
	1.091684434968017, -0.038097118260747485,
	1.0870488322717622, -0.036249053846934574,
	1.0824524312896406, -0.03440882023801921,
	1.0778947368421052, -0.03257635135096418,
	1.0733752620545074, -0.03075158193571686,
	1.068893528183716, -0.028934447561267493,
	1.0644490644490645, -0.027124884601999015,
	1.060041407867495, -0.02532283022431864,
	1.0556701030927835, -0.02352822237356711,
	1.051334702258727, -0.02174099976119644,
	1.047034764826176, -0.019961101852210554,
	1.0427698574338085, -0.018188468852862242,
	1.0385395537525355, -0.016423041698600736,
	1.0343434343434343, -0.014664762042262037,
	1.0301810865191148, -0.012913572242498677,
	1.0260521042084167, -0.011169415352440806,
	1.0219560878243512, -0.009432235108585003,
	1.0178926441351888, -0.007701975919903315,
	1.0138613861386139, -0.005978582857169383,
	1.009861932938856, -0.004262001642494803,
	1.005893909626719, -0.0025521786390719794,
	1.0, 0.0,
};
__constant__ double kLog10K[11] = {0.4342944819032518, -0.2171472409516259, 0.14476482730108395, -0.10857362047581295, 0.08685889638065036, -0.07238241365054197, 0.06204206884332169, -0.05428681023790648, 0.048254942433694645, 0.30102999566383914, 1.42050232272661e-13};
__device__ __forceinline__ double log10_fast(double x)
{
	const int hi = __double2hiint(x);
	if ((uint32_t)(hi - 0x00100000) >= 0x7FE00000u) return log10(x);
	const int i = (hi >> 13) & 0x7f;
	const int up = i >= 53 ? 1 : 0;
	const double m = __hiloint2double((hi & 0x000FFFFF) | (up ? 0x3FE00000 : 0x3FF00000), __double2loint(x));
	const double2 tc = __ldg(reinterpret_cast<const double2*>(kLog10Tab) + i);
	const double ed = u32_to_double((uint32_t)((hi >> 20) + up + 1)) - 1024.0;   // exponent - 1023 + up, without a conversion instruction
	const double r = fma(m, tc.x, -1.0);
	const double r2 = r * r;
	const double A = fma(kLog10K[1], r, kLog10K[0]), B = fma(kLog10K[3], r, kLog10K[2]), C = fma(kLog10K[5], r, kLog10K[4]), D = fma(kLog10K[7], r, kLog10K[6]);
	const double r4 = r2 * r2;
	const double AB = fma(B, r2, A), CD = fma(D, r2, C);
	const double in = fma(r4, kLog10K[8], CD);
	const double P = fma(r4, in, AB);
	const double t = fma(ed, kLog10K[9], tc.y);
	const double s = fma(ed, kLog10K[10], r * P);
	return t + s;
}

__device__ __forceinline__ double warp_sum(double v)
{
#pragma unroll
	for (int o = 16; o > 0; o >>= 1) v += __shfl_xor_sync(0xffffffffu, v, o);
	return v;
}
__device__ __forceinline__ int warp_sum(int v) { return __reduce_add_sync(0xffffffffu, v); }  // redux.sync: one instruction

}  // namespace crisp
