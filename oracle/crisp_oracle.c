/*
 * crisp_oracle.c — TEST INFRASTRUCTURE ONLY: a plain-C, single-threaded CPU restatement of CRISP's per-site
 * cross-pool statistics engine, written from the algorithm, working on the flat
 * struct-of-arrays batch of include/crispgpu.h instead of `struct VARIANT` + the read queue.
 *
 * It is the checker for the CUDA engine (tests/, __graft_entry__.smoke(), bench.py's cpu_baseline leg) and is
 * never imported, linked or executed by the product path (crisp_b200/).  Parity is PINNED: the reference has no
 * tests or golden vectors of its own, so tests/test_oracle.py checks every function
 * here against the unmodified reference compiled from /root/reference (oracle/_ref, see oracle/Makefile) and
 * against the committed fixtures in tests/golden/ that were generated from it.
 *
 * Each function cites the reference lines it restates.  Arithmetic is kept in the reference's order wherever a
 * result is compared or thresholded, so outputs agree to the last bit with the reference in nearly all cases.
 */
#define _GNU_SOURCE
#include <math.h>
#include <stdint.h>
#include <stdio.h>
#include <stdlib.h>
#include <string.h>

#include "oracle_abi.h"

#define NA 8 /* maxalleles */
#define MINP 1e-6
#define CT_EPS 1e-6 /* `epsilon`, FET/contables.h:9 */

/* ------------------------------------------------------------------------------------------------------------
 * scalar functions
 * ---------------------------------------------------------------------------------------------------------- */

/* log10 C(n,r) by r-term sum — FET/contables.c:80-87 */
static double o_ncr(int n, int r)
{
	if (r == 0 || n == r) return 0;
	if (2 * r > n) r = n - r;
	double ll = 0;
	for (int i = 0; i < r; i++) ll += log10((double)(n - i) / (i + 1));
	return ll;
}


/* two-sided 2x2 Fisher exact test in log10 space — FET/contables.c:131-171 */
static double o_fet(int c1, int r1, int c2, int r2)
{
	int minc = 0, maxc = c1 + c2;
	if (r2 < c1 + c2) minc = c1 + c2 - r2;
	if (r1 < c1 + c2) maxc = r1;
	double ptable = o_ncr(r1, c1) + o_ncr(r2, c2);
	int c11 = minc, c21 = c1 + c2 - minc;
	double p1 = o_ncr(r1, c11), p2 = o_ncr(r2, c21);
	double pvalue = -10000000;
	int flag = 0;
	while (c11 <= maxc) {
		if (p1 + p2 <= ptable + CT_EPS) {
			if (p1 + p2 <= pvalue) pvalue += log10(1.0 + pow(10, p1 + p2 - pvalue));
			else pvalue = p1 + p2 + log10(1.0 + pow(10, pvalue - p1 - p2));
			flag = 1;
		}
		if (r1 - c11 > 0) p1 += log10(r1 - c11) - log10(c11 + 1);
		if (c21 > 0) p2 += log10(c21) - log10(r2 - c21 + 1);
		c11 += 1;
		c21 -= 1;
	}
	if (flag == 0) pvalue = ptable;
	return pvalue - o_ncr(r1 + r2, c1 + c2);
}

/* lower binomial tail P(X <= A), X~Bin(R,e), log10 — FET/contables.c:197-209 */
static double o_minreads_pvalue(int R, int A, double e)
{
	double e1 = log10(e), e2 = log10(1.0 - e);
	double pvlog = R * e2, sum = pvlog;
	for (int r = 1; r <= A; r++) {
		pvlog += log10(R - r + 1) - log10(r) + e1 - e2;
		if (pvlog < sum) sum += log10(1.0 + pow(10, pvlog - sum));
		else sum = pvlog + log10(1.0 + pow(10, sum - pvlog));
	}
	return sum;
}

/* upper binomial tail P(X >= A), log10 — FET/contables.c:232-242 */
static double o_binomial_pvalue(int R, int A, double e)
{
	double e1 = log10(e), e2 = log10(1.0 - e);
	double pvlog = o_ncr(R, A) + A * e1 + (R - A) * e2, sum = pvlog;
	for (int r = A + 1; r < R + 1; r++) {
		pvlog += log10(R - r + 1) - log10(r) + e1 - e2;
		sum += log10(1.0 + pow(10, pvlog - sum));
	}
	return sum;
}

/* Chernoff-bound quality-value p-values per strand — FET/contables.c:342-364 */
static void o_chernoff(const int* counts, const double* stats, int refbase, int altbase, double ser, double* P0, double* P1)
{
	double p0, p1;
	if (ser == -1) {
		double mu0 = stats[altbase] + stats[refbase], mu1 = stats[altbase + NA] + stats[refbase + NA];
		double d0 = (counts[altbase] == 0 || mu0 == 0) ? 1 : counts[altbase] / mu0 - 1;
		double d1 = (counts[altbase + NA] == 0 || mu1 == 0) ? 1 : counts[altbase + NA] / mu1 - 1;
		p0 = d0 <= 0 ? 0 : mu0 * (d0 * log10(2.718281) - (1 + d0) * log10(1 + d0));
		p1 = d1 <= 0 ? 0 : mu1 * (d1 * log10(2.718281) - (1 + d1) * log10(1 + d1));
	} else {
		p0 = o_binomial_pvalue(counts[refbase] + counts[altbase], counts[altbase], ser);
		p1 = o_binomial_pvalue(counts[refbase + NA] + counts[altbase + NA], counts[altbase + NA], ser);
	}
	*P0 = p0;
	*P1 = p1;
}

/* ln Gamma — FET/kfunc.c:9-22 (AS245) */
static double o_lgamma(double z)
{
	double x = 0;
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

/* ln of the regularised lower incomplete gamma, series — FET/kfunc.c:73-83 */
static double o_gammap_series(double s, double z)
{
	double sum = 1., x = 1.;
	for (int k = 1; k < 100; ++k) {
		sum += (x *= z / (s + k));
		if (x / sum < 1e-14) break;
	}
	return s * log(z) - z - o_lgamma(s + 1.) + log(sum);
}

/* ln of the regularised upper incomplete gamma, modified Lentz continued fraction — FET/kfunc.c:85-105 */
static double o_gammaq_cf(double s, double z)
{
	double f = 1. + z - s, Cc = f, D = 0.;
	for (int j = 1; j < 100; ++j) {
		double a = j * (s - j), b = (j << 1) + 1 + z - s, d;
		D = b + a * D;
		if (D < 1e-290) D = 1e-290;
		Cc = b + a / Cc;
		if (Cc < 1e-290) Cc = 1e-290;
		D = 1. / D;
		d = Cc * D;
		f *= d;
		if (fabs(d - 1.) < 1e-14) break;
	}
	return s * log(z) - z - o_lgamma(s) - log(f);
}

/* ln Q(s,z) — FET/kfunc.c:112-116 */
static double o_gammaq(double s, double z)
{
	return (z <= 1. || z < s) ? log(1.0 - exp(o_gammap_series(s, z))) : o_gammaq_cf(s, z);
}

/* ------------------------------------------------------------------------------------------------------------
 * per-site working state
 * ---------------------------------------------------------------------------------------------------------- */
typedef struct oracle_ctx {
	crispgpu_config cfg;
	int P, maxploidy, K;
	int* ploidy;
	int* phenotypes;
	double** NC;      /* log10 C(n_i, j), parsebam/init_variant.c:19-25 */
	int* ic;          /* [P][24] working copy of indcounts (ambiguity-adjusted per slot) */
	int counts[24];
	double* qhigh;    /* [P][16] */
	double* stats;    /* [P][16] */
	double tstats[16];
	int* ctab;        /* [P][6]  ctable_stranded */
	double* ctabw;    /* [P][4]  ctable_weighted */
	int* ctab2;       /* [P][2]  ctable (N, alt) */
	double *G, *Gf, *Gr, *Gb; /* [P][maxploidy+1] */
	int rng_kind;
	char err[256];
} oracle_ctx;

#define IC(c, i, a) ((c)->ic[(i) * 24 + (a)])
#define GX(c, A, i, j) ((A)[(size_t)(i) * ((c)->maxploidy + 1) + (j)])

void* crisp_oracle_create(const crispgpu_config* cfg)
{
	if (!cfg || cfg->n_pools <= 0 || !cfg->ploidy) return NULL;
	oracle_ctx* c = (oracle_ctx*)calloc(1, sizeof(*c));
	c->cfg = *cfg;
	int P = c->P = cfg->n_pools;
	c->ploidy = (int*)malloc(sizeof(int) * P);
	c->phenotypes = (int*)calloc(P, sizeof(int));
	c->NC = (double**)malloc(sizeof(double*) * P);
	for (int i = 0; i < P; i++) {
		int n = c->ploidy[i] = cfg->ploidy[i];
		if (cfg->phenotypes) c->phenotypes[i] = cfg->phenotypes[i];
		if (n > c->maxploidy) c->maxploidy = n;
		c->K += n;
		c->NC[i] = (double*)calloc(n + 1, sizeof(double));
		for (int j = 1; j < n; j++) c->NC[i][j] = c->NC[i][j - 1] + log10(n - j + 1) - log10(j);
	}
	size_t g = (size_t)P * (c->maxploidy + 1);
	c->ic = (int*)malloc(sizeof(int) * P * 24);
	c->qhigh = (double*)malloc(sizeof(double) * P * 16);
	c->stats = (double*)malloc(sizeof(double) * P * 16);
	c->ctab = (int*)malloc(sizeof(int) * P * 6);
	c->ctabw = (double*)malloc(sizeof(double) * P * 4);
	c->ctab2 = (int*)malloc(sizeof(int) * P * 2);
	c->G = (double*)malloc(sizeof(double) * g);
	c->Gf = (double*)malloc(sizeof(double) * g);
	c->Gr = (double*)malloc(sizeof(double) * g);
	c->Gb = (double*)malloc(sizeof(double) * g);
	return c;
}

void crisp_oracle_destroy(void* h)
{
	oracle_ctx* c = (oracle_ctx*)h;
	if (!c) return;
	for (int i = 0; i < c->P; i++) free(c->NC[i]);
	free(c->NC); free(c->ploidy); free(c->phenotypes); free(c->ic); free(c->qhigh); free(c->stats);
	free(c->ctab); free(c->ctabw); free(c->ctab2); free(c->G); free(c->Gf); free(c->Gr); free(c->Gb);
	free(c);
}

const char* crisp_oracle_last_error(void* h) { return ((oracle_ctx*)h)->err; }

/* ------------------------------------------------------------------------------------------------------------
 * contingency tables and asymptotic tests
 * ---------------------------------------------------------------------------------------------------------- */

/* populate_contable — parsebam/allelecounts.c:208-261 (allele3 is never -1 on this path) */
static void o_populate_contable(oracle_ctx* c, int a1, int a2, int refbase, int a3)
{
	for (int i = 0; i < c->P; i++) {
		int* t = c->ctab + i * 6;
		double* w = c->ctabw + i * 4;
		const double* q = c->qhigh + i * 16;
		int c0 = IC(c, i, a1) + IC(c, i, NA + a1) + IC(c, i, 2 * NA + a1) + IC(c, i, a3) + IC(c, i, NA + a3) + IC(c, i, 2 * NA + a3);
		int c1 = IC(c, i, a2) + IC(c, i, NA + a2) + IC(c, i, 2 * NA + a2);
		int alt = (a1 == refbase || a2 != refbase) ? a2 : a1; /* which allele's count goes in the "alt" columns */
		c->ctab2[i * 2] = c0 + c1;
		c->ctab2[i * 2 + 1] = (alt == a2) ? c1 : c0;
		t[0] = IC(c, i, a1) + IC(c, i, a2) + IC(c, i, a3);
		t[2] = IC(c, i, a1 + NA) + IC(c, i, a2 + NA) + IC(c, i, a3 + NA);
		t[4] = IC(c, i, a1 + 2 * NA) + IC(c, i, a2 + 2 * NA);
		t[1] = IC(c, i, alt);
		t[3] = IC(c, i, alt + NA);
		t[5] = IC(c, i, alt + 2 * NA);
		w[0] = q[a1] + q[a2];
		w[0] += q[a3];
		w[2] = q[a1 + NA] + q[a2 + NA];
		w[2] += q[a3 + NA];
		w[1] = q[alt];
		w[3] = q[alt + NA];
	}
}

static double o_chi2_tail(double dof1, double stat, int positive) /* kf_gammaq((DOF-1)/2, stat/2)/ln 10, 0 when !positive */
{
	double r = o_gammaq(dof1 / 2, stat / 2);
	return positive ? r / log(10) : 0;
}

/* chi2pvalue_stratified (+ _pivotsample) — FET/chisquare-stratified.c:12-79, :84-203 */
static void o_chi2_stratified(oracle_ctx* c, int a1, int a2, int a3, double* pval, double chistat[5])
{
	const int P = c->P, pivot = c->cfg.pivot_sample;
	double cnt[2][6] = {{0}};
	double Ef[2], Er[2], Eb[2];
	double sj[2] = {0, 0}, sf[2] = {0, 0}, sr[2] = {0, 0}, snew[2] = {0, 0}, pair[2] = {0, 0};
	for (int i = 0; i < P; i++) {
		int g = (pivot > 0 && i >= pivot) ? 1 : 0;
		int a = IC(c, i, a1) + IC(c, i, a2) + IC(c, i, a3);
		int b = IC(c, i, a1 + NA) + IC(c, i, a2 + NA) + IC(c, i, a3 + NA);
		int cc = IC(c, i, a1 + 2 * NA) + IC(c, i, a2 + 2 * NA) + IC(c, i, a3 + 2 * NA);
		cnt[g][0] += a; cnt[g][1] += IC(c, i, a2);
		cnt[g][2] += b; cnt[g][3] += IC(c, i, a2 + NA);
		cnt[g][4] += cc; cnt[g][5] += IC(c, i, a2 + 2 * NA);
	}
	for (int g = 0; g < 2; g++) {
		Ef[g] = (cnt[g][1] + 0.25) / (cnt[g][0] + 0.25);
		Er[g] = (cnt[g][3] + 0.25) / (cnt[g][2] + 0.25);
		Eb[g] = (cnt[g][5] + 0.25) / (cnt[g][4] + 0.25);
		if (Ef[g] >= 0.999) Ef[g] = 0.999;
		if (Er[g] >= 0.999) Er[g] = 0.999;
		if (Eb[g] >= 0.999) Eb[g] = 0.999;
	}
	double total = cnt[0][0] + cnt[0][2] + cnt[1][0] + cnt[1][2] + cnt[0][4] + cnt[1][4];
	for (int i = 0; i < P; i++) {
		int g = (pivot > 0 && i >= pivot) ? 1 : 0;
		int a = IC(c, i, a1) + IC(c, i, a2) + IC(c, i, a3);
		int b = IC(c, i, a1 + NA) + IC(c, i, a2 + NA) + IC(c, i, a3 + NA);
		int cc = IC(c, i, a1 + 2 * NA) + IC(c, i, a2 + 2 * NA) + IC(c, i, a3 + 2 * NA);
		if (a + b + cc < 1 || total < 25) continue;
		/* :53 sums the squares in int; made explicit (two's-complement wrap, as the compiled reference behaves) */
		double df = 0, dr = 0, db = 0, nf = sqrt((double)(int)((unsigned)a * (unsigned)a + (unsigned)b * (unsigned)b + (unsigned)cc * (unsigned)cc));
		double w1 = a / nf, w2 = b / nf, w3 = cc / nf;
		if (a > 0) df = (IC(c, i, a2) - Ef[g] * a) / sqrt((1.0 - Ef[g]) * Ef[g] * a);
		if (b > 0) dr = (IC(c, i, a2 + NA) - Er[g] * b) / sqrt((1.0 - Er[g]) * Er[g] * b);
		if (cc > 0) db = (IC(c, i, a2 + 2 * NA) - Eb[g] * cc) / sqrt((1.0 - Eb[g]) * Er[g] * cc); /* Er, sic :57 */
		sj[g] += pow(w1 * df + w2 * dr + w3 * db, 2);
		sf[g] += pow(df, 2);
		sr[g] += pow(dr, 2);
		pair[g] += 2 * w1 * w2 * df * dr;
		if ((cnt[1][0] + cnt[0][0]) >= 25 || (cnt[1][2] + cnt[0][2]) >= 25) {
			double delta = w1 * (IC(c, i, a2) - Ef[g] * a) + w2 * (IC(c, i, a2 + NA) - Er[g] * b);
			double tf = w1 * w1 * Ef[g] * (1.0 - Ef[g]) * a + w2 * w2 * Er[g] * (1.0 - Er[g]) * b;
			snew[g] += delta * delta / tf;
		}
	}
	if (pivot == 0) {
		/* :68-75: DOF = P-1, tails at (DOF-1)/2 */
		int DOF = P - 1;
		chistat[0] = sj[0];
		*pval = o_chi2_tail(DOF - 1, sj[0], sj[0] > 0);
		chistat[1] = sf[0];
		chistat[3] = o_chi2_tail(DOF - 1, sf[0], sf[0] > 0);
		chistat[2] = sr[0];
		chistat[4] = o_chi2_tail(DOF - 1, sr[0], sr[0] > 0);
		return;
	}
	/* pivot sample variant :171-200 */
	int DOF = P - 2, DOF1 = pivot - 1, DOF2 = P - pivot - 1;
	double SF = sj[0] + sj[1];
	double pv2 = o_chi2_tail(DOF - 1, SF, SF > 0);
	*pval = pv2;
	chistat[0] = SF; chistat[1] = sf[0] + sf[1]; chistat[2] = sr[0] + sr[1]; chistat[3] = snew[0] + snew[1]; chistat[4] = pair[0] + pair[1];
	double pv0 = o_chi2_tail(DOF1 - 1, sj[0], sj[0] > DOF1 - 1);
	double pv1 = o_chi2_tail(DOF2 - 1, sj[1], sj[1] > DOF2 - 1);
	if (pv0 < pv2 && pv0 < pv1) {
		*pval = pv0;
		chistat[0] = sj[0]; chistat[1] = sf[0]; chistat[2] = sr[0]; chistat[3] = snew[0]; chistat[4] = pair[0];
	} else if (pv1 < pv2) {
		*pval = pv1;
		chistat[0] = sj[1]; chistat[1] = sf[1]; chistat[2] = sr[1]; chistat[3] = snew[1]; chistat[4] = pair[1];
	}
}

/* chi2pvalue on the weighted 2-strand table — FET/chisquare.c:16-136 */
static void o_chi2_weighted(oracle_ctx* c, double* pval, double chistat[5])
{
	const int P = c->P, pivot = c->cfg.pivot_sample;
	double cnt[2][4] = {{0}};
	double Ef[2], Er[2];
	double sj[2] = {0, 0}, sf[2] = {0, 0}, sr[2] = {0, 0}, snew[2] = {0, 0}, pair[2] = {0, 0};
	int DOF = P - 2;
	if (pivot == 0) DOF++;
	int DOF1 = pivot - 1, DOF2 = P - pivot - 1;
	for (int i = 0; i < P; i++) {
		int g = i < pivot ? 0 : 1;
		for (int k = 0; k < 4; k++) cnt[g][k] += c->ctabw[i * 4 + k];
	}
	for (int g = 0; g < 2; g++) {
		Ef[g] = (cnt[g][1] + 0.25) / (cnt[g][0] + 0.25);
		Er[g] = (cnt[g][3] + 0.25) / (cnt[g][2] + 0.25);
		if (Ef[g] >= 0.999) Ef[g] = 0.999;
		if (Er[g] >= 0.999) Er[g] = 0.999;
	}
	for (int i = 0; i < P; i++) {
		int g = i < pivot ? 0 : 1;
		const double* t = c->ctabw + i * 4;
		if (!(t[0] + t[2] >= 1 && cnt[0][0] + cnt[0][2] + cnt[1][0] + cnt[1][2] >= 25)) continue;
		double df = 0, dr = 0;
		double w1 = t[0] / sqrt(t[0] * t[0] + t[2] * t[2]);
		double w2 = sqrt(1.0 - w1 * w1);
		if (t[0] > 0) df = (t[1] - Ef[g] * t[0]) / sqrt((1.0 - Ef[g]) * Ef[g] * t[0]);
		if (t[2] > 0) dr = (t[3] - Er[g] * t[2]) / sqrt((1.0 - Er[g]) * Er[g] * t[2]);
		sj[g] += pow(w1 * df + w2 * dr, 2);
		sf[g] += pow(df, 2);
		sr[g] += pow(dr, 2);
		pair[g] += 2 * w1 * w2 * df * dr;
		if ((cnt[1][0] + cnt[0][0]) >= 25 || (cnt[1][2] + cnt[0][2]) >= 25) {
			double delta = w1 * (t[1] - Ef[g] * t[0]) + w2 * (t[3] - Er[g] * t[2]);
			double tf = w1 * w1 * Ef[g] * (1.0 - Ef[g]) * t[0] + w2 * w2 * Er[g] * (1.0 - Er[g]) * t[2];
			snew[g] += delta * delta / tf;
		}
	}
	double SF = sj[0] + sj[1];
	double pv2 = o_chi2_tail(DOF - 1, SF, SF > 0);
	*pval = pv2;
	chistat[0] = SF; chistat[1] = sf[0] + sf[1]; chistat[2] = sr[0] + sr[1]; chistat[3] = snew[0] + snew[1]; chistat[4] = pair[0] + pair[1];
	if (pivot > 0) {
		double pv0 = o_chi2_tail(DOF1 - 1, sj[0], sj[0] > DOF1 - 1);
		double pv1 = o_chi2_tail(DOF2 - 1, sj[1], sj[1] > DOF2 - 1);
		if (pv0 < pv2 && pv0 < pv1) {
			*pval = pv0;
			chistat[0] = sj[0]; chistat[1] = sf[0]; chistat[2] = sr[0]; chistat[3] = snew[0]; chistat[4] = pair[0];
		} else if (pv1 < pv2) {
			*pval = pv1;
			chistat[0] = sj[1]; chistat[1] = sf[1]; chistat[2] = sr[1]; chistat[3] = snew[1]; chistat[4] = pair[1];
		}
	}
}

/* ------------------------------------------------------------------------------------------------------------
 * Counter-based RNG shared (as a specification) with the CUDA engine: Philox4x32-10.
 * key = (seed_lo ^ pos, seed_hi ^ (tid*8 + slot)); counter = (permutation index, block, stratum, 0).
 * ---------------------------------------------------------------------------------------------------------- */
typedef struct philox_stream {
	uint32_t key[2];
	uint32_t ctr[4];
	uint32_t out[4];
	int have;
} philox_stream;

static void philox4x32_10(const uint32_t ctr[4], const uint32_t key[2], uint32_t out[4])
{
	uint32_t c0 = ctr[0], c1 = ctr[1], c2 = ctr[2], c3 = ctr[3], k0 = key[0], k1 = key[1];
	for (int r = 0; r < 10; r++) {
		uint64_t p0 = (uint64_t)0xD2511F53u * c0, p1 = (uint64_t)0xCD9E8D57u * c2;
		uint32_t n0 = (uint32_t)(p1 >> 32) ^ c1 ^ k0, n1 = (uint32_t)p1, n2 = (uint32_t)(p0 >> 32) ^ c3 ^ k1, n3 = (uint32_t)p0;
		c0 = n0; c1 = n1; c2 = n2; c3 = n3;
		k0 += 0x9E3779B9u;
		k1 += 0xBB67AE85u;
	}
	out[0] = c0; out[1] = c1; out[2] = c2; out[3] = c3;
}

static void ps_init(philox_stream* s, uint64_t seed, int tid, int pos, int slot, uint32_t perm, uint32_t stratum)
{
	s->key[0] = (uint32_t)seed ^ (uint32_t)pos;
	s->key[1] = (uint32_t)(seed >> 32) ^ ((uint32_t)tid * 8u + (uint32_t)slot);
	s->ctr[0] = perm; s->ctr[1] = 0; s->ctr[2] = stratum; s->ctr[3] = 0;
	s->have = 0;
}

static uint32_t ps_next(philox_stream* s)
{
	if (s->have == 0) {
		philox4x32_10(s->ctr, s->key, s->out);
		s->ctr[1]++;
		s->have = 4;
	}
	return s->out[4 - s->have--];
}

/* exact uniform integer in [0, M) (Lemire's multiply-shift with rejection) */
static uint32_t ps_below(philox_stream* s, uint32_t M)
{
	uint64_t m = (uint64_t)ps_next(s) * M;
	uint32_t l = (uint32_t)m;
	if (l < M) {
		uint32_t t = (0u - M) % M;
		while (l < t) {
			m = (uint64_t)ps_next(s) * M;
			l = (uint32_t)m;
		}
	}
	return (uint32_t)(m >> 32);
}

/* Draw `ones` of the M = sum N_i slots without replacement, returning in add[i] how many fell in pool i.
 * A uniform slot is drawn; inside its pool the first used[i] slots are taken to be the occupied ones
 * (slots of a pool are exchangeable), so the draw is rejected when it lands on one of them.  `visit` is called
 * for every accepted ball in draw order. */
typedef void (*ball_fn)(void* st, int pool);
static void sample_stratum(philox_stream* s, const uint32_t* cum, int P, uint32_t M, int ones, int* used, ball_fn visit, void* st)
{
	for (int d = 0; d < ones; d++) {
		for (;;) {
			uint32_t u = ps_below(s, M);
			int lo = 0, hi = P; /* largest i with cum[i] <= u */
			while (hi - lo > 1) {
				int mid = (lo + hi) >> 1;
				if (cum[mid] <= u) lo = mid; else hi = mid;
			}
			if (u - cum[lo] >= (uint32_t)used[lo]) {
				used[lo]++;
				visit(st, lo);
				break;
			}
		}
	}
}

/* ------------------------------------------------------------------------------------------------------------
 * permutation tests
 * ---------------------------------------------------------------------------------------------------------- */
typedef struct strand_state { const int* N; int* r; double stat; const double* L10; } strand_state;
static void strand_ball(void* p, int b)
{
	strand_state* s = (strand_state*)p;
	/* contables.c:316,322 accumulate log10((N_b - c_b)/(c_b + 1)); the engine uses the difference of two table
	 * entries log10(N_b - c_b) - log10(c_b + 1), which this restates */
	s->stat += s->L10[s->N[b] - s->r[b]] - s->L10[s->r[b] + 1];
	s->r[b]++;
}

/* pvalue_contable_iter_stranded — FET/contables.c:261-339 (the permutation part, :296-337).
 * rng_kind DRAND48: the reference's partial Fisher-Yates on a slot->pool index driven by drand48;
 * rng_kind PHILOX : the engine's sampler (same null law, same sequential stopping rule). */
static double o_perm_stranded(oracle_ctx* c, int maxiter, int tid, int pos, int slot, int* out_k, int* out_hits)
{
	const int P = c->P;
	int onesf = 0, readsf = 0, onesr = 0, readsr = 0, hits = 0, k;
	double clra = 0;
	for (int i = 0; i < P; i++) {
		const int* t = c->ctab + i * 6;
		onesf += t[1]; readsf += t[0]; onesr += t[3]; readsr += t[2];
	}
	for (int i = 0; i < P; i++) clra += o_ncr(c->ctab[i * 6] + c->ctab[i * 6 + 2], c->ctab[i * 6 + 1] + c->ctab[i * 6 + 3]);
	int* nalt = (int*)calloc(P, sizeof(int));
	if (c->rng_kind == CRISP_RNG_DRAND48) {
		int* index = (int*)malloc(sizeof(int) * (readsf + readsr + 1));
		int off = 0;
		for (int i = 0; i < P; i++) for (int j = 0; j < c->ctab[i * 6]; j++) index[off++] = i;
		for (int i = 0; i < P; i++) for (int j = 0; j < c->ctab[i * 6 + 2]; j++) index[off++] = i;
		for (k = 1; k <= maxiter; k++) {
			double stat = 0;
			memset(nalt, 0, sizeof(int) * P);
			for (int i = 0; i < onesf; i++) {
				int r = (int)(drand48() * (readsf - i)) + i;
				int b = index[r]; index[r] = index[i]; index[i] = b;
				int Nb = c->ctab[b * 6] + c->ctab[b * 6 + 2];
				stat += log10((double)(Nb - nalt[b]) / (nalt[b] + 1));
				nalt[b]++;
			}
			for (int i = 0; i < onesr; i++) {
				int r = (int)(drand48() * (readsr - i)) + i + readsf;
				int b = index[r]; index[r] = index[i + readsf]; index[i + readsf] = b;
				int Nb = c->ctab[b * 6] + c->ctab[b * 6 + 2];
				stat += log10((double)(Nb - nalt[b]) / (nalt[b] + 1));
				nalt[b]++;
			}
			if (stat <= clra + CT_EPS) hits++;
			if (hits >= 10) break;
			if (k <= 1000 && hits >= 5) break;
		}
		free(index);
	} else {
		uint32_t* cumf = (uint32_t*)malloc(sizeof(uint32_t) * (P + 1));
		uint32_t* cumr = (uint32_t*)malloc(sizeof(uint32_t) * (P + 1));
		int* N = (int*)malloc(sizeof(int) * P);
		int* usedf = (int*)malloc(sizeof(int) * P);
		int* usedr = (int*)malloc(sizeof(int) * P);
		int maxN = 0;
		cumf[0] = cumr[0] = 0;
		for (int i = 0; i < P; i++) {
			cumf[i + 1] = cumf[i] + c->ctab[i * 6];
			cumr[i + 1] = cumr[i] + c->ctab[i * 6 + 2];
			N[i] = c->ctab[i * 6] + c->ctab[i * 6 + 2];
			if (N[i] > maxN) maxN = N[i];
		}
		double* L10 = (double*)malloc(sizeof(double) * (maxN + 2));
		L10[0] = 0;
		for (int x = 1; x <= maxN + 1; x++) L10[x] = log10((double)x);
		for (k = 1; k <= maxiter; k++) {
			philox_stream ps;
			strand_state st = {N, nalt, 0.0, L10};
			memset(nalt, 0, sizeof(int) * P);
			memset(usedf, 0, sizeof(int) * P);
			memset(usedr, 0, sizeof(int) * P);
			ps_init(&ps, c->cfg.rng_seed, tid, pos, slot, (uint32_t)k, 0);
			sample_stratum(&ps, cumf, P, (uint32_t)readsf, onesf, usedf, strand_ball, &st);
			ps_init(&ps, c->cfg.rng_seed, tid, pos, slot, (uint32_t)k, 1);
			sample_stratum(&ps, cumr, P, (uint32_t)readsr, onesr, usedr, strand_ball, &st);
			if (st.stat <= clra + CT_EPS) hits++;
			if (hits >= 10) break;
			if (k <= 1000 && hits >= 5) break;
		}
		free(cumf); free(cumr); free(N); free(usedf); free(usedr); free(L10);
	}
	free(nalt);
	*out_k = k;
	*out_hits = hits;
	return log10(hits + 1) - log10(k + 1);
}

typedef struct low_state { int* r; } low_state;
static void low_ball(void* p, int b) { ((low_state*)p)->r[b]++; }
static int cmp_int(const void* a, const void* b) { return abs(*(const int*)a) - abs(*(const int*)b); }

/* pvalue_lowcoverage — FET/lowcovFET.c:23-149 */
static double o_perm_lowcov(oracle_ctx* c, int iter, int tid, int pos, int slot, int* out_k, int* out_hits, int* ran)
{
	const int P = c->P;
	int ones[3] = {0, 0, 0}, reads[3] = {0, 0, 0}, k;
	unsigned lo = 0, hi = 0;
	double clra = 0;
	int* N = (int*)malloc(sizeof(int) * P);
	for (int i = 0; i < P; i++) {
		const int* t = c->ctab + i * 6;
		N[i] = t[0] + t[2] + t[4];
		int Ri = t[1] + t[3] + t[5];
		for (int s = 0; s < 3; s++) { ones[s] += t[2 * s + 1]; reads[s] += t[2 * s]; }
		if (Ri > 0) clra += o_ncr(N[i], Ri);
	}
	*ran = 0;
	*out_k = 0;
	*out_hits = 0;
	if (ones[0] + ones[1] + ones[2] < 2) { free(N); return 0; }
	*ran = 1;
	int tot = ones[0] + ones[1] + ones[2];
	int* cnt = (int*)calloc(P, sizeof(int));
	if (c->rng_kind == CRISP_RNG_DRAND48) {
		int* index = (int*)malloc(sizeof(int) * (reads[0] + reads[1] + reads[2] + 1));
		int* ball = (int*)malloc(sizeof(int) * (tot + 1));
		int off = 0;
		for (int s = 0; s < 3; s++) for (int i = 0; i < P; i++) for (int j = 0; j < c->ctab[i * 6 + 2 * s]; j++) index[off++] = i;
		for (k = 1; k <= iter; k++) {
			int base = 0, nb = 0;
			for (int s = 0; s < 3; s++) {
				for (int i = 0; i < ones[s]; i++) {
					int r = (int)(drand48() * (reads[s] - i)) + i + base;
					int b = index[r]; index[r] = index[i + base]; index[i + base] = b;
					ball[nb++] = b;
				}
				base += reads[s];
			}
			qsort(ball, tot, sizeof(int), cmp_int);
			double stat = 0;
			for (int i = 0; i < tot;) { /* runs of equal pool index, ascending */
				int j = i;
				while (j < tot && ball[j] == ball[i]) j++;
				stat += o_ncr(N[ball[i]], j - i);
				i = j;
			}
			if (stat <= clra) lo++;
			if (stat >= clra) hi++;
			if ((lo >= 10 && hi >= 10) || (k <= 100 && lo >= 5 && hi >= 5)) break;
		}
		free(index); free(ball);
	} else {
		uint32_t* cum = (uint32_t*)malloc(sizeof(uint32_t) * (P + 1));
		int* used = (int*)malloc(sizeof(int) * P);
		for (k = 1; k <= iter; k++) {
			memset(cnt, 0, sizeof(int) * P);
			low_state st = {cnt};
			for (int s = 0; s < 3; s++) {
				philox_stream ps;
				cum[0] = 0;
				for (int i = 0; i < P; i++) cum[i + 1] = cum[i] + c->ctab[i * 6 + 2 * s];
				memset(used, 0, sizeof(int) * P);
				ps_init(&ps, c->cfg.rng_seed, tid, pos, slot, (uint32_t)k, (uint32_t)s);
				sample_stratum(&ps, cum, P, (uint32_t)reads[s], ones[s], used, low_ball, &st);
			}
			double stat = 0;
			for (int i = 0; i < P; i++) if (cnt[i] > 0) stat += o_ncr(N[i], cnt[i]);
			if (stat <= clra) lo++;
			if (stat >= clra) hi++;
			if ((lo >= 10 && hi >= 10) || (k <= 100 && lo >= 5 && hi >= 5)) break;
		}
		free(cum); free(used);
	}
	free(cnt); free(N);
	*out_k = k;
	if (hi < lo) { *out_hits = (int)hi; return log10((double)hi + 1) - log10(k + 1); }
	*out_hits = (int)lo;
	return log10((double)lo + 1) - log10(k + 1);
}

/* ------------------------------------------------------------------------------------------------------------
 * exact 2 x k test for small tables — FET/contables_exact.c:62-117 (not compiled into CRISP.binary; the specification
 * of the engine's exact_mode)
 * ---------------------------------------------------------------------------------------------------------- */
#define EXACT_MAXN 2001 /* MAXN, FET/contables.c:10 */

/* NCRtable[n][j], j < n: running sum of log10(n-j+1) - log10(j) (compute_NCRtable, contables.c:244-257); 0 for j >= n */
static void o_ncr_row(int n, int upto, double* row)
{
	double value = 0;
	row[0] = 0;
	for (int j = 1; j <= upto; j++) {
		if (j < n) { value += log10(n - j + 1) - log10(j); row[j] = value; }
		else row[j] = 0;
	}
}

static double o_exact_rec(const int* N, double* const* rows, int size, int offset, int R, int A, double prob)
{
	if (A > R) return -1;
	if (size == 1) {
		double cp = N[offset] > 0 ? rows[offset][A] : 0;
		return cp > prob ? -1 : cp;
	}
	double sum = 0;
	for (int i = 0; i <= N[offset] && i <= A; i++) {
		double cp = N[offset] > 0 ? rows[offset][i] : 0;
		double a = o_exact_rec(N, rows, size - 1, offset + 1, R - N[offset], A - i, prob - cp);
		if (a > -1) sum += pow(10, a + cp - prob);
	}
	if (sum > 0) return log10(sum) + prob;
	return -1;
}

/* gate of pvalue_contable_iter (contables_exact.c:104); returns 1 and the log10 p-value when the table qualifies */
static int o_exact_2xk(const int* N, const int* A, int size, double* log10p)
{
	int ones = 0, reads = 0, maxcol = 0;
	double cl = 0;
	for (int i = 0; i < size; i++) {
		cl += o_ncr(N[i], A[i]);
		ones += A[i];
		reads += N[i];
		if (N[i] > maxcol) maxcol = N[i];
	}
	if (!(ones < 70 && maxcol < EXACT_MAXN && size >= 2 && size < 8)) return 0;
	double* rows[8];
	for (int i = 0; i < size; i++) {
		rows[i] = (double*)calloc(ones + 2, sizeof(double));
		o_ncr_row(N[i], ones, rows[i]);
	}
	double p = o_exact_rec(N, rows, size, 0, reads, ones, cl);
	for (int i = 0; i < size; i++) free(rows[i]);
	*log10p = p - o_ncr(reads, ones);
	return 1;
}

/* ------------------------------------------------------------------------------------------------------------
 * evaluate_variant — crisp/newcrispcaller.c:34-132
 * ---------------------------------------------------------------------------------------------------------- */
typedef struct slot_out {
	int status; double paf, ctpval, chi2pval; int perm_k, perm_hits;
	double af_ml, delta, deltaf, hwe, qvpf, qvpr; int varpools, allelecount, variantpools, em_iters;
	double assoc[6]; /* p, LLR, OR, p1, LLR1, OR1 */
} slot_out;

static int o_evaluate(oracle_ctx* c, const crispgpu_batch* b, int s, int slot, int a1, int a2, int a3, double paf[2], slot_out* o)
{
	const int P = c->P;
	const crispgpu_config* g = &c->cfg;
	double p[2] = {0.0, 0.0}, pv1 = 0, pv2 = 0, chistat[6] = {0};
	int K = 0;
	for (int i = 0; i < P; i++) {
		int n = c->ploidy[i];
		K += n;
		double C1 = 0, C2 = 0, C3 = 0;
		for (int k = 0; k < 3; k++) { C1 += IC(c, i, a1 + k * NA); C2 += IC(c, i, a2 + k * NA); C3 += IC(c, i, a3 + k * NA); }
		double p1 = C2 / (C1 + C2 + C3 + 0.001);
		p1 *= n;
		if (p1 >= 0.2 && C2 >= 4) p[0] += p1;
		else if (p1 >= 0.4 && C2 >= 3) p[0] += p1;
		else if (n == 2 && p1 >= 0.25 && C2 >= 1) p[0] += p1;
		if (C3 >= 2) {
			double p2 = C3 / (C1 + C2 + C3 + 0.001);
			if (p2 >= 0.25 && C3 >= 4) p[1] += p2;
			else if (p2 >= 0.4 && C3 >= 3) p[1] += p2;
			else if (n == 2 && p2 >= 0.25 && C3 >= 1) p[1] += p2;
		}
	}
	o->paf = p[0];
	if (p[0] < 0.25 || (c->ploidy[0] == 2 && p[0] <= 0.4)) { o->status = CRISPGPU_SLOT_REJECT_AF; return 0; }
	o_populate_contable(c, a1, a2, b->refbase[s], a3);
	if (c->ploidy[0] > 2 && g->chisq_permutation == 1) {
		o_chi2_weighted(c, &pv2, chistat);
		int exact = 0;
		if (g->exact_mode == 1 && P < 8) {
			/* engine extension: small tables get the exact test on the unstranded table (N, alt) instead of the Monte-Carlo estimate */
			int Nn[8], Aa[8];
			for (int i = 0; i < P; i++) { Nn[i] = c->ctab2[i * 2]; Aa[i] = c->ctab2[i * 2 + 1]; }
			exact = o_exact_2xk(Nn, Aa, P, &pv1);
			if (exact) { o->perm_k = 0; o->perm_hits = -1; }
		}
		if (!exact) pv1 = o_perm_stranded(c, g->max_iter, b->tid[s], b->pos[s], slot, &o->perm_k, &o->perm_hits);
		o->ctpval = pv1;
		o->chi2pval = pv2;
	} else if (c->ploidy[0] == 2 && g->chisq_permutation == 1) {
		int ran;
		pv1 = o_perm_lowcov(c, g->max_iter, b->tid[s], b->pos[s], slot, &o->perm_k, &o->perm_hits, &ran);
		o->ctpval = pv1;
		o->chi2pval = 0;
	} else {
		o_chi2_stratified(c, a1, a2, a3, &pv1, chistat);
		o->ctpval = pv1;
		o->chi2pval = pv1;
	}
	o->status = CRISPGPU_SLOT_REJECT_CT;
	if (a2 >= 4) {
		int ilen = b->indel_len ? abs(b->indel_len[s * 8 + a2]) : 1;
		if (ilen < 1) ilen = 1;
		int hpl = (b->hplen ? b->hplen[s * 8 + a2] : 0) / ilen;
		if (g->fast_filter == 1 && o->ctpval >= -3 && c->ploidy[0] > 2 && hpl >= 4 && (p[0] / K) <= 0.05) return 0;
		else if (((g->fast_filter == 1) & (o->ctpval >= -3)) && c->ploidy[0] > 2 && hpl >= 8 && (p[0] / K) <= 0.1) return 0;
	}
	if (g->fast_filter == 1) {
		if (o->ctpval >= -3 && c->ploidy[0] > 2 && p[0] <= P / 4) return 0;
		else if (o->ctpval >= -2 && c->ploidy[0] > 2 && p[0] <= P / 2) return 0;
		else if (pv1 >= -2 && pv2 >= -2 && c->ploidy[0] == 2 && p[0] <= 5 && p[0] <= P / 4) return 0;
	}
	paf[0] = p[0];
	paf[1] = p[1];
	o->status = CRISPGPU_SLOT_NOCALL;
	return 1;
}

/* ------------------------------------------------------------------------------------------------------------
 * likelihoods — crisp/crispEM.c:36-138
 * ---------------------------------------------------------------------------------------------------------- */
static void o_likelihood_snp(oracle_ctx* c, const crispgpu_batch* b, int s, int a1, int a2)
{
	const int P = c->P;
	const double beta = c->cfg.agilent_ref_bias;
	for (int i = 0; i < P; i++)
		for (int j = 0; j <= c->ploidy[i]; j++) GX(c, c->G, i, j) = GX(c, c->Gf, i, j) = GX(c, c->Gr, i, j) = GX(c, c->Gb, i, j) = 0.0;
	for (int i = 0; i < P; i++) {
		const int n = c->ploidy[i];
		for (uint32_t r = b->read_off[(size_t)s * P + i]; r < b->read_off[(size_t)s * P + i + 1]; r++) {
			crispgpu_read w = b->reads[r];
			int base = (w >> 8) & 3;
			if (base != a1 && base != a2) continue;
			double ep = pow(0.1, ((double)(signed char)(w & 0xff) - c->cfg.qv_offset) / 10);
			double* S = (w >> 11) & 1 ? c->Gb : ((w >> 10) & 1 ? c->Gr : c->Gf);
			for (int j = 0; j <= n; j++) {
				double f = (double)j * (1.0 - beta);
				f /= f + (double)(n - j) * beta;
				double e = (1.0 - ep) * (1.0 - f) + ep * f;
				double le = base == a1 ? log10(e) : log10(1.0 - e);
				GX(c, c->G, i, j) += le;
				GX(c, S, i, j) += le;
			}
		}
	}
}

static double o_indel_bias(int ilen, int is_ins) /* crispEM.c:36-49, file-local READLENGTH=100, MINFLANK=5 */
{
	double het = 0, RL = 100, MF = 5;
	if (is_ins && ilen >= 1) het += (MF + ilen / 2) / (double)(RL - MF - ilen / 2);
	else if (ilen >= 1 && !is_ins) het += (MF) / (double)(RL - MF);
	return het;
}

static void o_likelihood_indel(oracle_ctx* c, int a1, int a2, int ilen, int is_ins, double E01[3], double E10[3])
{
	for (int i = 0; i < 3; i++) {
		if (E01[i] < MINP) E01[i] = MINP; else if (E01[i] > 1.0 - MINP) E01[i] = 1.0 - MINP;
		if (E10[i] < MINP) E10[i] = MINP; else if (E10[i] > 1.0 - MINP) E10[i] = 1.0 - MINP;
	}
	E01[2] = E01[0] * E01[1];
	E10[2] = E10[0] * E10[1];
	double hb = o_indel_bias(ilen, is_ins);
	for (int i = 0; i < c->P; i++) {
		int n = c->ploidy[i];
		for (int j = 0; j <= n; j++) {
			double f = (double)j / n;
			f /= (1 + hb);
			double ef = (1.0 - E01[0]) * (1.0 - f) + E10[0] * f, er = (1.0 - E01[1]) * (1.0 - f) + E10[1] * f, eb = (1.0 - E01[2]) * (1.0 - f) + E10[2] * f;
			GX(c, c->Gf, i, j) = IC(c, i, a1) * log10(ef) + IC(c, i, a2) * log10(1.0 - ef);
			GX(c, c->Gr, i, j) = IC(c, i, a1 + NA) * log10(er) + IC(c, i, a2 + NA) * log10(1.0 - er);
			GX(c, c->Gb, i, j) = IC(c, i, a1 + 2 * NA) * log10(eb) + IC(c, i, a2 + 2 * NA) * log10(1.0 - eb);
			GX(c, c->G, i, j) = GX(c, c->Gf, i, j) + GX(c, c->Gr, i, j) + GX(c, c->Gb, i, j);
		}
	}
}

/* ------------------------------------------------------------------------------------------------------------
 * calculate_genotypes — crisp/crispEM.c:141-203
 * ---------------------------------------------------------------------------------------------------------- */
static void o_genotypes(oracle_ctx* c, double p, int* AC, double* QV, int* variantpools, int* allelecount, double* hwe)
{
	const int P = c->P;
	double* gc = (double*)calloc(c->maxploidy + 1, sizeof(double));
	const double lp = log10(p), lq = log10(1.0 - p);
	*variantpools = 0;
	*allelecount = 0;
	for (int i = 0; i < P; i++) {
		const int n = c->ploidy[i];
		double Lb = -100000, best = -100000, second = -100000;
		int bg = 0, sg = 0;
		for (int j = 0; j <= n; j++) {
			double t = GX(c, c->G, i, j) + c->NC[i][j] + lp * j + lq * (n - j);
			if (t > best) { sg = bg; second = best; bg = j; best = t; }
			else if (t > second) { sg = j; second = t; }
			if (t > Lb || j == 0) Lb = t + log10(1.0 + pow(10, Lb - t));
			else Lb += log10(1.0 + pow(10, t - Lb));
		}
		(void)sg;
		for (int j = 0; j <= n; j++) {
			double t = GX(c, c->G, i, j) + c->NC[i][j] + lp * j + lq * (n - j);
			gc[j] += pow(10, t - Lb);
		}
		AC[i] = bg;
		QV[i] = 10 * (best - second);
		if (bg > 0) (*variantpools)++;
		*allelecount += bg;
	}
	if (c->cfg.varpoolsize == 0) {
		double stat = 0, dof = 0;
		const int n = c->ploidy[0];
		for (int j = 0; j <= n; j++) {
			double ec = pow(10, c->NC[0][j] + j * log10(p) + (n - j) * log10(1.0 - p) + log10(P));
			if (ec >= 2) { stat += pow(gc[j] - ec, 2) / ec; dof += 1; }
		}
		if (stat > 0 && dof >= 2) *hwe = o_gammaq((double)(dof - 1) / 2, (double)stat / 2) / log(10);
		else *hwe = 0;
	}
	free(gc);
}

/* ------------------------------------------------------------------------------------------------------------
 * EMmethod — crisp/crispEM.c:207-351 (outputs only: the `b` class and LLnull* feed nothing)
 * ---------------------------------------------------------------------------------------------------------- */
#define LSE_STEP(L, t, j) do { if ((j) == 0) L = (t); else if ((t) > L) L = (t) + log10(1.0 + pow(10, L - (t))); else L += log10(1.0 + pow(10, (t) - L)); } while (0)

static double o_em(oracle_ctx* c, int a1, int a2, int ilen, int is_ins, double p, slot_out* o, int* AC, double* QV)
{
	const int P = c->P, K = c->K;
	const int recompute = (a2 >= 4 || c->cfg.use_base_qvs == 0);
	double E01[3] = {0.001, 0.001, 0.001}, E10[3] = {0.001, 0.001, 0.001};
	double pf = p, pr = p, LL = 0, LLf = 0, LLr = 0, LLnull0 = 0, hwe = 0;
	double ALPHA = 0.0;
	for (int j = 1; j <= K; j++) ALPHA += 1.0 / j;
	ALPHA *= c->cfg.theta;
	int iter = 0, exitloop = 0;
	const double alpha = 1.5, beta = 50;
	while (iter++ < 1000 && exitloop == 0) {
		double pLL = LL, pLLf = LLf, pLLr = LLr;
		double pnew = 0, pnewf = 0, pnewr = 0;
		double T00[3] = {0, 0, 0}, T01[3] = {0, 0, 0}, T10[3] = {0, 0, 0}, T11[3] = {0, 0, 0};
		LL = LLf = LLr = 0;
		LLnull0 = 0.0;
		const double lp0 = log10(p), lp1 = log10(1.0 - p), lf0 = log10(pf), lf1 = log10(1.0 - pf), lr0 = log10(pr), lr1 = log10(1.0 - pr);
		for (int i = 0; i < P; i++) {
			const int n = c->ploidy[i];
			double Lj = -100000, Lf = -100000, Lr = -100000;
			LLnull0 += GX(c, c->G, i, 0);
			for (int j = 0; j <= n; j++) {
				double t = GX(c, c->G, i, j) + c->NC[i][j] + lp0 * j + lp1 * (n - j);
				double tf = GX(c, c->Gf, i, j) + c->NC[i][j] + lf0 * j + lf1 * (n - j);
				double tr = GX(c, c->Gr, i, j) + c->NC[i][j] + lr0 * j + lr1 * (n - j);
				LSE_STEP(Lj, t, j);
				LSE_STEP(Lf, tf, j);
				LSE_STEP(Lr, tr, j);
			}
			LL += Lj; LLf += Lf; LLr += Lr;
			LLr += GX(c, c->Gf, i, 0) + GX(c, c->Gb, i, 0);
			LLf += GX(c, c->Gr, i, 0) + GX(c, c->Gb, i, 0);
			for (int j = 0; j <= n; j++) {
				double f = (double)j / n;
				double t = GX(c, c->G, i, j) + c->NC[i][j] + lp0 * j + lp1 * (n - j);
				double tf = GX(c, c->Gf, i, j) + c->NC[i][j] + lf0 * j + lf1 * (n - j);
				double tr = GX(c, c->Gr, i, j) + c->NC[i][j] + lr0 * j + lr1 * (n - j);
				pnew += (double)j * pow(10, t - Lj);
				pnewf += (double)j * pow(10, tf - Lf);
				pnewr += (double)j * pow(10, tr - Lr);
				if (recompute) {
					for (int k = 0; k < 3; k++) {
						double w = pow(10, t - Lj);
						T00[k] += (double)IC(c, i, a1 + k * NA) * w * (1.0 - f) * (1.0 - E01[k]) / ((1.0 - f) * (1.0 - E01[k]) + f * E10[k]);
						T10[k] += (double)IC(c, i, a1 + k * NA) * w * f * E10[k] / ((1.0 - f) * (1.0 - E01[k]) + f * E10[k]);
						T11[k] += (double)IC(c, i, a2 + k * NA) * w * f * (1.0 - E10[k]) / ((1.0 - f) * E01[k] + f * (1.0 - E10[k]));
						T01[k] += (double)IC(c, i, a2 + k * NA) * w * (1.0 - f) * E01[k] / ((1.0 - f) * E01[k] + f * (1.0 - E10[k]));
					}
				}
			}
		}
		pnew /= K; pnewf /= K; pnewr /= K;
		if (recompute) {
			for (int k = 0; k < 3; k++) {
				E01[k] = (T01[k] + alpha - 1) / (T01[k] + T00[k] + alpha + beta - 1);
				E10[k] = (T10[k] + alpha - 1) / (T10[k] + T11[k] + beta + alpha - 1);
				if (E10[k] > 0.05) E10[k] = 0.05;
			}
			o_likelihood_indel(c, a1, a2, ilen, is_ins, E01, E10);
		}
		if (pnew < MINP) pnew = MINP;
		if (1.0 - pnew <= MINP) pnew = 1.0 - MINP;
		if (pnewf < MINP) pnewf = MINP;
		if (1.0 - pnewf <= MINP) pnewf = 1.0 - MINP;
		if (pnewr < MINP) pnewr = MINP;
		if (1.0 - pnewr <= MINP) pnewr = 1.0 - MINP;
		/* :304 — fabsf() of a double difference: the argument is rounded to float first, the thresholds are doubles */
		if (iter >= 2 && fabsf(LL - pLL) <= 0.0001 && fabsf(LLf - pLLf) <= 0.01 && fabsf(LLr - pLLr) <= 0.01) exitloop = 1;
		p = pnew; pf = pnewf; pr = pnewr;
	}
	o->em_iters = iter - 1;
	double deltaf = (LLf > LLr) ? LLf - LL - log10(2.0) : LLr - LL - log10(2.0);
	double LLrefadd = LLnull0 + log10(1.0 - ALPHA);
	LL += log10(ALPHA);
	if (LLrefadd < LL) LL += log10(1.0 + pow(10, LLrefadd - LL));
	else LL = LLrefadd + log10(1.0 + pow(10, LL - LLrefadd));
	double delta = LL - LLrefadd;
	if (delta >= 2) o_genotypes(c, p, AC, QV, &o->variantpools, &o->allelecount, &hwe);
	o->af_ml = p;
	o->delta = delta;
	o->deltaf = deltaf;
	o->hwe = hwe;
	return p;
}

/* calculate_likelihood_EM — crisp/LRstatistic.c:6-45: the EM restricted to the pools whose phenotype matches `flag` */
static double o_subset_em(oracle_ctx* c, int a1, int a2, int ilen, int is_ins, double p, double* dataLL, double E01[3], double E10[3], int flag)
{
	const int P = c->P;
	int K = 0, iter = 0, exitloop = 0;
	double LL = 0, LLprev = 0;
	for (int i = 0; i < P; i++) if (c->phenotypes[i] & flag) K += c->ploidy[i];
	while (iter++ < 1000 && exitloop == 0) {
		double pnew = 0;
		LL = 0;
		const double lp0 = log10(p), lp1 = log10(1.0 - p);
		for (int i = 0; i < P; i++) {
			if ((c->phenotypes[i] & flag) == 0) continue;
			const int n = c->ploidy[i];
			double Lj = 0;
			for (int j = 0; j <= n; j++) {
				double t = GX(c, c->G, i, j) + c->NC[i][j] + lp0 * j + lp1 * (n - j);
				LSE_STEP(Lj, t, j);
			}
			LL += Lj;
			for (int j = 0; j <= n; j++) {
				double t = GX(c, c->G, i, j) + c->NC[i][j] + lp0 * j + lp1 * (n - j);
				pnew += (double)j * pow(10, t - Lj);
			}
		}
		pnew /= K;
		if (a2 >= 4 || c->cfg.use_base_qvs == 0) o_likelihood_indel(c, a1, a2, ilen, is_ins, E01, E10);
		if (pnew < MINP) pnew = MINP;
		if (1.0 - pnew <= MINP) pnew = 1.0 - MINP;
		if (iter >= 2 && fabsf(LL - LLprev) <= 0.01) exitloop = 1;
		p = pnew;
		LLprev = LL;
	}
	*dataLL = LL;
	return p;
}

/* calculate_LRstatistic — crisp/LRstatistic.c:48-69; phenotype bits: 1 control, 2 case, 4 early onset */
static void o_lr_statistic(oracle_ctx* c, int a1, int a2, int ilen, int is_ins, double p, double E01[3], double E10[3], double out[6])
{
	double LLall, LLctl, LLcase, LLearly, LL1;
	o_subset_em(c, a1, a2, ilen, is_ins, p, &LLall, E01, E10, 7);
	double pcontrols = o_subset_em(c, a1, a2, ilen, is_ins, p, &LLctl, E01, E10, 1);
	double pcases = o_subset_em(c, a1, a2, ilen, is_ins, p, &LLcase, E01, E10, 6);
	double OR = pcases * (1.0 - pcontrols) / ((1.0 - pcases) * pcontrols);
	double LLR = LLcase + LLctl - LLall;
	double pvalue = o_gammaq(0.5, LLR * log(10)) / log(10);
	double pcases1 = o_subset_em(c, a1, a2, ilen, is_ins, p, &LLearly, E01, E10, 4);
	o_subset_em(c, a1, a2, ilen, is_ins, p, &LL1, E01, E10, 5);
	double OR1 = pcases1 * (1.0 - pcontrols) / ((1.0 - pcases1) * pcontrols);
	double LLR1 = LLearly + LLctl - LL1;
	double pvalue1 = o_gammaq(0.5, LLR1 * log(10)) / log(10);
	out[0] = pvalue; out[1] = LLR; out[2] = OR; out[3] = pvalue1; out[4] = LLR1; out[5] = OR1;
}

/* compute_GLL_pooled — crisp/crispEM.c:356-391 */
static int o_compute_gll(oracle_ctx* c, const crispgpu_batch* b, int s, int a1, int a2, double paf[2], slot_out* o, int* AC, double* QV)
{
	const int K = c->K;
	paf[0] /= K;
	if (1.0 - paf[0] < MINP) paf[0] = 1.0 - MINP;
	paf[1] /= K;
	if (1.0 - paf[1] < MINP) paf[1] = 1.0 - MINP;
	int ilen = 0, is_ins = 0;
	double E01[3] = {0.001, 0.001, 0.001}, E10[3] = {0.001, 0.001, 0.001};
	if (a2 >= 4 || c->cfg.use_base_qvs == 0) {
		if (a2 >= 4 && b->indel_len) { ilen = abs(b->indel_len[s * 8 + a2]); is_ins = b->indel_len[s * 8 + a2] > 0; }
		for (int k = 0; k < 3; k++) E01[k] = (double)(c->counts[a2 + k * NA] + 1.0) / (c->counts[a1 + k * NA] + c->counts[a2 + k * NA] + 1.0);
		if (E01[0] >= 0.005 || E01[1] >= 0.005) E01[0] = E01[1] = E01[2] = 0.005;
		for (int k = 0; k < 3; k++) E10[k] = E01[k];
		o_likelihood_indel(c, a1, a2, ilen, is_ins, E01, E10);
	} else o_likelihood_snp(c, b, s, a1, a2);
	double pem = o_em(c, a1, a2, ilen, is_ins, paf[0], o, AC, QV);
	if (c->cfg.association == 1 && pem >= 0.0001 && o->delta >= 3) o_lr_statistic(c, a1, a2, ilen, is_ins, pem, E01, E10, o->assoc);  /* crispEM.c:383 */
	paf[0] = pem;
	return o->delta >= 2 ? 1 : 0;
}

/* ------------------------------------------------------------------------------------------------------------
 * calculate_uniquereads + calculate_pool_statistics — parsebam/allelecounts.c:42-81, crisp/poolstats.c:5-60
 * ---------------------------------------------------------------------------------------------------------- */
static int cmp_alt(const void* x, const void* y)
{
	const crispgpu_altread* a = (const crispgpu_altread*)x; const crispgpu_altread* b = (const crispgpu_altread*)y;
	if (a->strand == b->strand) return (int)a->offset - (int)b->offset;
	return (int)a->strand - (int)b->strand;
}

static void o_uniquereads(const crispgpu_batch* b, int s, int slot, int pool, int cap, int* unique, int* middle)
{
	*unique = 1;
	*middle = 0;
	if (!b->alt_off) return;
	uint32_t lo = b->alt_off[s * 5 + slot], hi = b->alt_off[s * 5 + slot + 1];
	while (lo < hi && b->alt_reads[lo].pool < pool) lo++;
	uint32_t e = lo;
	while (e < hi && b->alt_reads[e].pool == pool) e++;
	int m = (int)(e - lo);
	if (m > cap) m = cap;
	if (m <= 0) return;
	crispgpu_altread* v = (crispgpu_altread*)malloc(sizeof(crispgpu_altread) * m);
	memcpy(v, b->alt_reads + lo, sizeof(crispgpu_altread) * m);
	if (m > 2) qsort(v, m, sizeof(crispgpu_altread), cmp_alt);
	for (int i = 0; i < m - 1; i++)
		if (v[i].strand != v[i + 1].strand || v[i].offset != v[i + 1].offset) (*unique)++;
	for (int i = 0; i < m; i++)
		if (v[i].offset >= 5 && (int)v[i].aligned - (int)v[i].offset >= 5) (*middle)++;
	free(v);
}

static int o_pool_statistics(oracle_ctx* c, const crispgpu_batch* b, int s, int slot, int a2)
{
	const int MR = c->cfg.min_reads;
	int variantpools = 0;
	for (int a = 0; a < c->P; a++) {
		int f = IC(c, a, a2), r = IC(c, a, a2 + NA), bd = IC(c, a, a2 + 2 * NA);
		if (f + r + bd < 2) continue;
		int uq, mid;
		o_uniquereads(b, s, slot, a, f + r + bd, &uq, &mid);
		int varpool = 0;
		if (f + r + bd >= MR && mid > 0 && uq >= MR - 1) varpool = 1;
		else if (f + r + bd >= MR - 1 && f + bd > 0 && r + bd > 0 && mid > 0 && uq >= MR - 1) varpool = 1;
		variantpools += varpool;
	}
	return variantpools;
}

/* ------------------------------------------------------------------------------------------------------------
 * --EM 0: pooled_variantcaller + call_genotypes — crisp/oldcrispmethod.c:131-204, 4-110
 * ---------------------------------------------------------------------------------------------------------- */
static int o_legacy_pools(oracle_ctx* c, const crispgpu_batch* b, int s, int slot, int a1, int a2, double ctpval, int* strandpass)
{
	const crispgpu_config* g = &c->cfg;
	const int P = c->P;
	int variants = 0, spools = 0;
	double* pvstrand = (double*)calloc(P, sizeof(double));
	*strandpass = 0;
	for (int a = 0; a < P; a++) {
		const int* t = c->ctab + a * 6;
		const int n = c->ploidy[a];
		int altf = IC(c, a, a2), altr = IC(c, a, a2 + NA), altb = IC(c, a, a2 + 2 * NA);
		if (altf + altr + altb < 2) continue;
		int uq, mid;
		o_uniquereads(b, s, slot, a, altf + altr + altb, &uq, &mid);
		double pvallow = o_minreads_pvalue(c->ctab2[a * 2], c->ctab2[a * 2 + 1], 1.0 / n);
		int C1 = altf, N1 = C1 + IC(c, a, a1), C2 = altr, N2 = C2 + IC(c, a, a1 + NA);
		pvstrand[a] = o_fet(C1, N1, C2, N2);
		double minpvlog = o_minreads_pvalue(t[0], t[1], 1.0 / n) + o_minreads_pvalue(t[2], t[3], 1.0 / n);
		double minpvlog1 = o_minreads_pvalue(t[0], t[1], g->relax / n) + o_minreads_pvalue(t[2], t[3], g->relax / n);
		if (n > 2) {
			int pass = 0;
			if (minpvlog >= g->thresh3 || (minpvlog1 >= g->thresh3 && ctpval - minpvlog <= g->thresh1 && n >= 10) || (a2 >= 4 && minpvlog1 >= g->thresh3)) pass++;
			else if (minpvlog < g->thresh3 && (pvallow + pvstrand[a] >= -3)) pass++;
			if ((t[1] > 0 && t[3] > 0 && uq >= 3) || uq >= 4) pass++;
			if (c->ctab2[a * 2 + 1] >= g->min_reads) pass++;
			if (mid > 0) pass++;
			if (pass == 4) variants += 1;
		} else {
			if (minpvlog >= g->thresh3 && t[1] > 0 && t[3] > 0 && uq >= 2) variants += 1;
			else if (minpvlog >= g->thresh3 && c->ctab2[a * 2 + 1] >= g->min_reads && uq >= 2) variants += 1;
			else if (minpvlog1 >= g->thresh3 && a2 >= 4 && c->ctab2[a * 2 + 1] >= g->min_reads && uq >= 2) variants += 1;
			else if (minpvlog1 >= g->thresh3 && a2 < 4 && c->ctab2[a * 2 + 1] >= 5 && uq >= 3) variants += 1;
		}
		if (minpvlog1 >= g->thresh3) spools++;
		else pvstrand[a] = 0;
	}
	for (int a = 0; a < P; a++)
		if (pvstrand[a] < (log10(0.01) - log10(1.0 + spools))) (*strandpass)++;
	free(pvstrand);
	return variants;
}

static int o_legacy_caller(oracle_ctx* c, const crispgpu_batch* b, int s, int slot, int a1, int a2, slot_out* o, int* strandpass)
{
	const crispgpu_config* g = &c->cfg;
	double ser = (a1 >= 4 || a2 >= 4) ? 0.05 : g->ser;
	double pall[2] = {0, 0};
	int qpvars = 0;
	if (a1 < 4 && a2 < 4) {
		for (int i = 0; i < c->P; i++) {
			double q0, q1;
			o_chernoff(c->ic + i * 24, c->stats + i * 16, a1, a2, ser, &q0, &q1);
			if (q0 <= g->min_qpv && q1 <= g->min_qpv) qpvars += 1;
		}
		o_chernoff(c->counts, c->tstats, a1, a2, pow(0.1, 0.1 * g->qmin), &pall[0], &pall[1]);
	}
	o->qvpf = pall[0];
	o->qvpr = pall[1];
	int C1 = c->counts[a2], N1 = C1 + c->counts[a1], C2 = c->counts[a2 + NA], N2 = C2 + c->counts[a1 + NA];
	double strandpv = o_fet(C1, N1, C2, N2);
	int type = 0;
	if (o->ctpval <= g->thresh1) type = 1;
	else if (o->ctpval <= -3 && pall[0] <= -5 && pall[1] <= -5) type = 2;
	else if (o->ctpval <= -2 && pall[0] + pall[1] <= -10 && strandpv >= -1) type = 3;
	else if (strandpv >= -1 && pall[0] + pall[1] <= -20 && c->counts[a2] >= 5 && c->counts[a2 + NA] >= 5) type = 4; /* low_coverage_strand is always 1, :156-166 */
	else if (qpvars >= 2) type = 5;
	if (type > 0) {
		int variants = o_legacy_pools(c, b, s, slot, a1, a2, o->ctpval, strandpass);
		if (variants > 0) {
			o->varpools = variants;
			o->variantpools = variants;
			o->allelecount = 0;
			return 1;
		}
	}
	return 0;
}

/* ------------------------------------------------------------------------------------------------------------
 * per-site driver — crisp/newcrispcaller.c:184-270
 * ---------------------------------------------------------------------------------------------------------- */
static void load_site(oracle_ctx* c, const crispgpu_batch* b, int s)
{
	const int P = c->P;
	memcpy(c->ic, b->indcounts + (size_t)s * P * 24, sizeof(int) * P * 24);
	memcpy(c->counts, b->counts + (size_t)s * 24, sizeof(int) * 24);
	for (int i = 0; i < P; i++)
		for (int q = 0; q < 16; q++) {
			c->qhigh[i * 16 + q] = b->qhigh ? b->qhigh[((size_t)s * P + i) * 16 + q] : (double)IC(c, i, q);
			c->stats[i * 16 + q] = b->stats ? b->stats[((size_t)s * P + i) * 16 + q] : 0.0;
		}
	for (int q = 0; q < 16; q++) c->tstats[q] = b->tstats ? b->tstats[(size_t)s * 16 + q] : 0.0;
}

/* remove_ambiguous_bases(+-1) — parsebam/process_indel_variant.c:20-66, from the precomputed decrement row */
static void apply_ambiguity(oracle_ctx* c, const crispgpu_batch* b, int row, int refbase, int sign)
{
	for (int i = 0; i < c->P; i++) {
		const int32_t* d = b->amb_dec + ((size_t)row * c->P + i) * 4;
		IC(c, i, refbase) -= sign * d[0];
		IC(c, i, refbase + NA) -= sign * d[1];
		IC(c, i, refbase + 2 * NA) -= sign * (d[2] + d[3]);
		c->counts[refbase] -= sign * d[0];
		c->counts[refbase + NA] -= sign * d[1];
		c->counts[refbase + 2 * NA] -= sign * (d[2] + d[3]);
		for (int st = 0; st < 2; st++) {
			int cnt = d[st] + d[2 + st];
			double ep = b->amb_ep ? b->amb_ep[((size_t)row * c->P + i) * 2 + st] : 0.0;
			c->stats[i * 16 + refbase + st * NA] -= sign * ep;
			c->tstats[refbase + st * NA] -= sign * ep;
			c->qhigh[i * 16 + refbase + st * NA] -= c->cfg.use_qv == 1 ? sign * (cnt - ep) : sign * cnt;
		}
	}
}

int crisp_oracle_run(void* h, const crispgpu_batch* b, crisp_host_results* out, int rng_kind, long seed_unused, char* vcf, size_t vcf_cap, size_t* vcf_len)
{
	oracle_ctx* c = (oracle_ctx*)h;
	if (!c || !b || !out) return CRISPGPU_ERR_ARG;
	const int P = c->P, n = b->n_sites;
	c->rng_kind = rng_kind;
	if (rng_kind == CRISP_RNG_DRAND48) srand48(seed_unused);
	if (vcf && vcf_cap) vcf[0] = 0;
	if (vcf_len) *vcf_len = 0;
	out->n_sites = n;
	out->n_calls = 0;
	int* AC = (int*)malloc(sizeof(int) * P);
	double* QV = (double*)malloc(sizeof(double) * P);
	for (int s = 0; s < n; s++) {
		load_site(c, b, s);
		int varalleles = 0, strandpass = 0;
		const uint8_t* mal = b->maxvec_al + s * 8;
		const int32_t* mct = b->maxvec_ct + s * 8;
		double paf[2] = {0.0, 0.0};
		for (int al = 1; al <= 5; al++) {
			const int slot = al - 1, oix = s * 5 + slot;
			slot_out o;
			memset(&o, 0, sizeof(o));
			o.status = CRISPGPU_SLOT_UNTESTED;
			int called = 0;
			if (mct[al] >= 3) {
				int a1 = mal[0], a2 = mal[al], a3 = mal[al == 1 ? 2 : 1];
				int hp2 = b->hplen ? b->hplen[s * 8 + a2] : 0, hp3 = b->hplen ? b->hplen[s * 8 + a3] : 0;
				int row = -1;
				if ((a2 >= 4 && hp2 > 0) || (a3 >= 4 && hp3 >= 100)) {
					row = b->amb_idx ? b->amb_idx[s * 5 + slot] : -1;
					if (row >= 0) apply_ambiguity(c, b, row, b->refbase[s], +1);
				}
				if (o_evaluate(c, b, s, slot, a1, a2, a3, paf, &o)) {
					if (c->cfg.em_flag >= 1) {
						called = o_compute_gll(c, b, s, a1, a2, paf, &o, AC, QV);
						if (out->genll) {
							double* src[4] = {c->G, c->Gf, c->Gr, c->Gb};
							for (int t = 0; t < 4; t++)
								for (int i = 0; i < P; i++)
									for (int j = 0; j <= c->ploidy[i] && j < out->genll_stride; j++)
										out->genll[((((size_t)s * 5 + slot) * 4 + t) * P + i) * out->genll_stride + j] = GX(c, src[t], i, j);
						}
						if (called) o.varpools = o_pool_statistics(c, b, s, slot, a2);
					} else {
						called = o_legacy_caller(c, b, s, slot, a1, a2, &o, &strandpass);
					}
				}
				if (row >= 0) apply_ambiguity(c, b, row, b->refbase[s], -1);
			}
			out->status[oix] = called ? CRISPGPU_SLOT_CALLED : o.status;
			out->allele[oix] = mal[al];
			out->call_rank[oix] = called ? varalleles : -1;
			out->call_row[oix] = -1;
			out->paf[oix] = o.paf; out->ctpval[oix] = o.ctpval; out->chi2pval[oix] = o.chi2pval;
			out->af_ml[oix] = o.af_ml; out->delta[oix] = o.delta; out->deltaf[oix] = o.deltaf; out->hwe[oix] = o.hwe;
			out->qvpf[oix] = o.qvpf; out->qvpr[oix] = o.qvpr;
			out->varpools[oix] = o.varpools; out->allelecount[oix] = o.allelecount; out->variantpools[oix] = o.variantpools;
			if (out->em_iters) out->em_iters[oix] = o.em_iters;
			if (out->perm_k) out->perm_k[oix] = o.perm_k;
			if (out->perm_hits) out->perm_hits[oix] = o.perm_hits;
			if (out->assoc_p) { out->assoc_p[oix] = o.assoc[0]; out->assoc_llr[oix] = o.assoc[1]; out->assoc_or[oix] = o.assoc[2]; out->assoc_p1[oix] = o.assoc[3]; out->assoc_llr1[oix] = o.assoc[4]; out->assoc_or1[oix] = o.assoc[5]; }
			if (called) {
				varalleles++;
				if (c->cfg.em_flag >= 1 && out->n_calls < out->max_calls) {
					int row = out->n_calls++;
					out->call_row[oix] = row;
					memcpy(out->ac + (size_t)row * P, AC, sizeof(int) * P);
					memcpy(out->qv + (size_t)row * P, QV, sizeof(double) * P);
				}
			}
		}
		out->varalleles[s] = varalleles;
		out->strandpass[s] = strandpass;
	}
	free(AC);
	free(QV);
	return 0;
}

/* scalar known-answer entry points */
double crisp_oracle_ncr(int n, int r) { return o_ncr(n, r); }
double crisp_oracle_fet(int c1, int r1, int c2, int r2) { return o_fet(c1, r1, c2, r2); }
double crisp_oracle_minreads_pvalue(int R, int A, double e) { return o_minreads_pvalue(R, A, e); }
double crisp_oracle_binomial_pvalue(int R, int A, double e) { return o_binomial_pvalue(R, A, e); }
double crisp_oracle_kf_lgamma(double z) { return o_lgamma(z); }
double crisp_oracle_kf_gammaq(double s, double z) { return o_gammaq(s, z); }
void crisp_oracle_chernoff(int* counts, double* stats, int refbase, int altbase, double ser, double* p0, double* p1)
{
	o_chernoff(counts, stats, refbase, altbase, ser, p0, p1);
}
double crisp_oracle_exact_2xk(const int* N, const int* A, int size)
{
	double p = 1.0;
	o_exact_2xk(N, A, size, &p);
	return p;
}
void crisp_oracle_philox(const uint32_t ctr[4], const uint32_t key[2], uint32_t out[4]) { philox4x32_10(ctr, key, out); }

/* ------------------------------------------------------------------------------------------------------------
 * candidate screen — variantcalls.c:62-92
 * ---------------------------------------------------------------------------------------------------------- */
/* sort_allelecounts, parsebam/allelecounts.c:134-183 with IUPAC off: read count per allele, the reference allele's entry
 * swapped to the front, then `if (ct[i] < ct[j]) swap` for i = 1..6, j = i+1..7 */
static void o_sort_allelecounts(const int32_t* counts, int refbase, uint8_t* al_out, int32_t* ct_out)
{
	int ct[8], al[8];
	for (int i = 0; i < 8; i++) { ct[i] = counts[i] + counts[i + 8] + counts[i + 16]; al[i] = i; }
	if (refbase >= 1 && refbase < 8) {
		int t = ct[0]; ct[0] = ct[refbase]; ct[refbase] = t;
		t = al[0]; al[0] = al[refbase]; al[refbase] = t;
	}
	for (int i = 1; i < 7; i++)
		for (int j = i + 1; j < 8; j++)
			if (ct[i] < ct[j]) {
				int t = ct[i]; ct[i] = ct[j]; ct[j] = t;
				t = al[i]; al[i] = al[j]; al[j] = t;
			}
	for (int i = 0; i < 8; i++) { al_out[i] = (uint8_t)al[i]; ct_out[i] = ct[i]; }
}

/* potentialvariant of jointvariantcaller (variantcalls.c:76-92) for PICALL == 3 */
static int o_potential_variant(const int32_t* indcounts, const int32_t* ploidy, int P, int refbase)
{
	int pot = 0;
	for (int i = 0; i < 8; i++) {
		if (i == refbase) continue;
		for (int j = 0; j < P; j++) {
			const int32_t* r = indcounts + (size_t)j * 24;
			const int t = r[i] + r[i + 8] + r[i + 16];
			if (t >= 3) pot += 2;
			else if (t >= 1 && ploidy[j] == 2) pot += t;
		}
	}
	return pot;
}

void crisp_oracle_screen(void* h, int n, const uint8_t* refbase, const int32_t* counts, const int32_t* indcounts, uint8_t* maxvec_al, int32_t* maxvec_ct, int32_t* potential)
{
	oracle_ctx* c = (oracle_ctx*)h;
	for (int s = 0; s < n; s++) {
		o_sort_allelecounts(counts + (size_t)s * 24, refbase[s], maxvec_al + (size_t)s * 8, maxvec_ct + (size_t)s * 8);
		potential[s] = o_potential_variant(indcounts + (size_t)s * c->P * 24, c->ploidy, c->P, refbase[s]);
	}
}
