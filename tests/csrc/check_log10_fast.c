/* Accuracy of log10_fast (crisp_b200/csrc/dmath.cuh) on the CPU: the same operations in the same order (fma() is the
 * device's DFMA; compile with -ffp-contract=off), the table and constants taken from dmath.cuh by the test that builds
 * this file (log10_tab.h).  Prints the maximum error in ulp against the 80-bit log10l. */
#include <math.h>
#include <stdint.h>
#include <stdio.h>
#include <stdlib.h>
#include <string.h>
#include "log10_tab.h"

static double log10_fast(double x)
{
	uint64_t b;
	memcpy(&b, &x, 8);
	const int32_t hi = (int32_t)(b >> 32);
	if ((uint32_t)(hi - 0x00100000) >= 0x7FE00000u) return log10(x);
	const int i = (hi >> 13) & 0x7f;
	const int up = i >= 53;
	const uint64_t mb = ((uint64_t)((uint32_t)(hi & 0x000FFFFF) | (up ? 0x3FE00000u : 0x3FF00000u)) << 32) | (b & 0xffffffffu);
	double m;
	memcpy(&m, &mb, 8);
	const double rc = kLog10Tab[2 * i], lc = kLog10Tab[2 * i + 1];
	const double ed = (double)((hi >> 20) + up - 1023);
	const double r = fma(m, rc, -1.0);
	const double r2 = r * r;
	const double A = fma(kLog10K[1], r, kLog10K[0]), B = fma(kLog10K[3], r, kLog10K[2]), C = fma(kLog10K[5], r, kLog10K[4]), D = fma(kLog10K[7], r, kLog10K[6]);
	const double r4 = r2 * r2;
	const double AB = fma(B, r2, A), CD = fma(D, r2, C);
	const double in = fma(r4, kLog10K[8], CD);
	const double P = fma(r4, in, AB);
	const double t = fma(ed, kLog10K[9], lc);
	const double s = fma(ed, kLog10K[10], r * P);
	return t + s;
}

int main(int argc, char** argv)
{
	const long n_points = argc > 1 ? atol(argv[1]) : 4000000;
	double maxulp = 0, maxabs1 = 0;
	srand48(1);
	for (long n = 0; n < n_points; n++) {
		double x;
		switch (n % 4) {
		case 0: x = pow(10.0, -12 + 16 * drand48()); break;                       /* the range of p, 1-p and the sums */
		case 1: x = 1.0 + (drand48() - 0.5) * pow(10.0, -8 * drand48()); break;   /* around 1 */
		case 2: x = 1e-3 + drand48() * 1e3; break;
		default: x = 0.5 + drand48() * 1.5;
		}
		const double got = log10_fast(x);
		const long double want = log10l((long double)x);
		const double err = fabs((double)((long double)got - want));
		const double a = fabs((double)want);
		const double ulp = err / (nextafter(a, INFINITY) - a);
		if (ulp > maxulp) maxulp = ulp;
		if (a < 1e-2 && err > maxabs1) maxabs1 = err;
	}
	printf("%.4f %.4g\n", maxulp, maxabs1);
	return !(log10_fast(1.0) == 0.0 && log10_fast(10.0) == 1.0 && isinf(log10_fast(0.0)) && isnan(log10_fast(-1.0)));
}
