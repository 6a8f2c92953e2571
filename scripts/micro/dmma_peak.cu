// FP64 tensor-core (DMMA.8x8x4) versus DFMA issue throughput on the device: does the dense likelihood contraction of
// k_em belong on mma.sync?   nvcc -O3 -gencode arch=compute_100a,code=sm_100a -o dmma_peak dmma_peak.cu
#include <cstdio>
__global__ void k_dmma(double* out, int iters)
{
	double a = threadIdx.x * 1e-9, b = 1.0 + threadIdx.x * 1e-9;
	double c[8][2] = {};
	for (int i = 0; i < iters; i++) {
#pragma unroll
		for (int u = 0; u < 8; u++)
			asm volatile("mma.sync.aligned.m8n8k4.row.col.f64.f64.f64.f64 {%0,%1}, {%2}, {%3}, {%0,%1};" : "+d"(c[u][0]), "+d"(c[u][1]) : "d"(a), "d"(b));
	}
	double s = 0;
	for (int u = 0; u < 8; u++) s += c[u][0] + c[u][1];
	if (s == 12345.678) out[0] = s;
}
__global__ void k_dfma(double* out, int iters)
{
	double a[16];
	for (int u = 0; u < 16; u++) a[u] = threadIdx.x * 1e-9 + u;
	const double m = 1.0000001, c = 1e-7;
	for (int i = 0; i < iters; i++) {
#pragma unroll
		for (int u = 0; u < 16; u++) a[u] = fma(a[u], m, c);
	}
	double s = 0;
	for (int u = 0; u < 16; u++) s += a[u];
	if (s == 12345.678) out[0] = s;
}
int main()
{
	double* d; cudaMalloc(&d, 8);
	cudaDeviceProp p; cudaGetDeviceProperties(&p, 0);
	cudaEvent_t e0, e1; cudaEventCreate(&e0); cudaEventCreate(&e1);
	const int iters = 20000, grid = p.multiProcessorCount * 8, threads = 256;
	float ms;
	k_dmma<<<grid, threads>>>(d, 100); k_dfma<<<grid, threads>>>(d, 100); cudaDeviceSynchronize();
	cudaEventRecord(e0); k_dmma<<<grid, threads>>>(d, iters); cudaEventRecord(e1); cudaEventSynchronize(e1); cudaEventElapsedTime(&ms, e0, e1);
	double fl = (double)grid * (threads / 32) * iters * 8 * (8 * 8 * 4 * 2);
	printf("DMMA.8x8x4: %.3f ms, %.2f TFLOP/s\n", ms, fl / ms / 1e9);
	cudaEventRecord(e0); k_dfma<<<grid, threads>>>(d, iters); cudaEventRecord(e1); cudaEventSynchronize(e1); cudaEventElapsedTime(&ms, e0, e1);
	fl = (double)grid * threads * iters * 16 * 2;
	printf("DFMA:       %.3f ms, %.2f TFLOP/s\n", ms, fl / ms / 1e9);
	printf("%s\n", cudaGetErrorString(cudaGetLastError()));
	return 0;
}
