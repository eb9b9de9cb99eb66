// latency microbenchmarks for the block-wavefront sweep: dependent DFMA, LDS->DFMA->STS round trip, CTA barrier (256 / 192 threads)
#include <cstdio>
#include <cuda_runtime.h>
__global__ void k(double *out, long long *t, int n) {
    __shared__ double sh[512];
    const int tid = threadIdx.x;
    sh[tid] = 1.0 + tid; sh[tid + 256] = 0.5;
    __syncthreads();
    double a = sh[tid], b = 1.0000001, c = 1e-9;
    long long t0 = clock64();
#pragma unroll 16
    for (int i = 0; i < n; ++i) a = __fma_rn(a, b, c);
    long long t1 = clock64();
    // LDS -> DFMA -> STS chain through shared memory (same thread, address depends on nothing)
    volatile double *vs = sh;
    for (int i = 0; i < n; ++i) { double y = vs[tid]; y = __fma_rn(y, b, c); vs[tid] = y; }
    long long t2 = clock64();
    for (int i = 0; i < n; ++i) __syncthreads();
    long long t3 = clock64();
    for (int i = 0; i < n; ++i) { double y = vs[(tid + 1) & 255]; y = __fma_rn(y, b, c); __syncthreads(); vs[tid] = y; __syncthreads(); }
    long long t4 = clock64();
    if (tid < 192) for (int i = 0; i < n; ++i) asm volatile("bar.sync 1, 192;" ::: "memory");
    long long t5 = clock64();
    if (tid < 32) for (int i = 0; i < n; ++i) __syncwarp();
    long long t6 = clock64();
    if (tid == 0) { t[0] = t1 - t0; t[1] = t2 - t1; t[2] = t3 - t2; t[3] = t4 - t3; t[4] = t5 - t4; t[5] = t6 - t5; }
    out[tid] = a + vs[tid];
}
int main() {
    double *o; long long *t, h[6]; const int n = 4096;
    cudaMalloc(&o, 8 * 256); cudaMalloc(&t, 8 * 6);
    for (int r = 0; r < 2; ++r) k<<<1, 256>>>(o, t, n);
    cudaMemcpy(h, t, 48, cudaMemcpyDeviceToHost);
    printf("cycles per: dependent DFMA %.1f | LDS+DFMA+STS %.1f | barrier(256) %.1f | LDS+DFMA+bar+STS+bar %.1f | bar.sync 192 %.1f | syncwarp %.1f\n",
           h[0] / (double)n, h[1] / (double)n, h[2] / (double)n, h[3] / (double)n, h[4] / (double)n, h[5] / (double)n);
    return 0;
}
