// Single-warp issue intervals on sm_100a: independent DADDs, LDS.64, STS.64, mixed (clock64).
#include <cstdio>
#include <cuda_runtime.h>
__device__ __forceinline__ long long clk() { long long c; asm volatile("mov.u64 %0, %%clock64;" : "=l"(c)); return c; }
__global__ void k(double* out, long long* cyc, double a)
{
    __shared__ double sm[1024];
    for (int i = threadIdx.x; i < 1024; i += blockDim.x) sm[i] = a * i;
    __syncthreads();
    double v[8];
#pragma unroll
    for (int u = 0; u < 8; u++) v[u] = threadIdx.x + u;
    long long t0 = clk();
#pragma unroll 8
    for (int i = 0; i < 512; i++) {
#pragma unroll
        for (int u = 0; u < 8; u++) v[u] = __dadd_rn(v[u], a);
    }
    long long t1 = clk();
    double acc = 0;
    const double* p = sm + threadIdx.x;
#pragma unroll 8
    for (int i = 0; i < 512; i++) {
        double l[8];
#pragma unroll
        for (int u = 0; u < 8; u++) l[u] = p[((i + u) & 15) * 32];
#pragma unroll
        for (int u = 0; u < 8; u++) acc += l[u];
    }
    long long t2 = clk();
    double* q = sm + threadIdx.x;
#pragma unroll 8
    for (int i = 0; i < 512; i++) {
#pragma unroll
        for (int u = 0; u < 8; u++) q[((i + u) & 15) * 32] = v[u];
    }
    long long t3 = clk();
    float f[8];
#pragma unroll
    for (int u = 0; u < 8; u++) f[u] = threadIdx.x + u;
#pragma unroll 8
    for (int i = 0; i < 512; i++) {
#pragma unroll
        for (int u = 0; u < 8; u++) f[u] = __fadd_rn(f[u], 1.5f);
    }
    long long t4 = clk();
    double s = 0;
#pragma unroll
    for (int u = 0; u < 8; u++) s += v[u] + f[u];
    out[threadIdx.x] = s + acc;
    if (threadIdx.x == 0) { cyc[0] = t1 - t0; cyc[1] = t2 - t1; cyc[2] = t3 - t2; cyc[3] = t4 - t3; }
}
int main()
{
    double* o; long long* c;
    cudaMalloc(&o, 8 * 1024); cudaMalloc(&c, 64);
    for (int nt : {32, 128, 512}) {
        k<<<1, nt>>>(o, c, 1.0000001);
        long long h[4];
        cudaMemcpy(h, c, 32, cudaMemcpyDeviceToHost);
        printf("threads %3d: 8 indep DADD chains %.2f cyc/DADD; LDS.64+DADD %.2f cyc/pair; STS.64 %.2f cyc; 8 indep FADD %.2f cyc\n", nt,
               h[0] / 4096.0, h[1] / 4096.0, h[2] / 4096.0, h[3] / 4096.0);
    }
    return 0;
}
