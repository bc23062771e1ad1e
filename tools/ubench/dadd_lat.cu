// Dependent-issue latency of DADD / FADD / LDS.64 round trip on sm_100a (single warp, clock64).
#include <cstdio>
#include <cuda_runtime.h>
__global__ void k(double* out, long long* cyc, double a, float fa)
{
    __shared__ double sm[256];
    sm[threadIdx.x] = a;
    __syncthreads();
    double s = threadIdx.x;
    long long t0 = clock64();
#pragma unroll 64
    for (int i = 0; i < 4096; i++) s = __dadd_rn(s, a);
    long long t1 = clock64();
    float f = threadIdx.x;
#pragma unroll 64
    for (int i = 0; i < 4096; i++) f = __fadd_rn(f, fa);
    long long t2 = clock64();
    // chain through shared memory: S += sm[idx]; sm[idx] = S (store->load of a different address: no dependency)
    double s2 = s;
    int idx = threadIdx.x;
#pragma unroll 16
    for (int i = 0; i < 1024; i++) { s2 = __dadd_rn(s2, sm[(idx + i) & 255]); }
    long long t3 = clock64();
    // two independent chains per thread
    double u = s, v = s2;
#pragma unroll 32
    for (int i = 0; i < 4096; i++) { u = __dadd_rn(u, a); v = __dadd_rn(v, a); }
    long long t4 = clock64();
    out[threadIdx.x] = s + f + s2 + u + v;
    if (threadIdx.x == 0) { cyc[0] = t1 - t0; cyc[1] = t2 - t1; cyc[2] = t3 - t2; cyc[3] = t4 - t3; }
}
int main()
{
    double* o; long long* c;
    cudaMalloc(&o, 8 * 1024); cudaMalloc(&c, 64);
    for (int nt : {32, 15, 128}) {
        k<<<1, nt>>>(o, c, 1.0000001, 1.0001f);
        long long h[4];
        cudaMemcpy(h, c, 32, cudaMemcpyDeviceToHost);
        printf("threads %3d: DADD chain %.2f cyc/op, FADD chain %.2f, DADD+LDS operand chain %.2f, 2 indep DADD chains %.2f cyc/pair\n", nt,
               h[0] / 4096.0, h[1] / 4096.0, h[2] / 1024.0, h[3] / 4096.0);
    }
    return 0;
}
