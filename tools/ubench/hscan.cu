// Phase-H micro-benchmark: one warp, 30 lanes, each running a sequential scan over a 116-column line in shared memory.
#include <cstdio>
#include <cuda_runtime.h>
constexpr int LS = 129, NL = 30, NC = 116, OFF = 5;
__device__ __forceinline__ long long clk() { long long c; asm volatile("mov.u64 %0, %%clock64;" : "=l"(c)); return c; }

// variant 0: as in k_flow_iter_win (3 register sets, diffs formed by the scanning lane)
__device__ double scan_v0(double* line, double S, int nch)
{
    double A0, A1, A2, A3, B0, B1, B2, B3, C0, C1, C2, C3;
#define LOAD4(P, ch) { const double* q_ = line + 4 * (ch); P##0 = __dsub_rn(q_[OFF], q_[0]); P##1 = __dsub_rn(q_[OFF + 1], q_[1]); P##2 = __dsub_rn(q_[OFF + 2], q_[2]); P##3 = __dsub_rn(q_[OFF + 3], q_[3]); }
#define CHAIN4(P, ch) { double* q_ = line + 4 * (ch); S = __dadd_rn(S, P##0); q_[0] = S; S = __dadd_rn(S, P##1); q_[1] = S; S = __dadd_rn(S, P##2); q_[2] = S; S = __dadd_rn(S, P##3); q_[3] = S; }
    LOAD4(A, 0);
    if (nch > 1) LOAD4(B, 1);
    for (int ch = 0;;) {
        if (ch + 2 < nch) LOAD4(C, ch + 2);
        CHAIN4(A, ch);
        if (++ch == nch) break;
        if (ch + 2 < nch) LOAD4(A, ch + 2);
        CHAIN4(B, ch);
        if (++ch == nch) break;
        if (ch + 2 < nch) LOAD4(B, ch + 2);
        CHAIN4(C, ch);
        if (++ch == nch) break;
    }
    return S;
}
// variant 1: differences already in the line (someone else formed them): load, add, store
__device__ double scan_v1(double* line, double S, int nch)
{
    for (int ch = 0; ch < nch; ch++) {
        double* q = line + 4 * ch;
        const double d0 = q[0], d1 = q[1], d2 = q[2], d3 = q[3];
        S = __dadd_rn(S, d0); q[0] = S; S = __dadd_rn(S, d1); q[1] = S;
        S = __dadd_rn(S, d2); q[2] = S; S = __dadd_rn(S, d3); q[3] = S;
    }
    return S;
}
// variant 2: like 1, fully unrolled by 8 columns with the next 8 differences loaded ahead
__device__ double scan_v2(double* line, double S, int nch)
{
    double a[8], b[8];
#pragma unroll
    for (int u = 0; u < 8; u++) a[u] = line[u];
    int i = 0;
    for (; i + 16 <= 4 * nch; i += 8) {
#pragma unroll
        for (int u = 0; u < 8; u++) b[u] = line[i + 8 + u];
#pragma unroll
        for (int u = 0; u < 8; u++) { S = __dadd_rn(S, a[u]); line[i + u] = S; }
#pragma unroll
        for (int u = 0; u < 8; u++) a[u] = b[u];
    }
#pragma unroll
    for (int u = 0; u < 8; u++) { S = __dadd_rn(S, a[u]); line[i + u] = S; }
    i += 8;
    for (; i < 4 * nch; i++) { S = __dadd_rn(S, line[i]); line[i] = S; }
    return S;
}
// variant 3: chain only; the 8 sums of a chunk stay in registers and are stored after the chain (in-order issue: a store
// right behind its DADD blocks the next DADD of the chain for the store's issue time)
__device__ double scan_v3(double* line, double S, int nch)
{
    double a[8], b[8], s[8];
#pragma unroll
    for (int u = 0; u < 8; u++) a[u] = line[u];
    int i = 0;
    for (; i + 16 <= 4 * nch; i += 8) {
#pragma unroll
        for (int u = 0; u < 8; u++) b[u] = line[i + 8 + u];
#pragma unroll
        for (int u = 0; u < 8; u++) { S = __dadd_rn(S, a[u]); s[u] = S; }
#pragma unroll
        for (int u = 0; u < 8; u++) line[i + u] = s[u];
#pragma unroll
        for (int u = 0; u < 8; u++) a[u] = b[u];
    }
#pragma unroll
    for (int u = 0; u < 8; u++) { S = __dadd_rn(S, a[u]); line[i + u] = S; }
    i += 8;
    for (; i < 4 * nch; i++) { S = __dadd_rn(S, line[i]); line[i] = S; }
    return S;
}
// variant 4: diff + chain, chunk of 8, diffs of the next chunk formed before the chain, stores after the chain
__device__ double scan_v4(double* line, double S, int nch)
{
    double a[8], b[8], s[8];
#pragma unroll
    for (int u = 0; u < 8; u++) a[u] = __dsub_rn(line[u + OFF], line[u]);
    int i = 0;
    for (; i + 16 <= 4 * nch; i += 8) {
#pragma unroll
        for (int u = 0; u < 8; u++) b[u] = __dsub_rn(line[i + 8 + u + OFF], line[i + 8 + u]);
#pragma unroll
        for (int u = 0; u < 8; u++) { S = __dadd_rn(S, a[u]); s[u] = S; }
#pragma unroll
        for (int u = 0; u < 8; u++) line[i + u] = s[u];
#pragma unroll
        for (int u = 0; u < 8; u++) a[u] = b[u];
    }
#pragma unroll
    for (int u = 0; u < 8; u++) { S = __dadd_rn(S, a[u]); line[i + u] = S; }
    i += 8;
    for (; i < 4 * nch; i++) { S = __dadd_rn(S, __dsub_rn(line[i + OFF], line[i])); line[i] = S; }
    return S;
}
// variant 5: like 4 but written with inline PTX so that the order chain / diffs / stores is ours:
// per column: chain add, then one diff of the next chunk, stores delayed by one chunk
__device__ double scan_v5(double* line, double S, int nch)
{
    double a[4], b[4], s[4], sp[4];
#pragma unroll
    for (int u = 0; u < 4; u++) a[u] = __dsub_rn(line[u + OFF], line[u]);
    // chunk ch: chain over a[]; meanwhile form b[] = diffs of chunk ch+1 and store sp[] (sums of chunk ch-1)
    int ch = 0;
    bool havep = false;
    for (; ch < nch; ch++) {
        const bool nxt = ch + 1 < nch;
        double l0[4], l1[4];
        if (nxt) {
#pragma unroll
            for (int u = 0; u < 4; u++) { l0[u] = line[4 * (ch + 1) + u + OFF]; l1[u] = line[4 * (ch + 1) + u]; }
        }
#pragma unroll
        for (int u = 0; u < 4; u++) {
            S = __dadd_rn(S, a[u]); s[u] = S;
            if (nxt) b[u] = __dsub_rn(l0[u], l1[u]);
            if (havep) line[4 * (ch - 1) + u] = sp[u];
        }
#pragma unroll
        for (int u = 0; u < 4; u++) { sp[u] = s[u]; a[u] = b[u]; }
        havep = true;
    }
#pragma unroll
    for (int u = 0; u < 4; u++) line[4 * (nch - 1) + u] = sp[u];
    return S;
}
// variant 3: differences in the line, results kept in registers only every column but stored as pairs (even LS needed) -- here plain
template <int V>
__global__ void k(double* out, long long* cyc, int nch, int lanes)
{
    __shared__ double tile[NL * LS + 64];
    for (int i = threadIdx.x; i < NL * LS; i += blockDim.x) tile[i] = 1.0 + i * 1e-3;
    __syncthreads();
    long long t0 = 0, t1 = 0;
    double S = 0;
    if (threadIdx.x < 32) {
        t0 = clk();
        for (int rep = 0; rep < 8; rep++) {
            if (threadIdx.x < lanes) {
                double* line = tile + threadIdx.x * LS;
                if (V == 0) S = scan_v0(line, S, nch);
                if (V == 1) S = scan_v1(line, S, nch);
                if (V == 2) S = scan_v2(line, S, nch);
                if (V == 3) S = scan_v3(line, S, nch);
                if (V == 4) S = scan_v4(line, S, nch);
                if (V == 5) S = scan_v5(line, S, nch);
            }
            __syncwarp();
        }
        t1 = clk();
    }
    out[threadIdx.x] = S;
    if (threadIdx.x == 0) cyc[0] = (t1 - t0) / 8;
}
int main()
{
    double* o; long long* c;
    cudaMalloc(&o, 8 * 1024); cudaMalloc(&c, 64);
    long long h;
    for (int lanes : {30}) {
        k<0><<<1, 128>>>(o, c, NC / 4, lanes); cudaMemcpy(&h, c, 8, cudaMemcpyDeviceToHost);
        printf("lanes %2d v0 (diff+chain): %lld cycles per %d columns = %.1f cyc/col\n", lanes, h, NC, (double)h / NC);
        k<1><<<1, 128>>>(o, c, NC / 4, lanes); cudaMemcpy(&h, c, 8, cudaMemcpyDeviceToHost);
        printf("lanes %2d v1 (chain only, rolled): %lld = %.1f cyc/col\n", lanes, h, (double)h / NC);
        k<2><<<1, 128>>>(o, c, NC / 4, lanes); cudaMemcpy(&h, c, 8, cudaMemcpyDeviceToHost);
        printf("lanes %2d v2 (chain only, lookahead 8): %lld = %.1f cyc/col\n", lanes, h, (double)h / NC);
        k<3><<<1, 128>>>(o, c, NC / 4, lanes); cudaMemcpy(&h, c, 8, cudaMemcpyDeviceToHost);
        printf("lanes %2d v3 (chain only, stores after chain of 8): %lld = %.1f cyc/col\n", lanes, h, (double)h / NC);
        k<4><<<1, 128>>>(o, c, NC / 4, lanes); cudaMemcpy(&h, c, 8, cudaMemcpyDeviceToHost);
        printf("lanes %2d v4 (diff+chain, stores after chain of 8): %lld = %.1f cyc/col\n", lanes, h, (double)h / NC);
        k<5><<<1, 128>>>(o, c, NC / 4, lanes); cudaMemcpy(&h, c, 8, cudaMemcpyDeviceToHost);
        printf("lanes %2d v5 (diff+chain interleaved, stores one chunk late): %lld = %.1f cyc/col\n", lanes, h, (double)h / NC);
    }
    return 0;
}
