// Stage 4: bilinear remap of a neighbour slice fused with the Gaussian-weighted accumulation; plus the
// no-OF wrap-around 1-D Gaussian and the batched transpose used by the X pass. sm_100a.
//
// Reference behaviour:
//   warp_slice                  /root/reference/src/flowdenoising.py:55-63  (cv2.remap INTER_LINEAR,
//                               BORDER_REPLICATE, float32 map -> OpenCV's 1/32-pixel quantiser)
//   tmp_slice += warped*k[i]    :316, :317, :323  (NumPy >= 2: float64 product and sum, rounded to float32
//                               once per tap)
//   GaussianDenoising slices    :133-158 (no-OF path)
// Compiled with -fmad=false.
#include "fdn_internal.cuh"

namespace fdn {

// ------------------------------------------------------------------------------------------------
// K4: acc = f32(f64(acc) + f64(remap(neigh, flow)) * weight)
// ------------------------------------------------------------------------------------------------
__global__ void __launch_bounds__(128)
k_warp_acc(const float* __restrict__ neigh, int64_t n_ss, int64_t n_rs, SlotMap n_map, const float2* __restrict__ flow,
           double weight, float* __restrict__ acc, int64_t a_ss, int64_t a_rs, int H, int W, int first)
{
    const int x = blockIdx.x * 128 + threadIdx.x;
    const int y = blockIdx.y;
    const int b = blockIdx.z;
    if (x >= W) return;
    const float* src = neigh + (int64_t)n_map.slot(b) * n_ss;
    float v;
    if (flow) {
        const float2 f = flow[((int64_t)b * H + y) * W + x];
        // map = (flow + grid).astype(float32): one rounding of the exact sum
        const float mx = __fadd_rn(f.x, (float)x), my = __fadd_rn(f.y, (float)y);
        // cvRound(map * INTER_TAB_SIZE): round half to even
        const int sx = __float2int_rn(__fmul_rn(mx, 32.f)), sy = __float2int_rn(__fmul_rn(my, 32.f));
        const int ax = sx & 31, ay = sy & 31;
        int ix = sx >> 5, iy = sy >> 5;
        ix = min(max(ix, -32768), 32767);  // saturate_cast<short>
        iy = min(max(iy, -32768), 32767);
        const float tx1 = __fmul_rn((float)ax, 0.03125f), tx0 = __fsub_rn(1.f, tx1);
        const float ty1 = __fmul_rn((float)ay, 0.03125f), ty0 = __fsub_rn(1.f, ty1);
        const float w0 = __fmul_rn(ty0, tx0), w1 = __fmul_rn(ty0, tx1), w2 = __fmul_rn(ty1, tx0),
                    w3 = __fmul_rn(ty1, tx1);
        const int x0 = min(max(ix, 0), W - 1), x1 = min(max(ix + 1, 0), W - 1);
        const int y0 = min(max(iy, 0), H - 1), y1 = min(max(iy + 1, 0), H - 1);
        const float* r0 = src + (int64_t)y0 * n_rs;
        const float* r1 = src + (int64_t)y1 * n_rs;
        const float v0 = r0[x0], v1 = r0[x1], v2 = r1[x0], v3 = r1[x1];
        v = __fadd_rn(__fadd_rn(__fadd_rn(__fmul_rn(v0, w0), __fmul_rn(v1, w1)), __fmul_rn(v2, w2)),
                      __fmul_rn(v3, w3));
    } else {
        v = src[(int64_t)y * n_rs + x];
    }
    float* ap = acc + (int64_t)b * a_ss + (int64_t)y * a_rs + x;
    const double prod = __dmul_rn((double)v, weight);
    *ap = first ? (float)prod : (float)__dadd_rn((double)*ap, prod);
}

int launch_warp_acc(const float* neigh, int64_t n_ss, int64_t n_rs, SlotMap n_map, const float* flow, double weight,
                    float* acc, int64_t a_ss, int64_t a_rs, int n, int H, int W, int first, cudaStream_t st)
{
    for (int b0 = 0; b0 < n; b0 += 65535) {
        const int nb = n - b0 < 65535 ? n - b0 : 65535;
        SlotMap m = n_map;
        m.base += b0;
        dim3 grid((unsigned)cdiv(W, 128), (unsigned)H, (unsigned)nb);
        ProfScope ps(K_WARP_ACC, (double)nb * H * W * (4.0 + (flow ? 8.0 : 0.0) + (first ? 4.0 : 8.0)), st);
        k_warp_acc<<<grid, 128, 0, st>>>(neigh, n_ss, n_rs, m,
                                         flow ? reinterpret_cast<const float2*>(flow) + (int64_t)b0 * H * W : nullptr,
                                         weight, acc + (int64_t)b0 * a_ss, a_ss, a_rs, H, W, first);
        FDN_LAUNCHED("k_warp_acc");
    }
    return FDN_OK;
}

// ------------------------------------------------------------------------------------------------
// K5: no-OF pass, out[s] = fold_i f32(f64(acc) + f64(in[(s+halo+i-r) wrap]) * k[i])
// ------------------------------------------------------------------------------------------------
#define FDN_MAX_KLEN 257
struct Taps64 {
    int klen;
    double k[FDN_MAX_KLEN];
};
struct Taps32 {
    int klen;
    float k[FDN_MAX_KLEN];
};

// Filter along the slice axis of a view. One thread per (y, x) column position and output slice; neighbouring
// slices are re-read through L2 (the 2r+1 slices of a tile stay resident there).
template <bool EXACT>
__global__ void __launch_bounds__(256)
k_gauss_axis(const float* __restrict__ in, float* __restrict__ out, fdn_view v, Taps64 t64, Taps32 t32)
{
    const int x = blockIdx.x * 256 + threadIdx.x;
    const int y = blockIdx.y;
    const int s = blockIdx.z;
    if (x >= v.W) return;
    const int klen = EXACT ? t64.klen : t32.klen;
    const int r = klen >> 1;
    const int64_t off = (int64_t)y * v.in_row_stride + x;
    float acc = 0.f;
    for (int i = 0; i < klen; i++) {
        int j = s + v.halo + i - r;
        if (v.periodic) {
            j %= v.n_in;
            if (j < 0) j += v.n_in;
        }
        const float val = in[(int64_t)j * v.in_slice_stride + off];
        if (EXACT) acc = (float)__dadd_rn((double)acc, __dmul_rn((double)val, t64.k[i]));
        else acc = fmaf(val, t32.k[i], acc);
    }
    out[(int64_t)s * v.out_slice_stride + (int64_t)y * v.out_row_stride + x] = acc;
}

int launch_gauss_axis(const float* in, float* out, const fdn_view& v, const double* k, int klen, int exact,
                      cudaStream_t st)
{
    FDN_CHECK_ARG(klen >= 1 && klen <= FDN_MAX_KLEN && (klen & 1), "kernel length %d unsupported (odd, <= %d)", klen,
                  FDN_MAX_KLEN);
    Taps64 t64;
    Taps32 t32;
    t64.klen = t32.klen = klen;
    for (int i = 0; i < klen; i++) { t64.k[i] = k[i]; t32.k[i] = (float)k[i]; }
    FDN_CHECK_ARG(v.H <= 65535 && v.n_out <= 65535, "view too large for one launch");
    dim3 grid((unsigned)cdiv(v.W, 256), (unsigned)v.H, (unsigned)v.n_out);
    ProfScope ps(K_GAUSS_AXIS, 8.0 * v.n_out * v.H * v.W, st);
    if (exact) k_gauss_axis<true><<<grid, 256, 0, st>>>(in, out, v, t64, t32);
    else k_gauss_axis<false><<<grid, 256, 0, st>>>(in, out, v, t64, t32);
    FDN_LAUNCHED("k_gauss_axis");
    return FDN_OK;
}

// Filter along contiguous rows (x axis, periodic). Row segment + wrap halo staged in shared memory.
#define GR_TILE 512
template <bool EXACT>
__global__ void __launch_bounds__(256)
k_gauss_rows(const float* __restrict__ in, float* __restrict__ out, int W, Taps64 t64, Taps32 t32)
{
    extern __shared__ float s_seg[];  // GR_TILE + klen - 1
    const int klen = EXACT ? t64.klen : t32.klen;
    const int r = klen >> 1;
    const int64_t row = blockIdx.y;
    const int x0 = blockIdx.x * GR_TILE;
    const float* src = in + row * W;
    const int nload = min(GR_TILE, W - x0) + 2 * r;
    for (int i = threadIdx.x; i < nload; i += 256) {
        int j = (x0 - r + i) % W;
        if (j < 0) j += W;
        s_seg[i] = src[j];
    }
    __syncthreads();
    for (int i = threadIdx.x; i < GR_TILE && x0 + i < W; i += 256) {
        float acc = 0.f;
        for (int t = 0; t < klen; t++) {
            const float val = s_seg[i + t];
            if (EXACT) acc = (float)__dadd_rn((double)acc, __dmul_rn((double)val, t64.k[t]));
            else acc = fmaf(val, t32.k[t], acc);
        }
        out[row * W + x0 + i] = acc;
    }
}

int launch_gauss_rows(const float* in, float* out, int64_t rows, int W, const double* k, int klen, int exact,
                      cudaStream_t st)
{
    FDN_CHECK_ARG(klen >= 1 && klen <= FDN_MAX_KLEN && (klen & 1), "kernel length %d unsupported (odd, <= %d)", klen,
                  FDN_MAX_KLEN);
    Taps64 t64;
    Taps32 t32;
    t64.klen = t32.klen = klen;
    for (int i = 0; i < klen; i++) { t64.k[i] = k[i]; t32.k[i] = (float)k[i]; }
    const size_t smem = sizeof(float) * (GR_TILE + klen - 1);
    for (int64_t r0 = 0; r0 < rows; r0 += 65535) {
        const int64_t nr = rows - r0 < 65535 ? rows - r0 : 65535;
        dim3 grid((unsigned)cdiv(W, GR_TILE), (unsigned)nr);
        ProfScope ps(K_GAUSS_ROWS, 8.0 * nr * W, st);
        if (exact) k_gauss_rows<true><<<grid, 256, smem, st>>>(in + r0 * W, out + r0 * W, W, t64, t32);
        else k_gauss_rows<false><<<grid, 256, smem, st>>>(in + r0 * W, out + r0 * W, W, t64, t32);
        FDN_LAUNCHED("k_gauss_rows");
    }
    return FDN_OK;
}

// ------------------------------------------------------------------------------------------------
// Batched transpose of the last two axes with arbitrary outer/row strides (32x32 shared-memory tiles):
//   out[n*out_sn + b*out_sb + a] = in[n*in_sn + a*in_sa + b],  a < A, b < B
// Dense case ([n][A][B] -> [n][B][A]): in_sn = out_sn = A*B, in_sa = B, out_sb = A. The strided form is the
// transposing unpack of the multi-GPU re-slab (flowdenoising_b200/dist.py).
// ------------------------------------------------------------------------------------------------
__global__ void __launch_bounds__(256)
k_transpose(const float* __restrict__ in, int64_t in_sn, int64_t in_sa, float* __restrict__ out, int64_t out_sn,
            int64_t out_sb, int A, int B)
{
    __shared__ float tile[32][33];
    const int64_t img = blockIdx.z;
    const int b0 = blockIdx.x * 32, a0 = blockIdx.y * 32;
    const int tx = threadIdx.x & 31, ty = threadIdx.x >> 5;  // 32 x 8
    const float* src = in + img * in_sn;
    float* dst = out + img * out_sn;
    for (int i = ty; i < 32; i += 8) {
        int a = a0 + i, b = b0 + tx;
        if (a < A && b < B) tile[i][tx] = src[(int64_t)a * in_sa + b];
    }
    __syncthreads();
    for (int i = ty; i < 32; i += 8) {
        int b = b0 + i, a = a0 + tx;
        if (a < A && b < B) dst[(int64_t)b * out_sb + a] = tile[tx][i];
    }
}

int launch_transpose_strided(const float* in, int64_t in_sn, int64_t in_sa, float* out, int64_t out_sn, int64_t out_sb,
                             int n, int A, int B, cudaStream_t st)
{
    FDN_CHECK_ARG(cdiv(A, 32) <= 65535, "transpose: A too large");
    for (int b0 = 0; b0 < n; b0 += 65535) {
        const int nb = n - b0 < 65535 ? n - b0 : 65535;
        dim3 grid((unsigned)cdiv(B, 32), (unsigned)cdiv(A, 32), (unsigned)nb);
        ProfScope ps(K_TRANSPOSE, 8.0 * nb * A * B, st);
        k_transpose<<<grid, 256, 0, st>>>(in + (int64_t)b0 * in_sn, in_sn, in_sa, out + (int64_t)b0 * out_sn, out_sn,
                                          out_sb, A, B);
        FDN_LAUNCHED("k_transpose");
    }
    return FDN_OK;
}

int launch_transpose(const float* in, float* out, int n, int A, int B, cudaStream_t st)
{
    return launch_transpose_strided(in, (int64_t)A * B, B, out, (int64_t)A * B, A, n, A, B, st);
}

// ------------------------------------------------------------------------------------------------
// Strided 3-D copy with periodic index offsets (the packing side of the re-slab / halo exchange):
//   out[a*out_sa + b*out_sb + c] = in[a*in_sa + ((b0 + b) mod bw)*in_sb + ((c0 + c) mod cw)]
// ------------------------------------------------------------------------------------------------
__global__ void __launch_bounds__(256)
k_copy3d(const float* __restrict__ in, int64_t in_sa, int64_t in_sb, int b0, int bw, int c0, int cw,
         float* __restrict__ out, int64_t out_sa, int64_t out_sb, int B, int C)
{
    const int c = blockIdx.x * 256 + threadIdx.x;
    const int b = blockIdx.y;
    const int64_t a = blockIdx.z;
    if (c >= C) return;
    int bi = (b0 + b) % bw;
    if (bi < 0) bi += bw;
    int ci = (c0 + c) % cw;
    if (ci < 0) ci += cw;
    (void)B;
    out[a * out_sa + (int64_t)b * out_sb + c] = in[a * in_sa + (int64_t)bi * in_sb + ci];
}

int launch_copy3d(const float* in, int64_t in_sa, int64_t in_sb, int b0, int bw, int c0, int cw, float* out,
                  int64_t out_sa, int64_t out_sb, int A, int B, int C, cudaStream_t st)
{
    FDN_CHECK_ARG(B <= 65535, "copy3d: B too large");
    FDN_CHECK_ARG(bw >= 1 && cw >= 1, "copy3d: wrap lengths must be >= 1");
    for (int a0 = 0; a0 < A; a0 += 65535) {
        const int na = A - a0 < 65535 ? A - a0 : 65535;
        dim3 grid((unsigned)cdiv(C, 256), (unsigned)B, (unsigned)na);
        ProfScope ps(K_COPY3D, 8.0 * na * B * C, st);
        k_copy3d<<<grid, 256, 0, st>>>(in + (int64_t)a0 * in_sa, in_sa, in_sb, b0, bw, c0, cw,
                                       out + (int64_t)a0 * out_sa, out_sa, out_sb, B, C);
        FDN_LAUNCHED("k_copy3d");
    }
    return FDN_OK;
}

}  // namespace fdn
