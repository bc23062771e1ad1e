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
    *ap = (float)__dadd_rn(first ? 0.0 : (double)*ap, prod);   // tmp = zeros; tmp += v * k (:308, :316)
}

// Four consecutive pixels per thread (128-bit flow / accumulator accesses, four independent gathers in flight).
// Same arithmetic as k_warp_acc. Needs W % 4 == 0 and 16-byte aligned rows.
__device__ __forceinline__ float remap_px(const float* __restrict__ src, int64_t n_rs, float fx, float fy, int x, int y,
                                          int H, int W)
{
    const float mx = __fadd_rn(fx, (float)x), my = __fadd_rn(fy, (float)y);
    const int sx = __float2int_rn(__fmul_rn(mx, 32.f)), sy = __float2int_rn(__fmul_rn(my, 32.f));
    const int ax = sx & 31, ay = sy & 31;
    int ix = sx >> 5, iy = sy >> 5;
    ix = min(max(ix, -32768), 32767);
    iy = min(max(iy, -32768), 32767);
    const float tx1 = __fmul_rn((float)ax, 0.03125f), tx0 = __fsub_rn(1.f, tx1);
    const float ty1 = __fmul_rn((float)ay, 0.03125f), ty0 = __fsub_rn(1.f, ty1);
    const float w0 = __fmul_rn(ty0, tx0), w1 = __fmul_rn(ty0, tx1), w2 = __fmul_rn(ty1, tx0), w3 = __fmul_rn(ty1, tx1);
    const int x0 = min(max(ix, 0), W - 1), x1 = min(max(ix + 1, 0), W - 1);
    const int y0 = min(max(iy, 0), H - 1), y1 = min(max(iy + 1, 0), H - 1);
    const float* r0 = src + (int64_t)y0 * n_rs;
    const float* r1 = src + (int64_t)y1 * n_rs;
    const float v0 = __ldg(r0 + x0), v1 = __ldg(r0 + x1), v2 = __ldg(r1 + x0), v3 = __ldg(r1 + x1);
    return __fadd_rn(__fadd_rn(__fadd_rn(__fmul_rn(v0, w0), __fmul_rn(v1, w1)), __fmul_rn(v2, w2)), __fmul_rn(v3, w3));
}

__global__ void __launch_bounds__(128)
k_warp_acc4(const float* __restrict__ neigh, int64_t n_ss, int64_t n_rs, SlotMap n_map, const float4* __restrict__ flow,
            double weight, float* __restrict__ acc, int64_t a_ss, int64_t a_rs, int H, int W, int first)
{
    const int x = (blockIdx.x * 128 + threadIdx.x) * 4;
    const int y = blockIdx.y;
    const int b = blockIdx.z;
    if (x >= W) return;
    const float* src = neigh + (int64_t)n_map.slot(b) * n_ss;
    float v[4];
    if (flow) {
        const float4* fp = flow + (((int64_t)b * H + y) * W + x) / 2;   // two float2 flows per float4
        const float4 f01 = __ldg(fp), f23 = __ldg(fp + 1);
        v[0] = remap_px(src, n_rs, f01.x, f01.y, x, y, H, W);
        v[1] = remap_px(src, n_rs, f01.z, f01.w, x + 1, y, H, W);
        v[2] = remap_px(src, n_rs, f23.x, f23.y, x + 2, y, H, W);
        v[3] = remap_px(src, n_rs, f23.z, f23.w, x + 3, y, H, W);
    } else {
        const float4 c = __ldg(reinterpret_cast<const float4*>(src + (int64_t)y * n_rs + x));
        v[0] = c.x; v[1] = c.y; v[2] = c.z; v[3] = c.w;
    }
    float4* ap = reinterpret_cast<float4*>(acc + (int64_t)b * a_ss + (int64_t)y * a_rs + x);
    float o[4];
    if (first) {
#pragma unroll
        for (int i = 0; i < 4; i++) o[i] = (float)__dadd_rn(0.0, __dmul_rn((double)v[i], weight));   // tmp = zeros; tmp += v * k
    } else {
        const float4 a = *ap;
        const float av[4] = {a.x, a.y, a.z, a.w};
#pragma unroll
        for (int i = 0; i < 4; i++) o[i] = (float)__dadd_rn((double)av[i], __dmul_rn((double)v[i], weight));
    }
    *ap = make_float4(o[0], o[1], o[2], o[3]);
}

int launch_warp_acc(const float* neigh, int64_t n_ss, int64_t n_rs, SlotMap n_map, const float* flow, double weight,
                    float* acc, int64_t a_ss, int64_t a_rs, int n, int H, int W, int first, cudaStream_t st)
{
    const bool vec4 = W % 4 == 0 && n_ss % 4 == 0 && n_rs % 4 == 0 && a_ss % 4 == 0 && a_rs % 4 == 0 &&
                      (reinterpret_cast<uintptr_t>(neigh) & 15) == 0 && (reinterpret_cast<uintptr_t>(acc) & 15) == 0 &&
                      (reinterpret_cast<uintptr_t>(flow) & 15) == 0;
    if (vec4) {
        for (int b0 = 0; b0 < n; b0 += 65535) {
            const int nb = n - b0 < 65535 ? n - b0 : 65535;
            SlotMap m = n_map;
            m.base += b0;
            dim3 grid((unsigned)cdiv(W, 512), (unsigned)H, (unsigned)nb);
            ProfScope ps(K_WARP_ACC, (double)nb * H * W * (4.0 + (flow ? 8.0 : 0.0) + (first ? 4.0 : 8.0)), st);
            k_warp_acc4<<<grid, 128, 0, st>>>(neigh, n_ss, n_rs, m,
                                              flow ? reinterpret_cast<const float4*>(flow + (int64_t)b0 * H * W * 2) : nullptr,
                                              weight, acc + (int64_t)b0 * a_ss, a_ss, a_rs, H, W, first);
            FDN_LAUNCHED("k_warp_acc4");
        }
        return FDN_OK;
    }
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
// Both chain directions in one launch (the two chains of src/flowdenoising.py:311-316 and :319-324 advance together):
// images [0, n) are the backward neighbours -- remapped and accumulated in place, the reference's order -- and images
// [n, 2n) the forward neighbours, whose remapped values wait in `stash` until the backward chain and the centre tap
// are in the accumulator (k_acc_finish applies them in the reference's order: centre, +1, ..., +r).
// ------------------------------------------------------------------------------------------------
template <int VEC>
__global__ void __launch_bounds__(128)
k_warp_pair(const float* __restrict__ neigh, int64_t n_ss, int64_t n_rs, SlotMap n_map, const float* __restrict__ flow,
            double weight, float* __restrict__ acc, int64_t a_ss, int64_t a_rs, float* __restrict__ stash, int n, int b0,
            int H, int W, int first)
{
    const int x = (blockIdx.x * 128 + threadIdx.x) * VEC;
    const int y = blockIdx.y;
    const int b = b0 + blockIdx.z;   // 0 .. 2n-1
    if (x >= W) return;
    const float* src = neigh + (int64_t)n_map.slot(b) * n_ss;
    const int64_t fpx = ((int64_t)b * H + y) * W + x;
    float v[VEC];
    if (VEC == 4) {
        const float4* fp = reinterpret_cast<const float4*>(flow) + fpx / 2;   // two float2 flows per float4
        const float4 f01 = __ldg(fp), f23 = __ldg(fp + 1);
        v[0] = remap_px(src, n_rs, f01.x, f01.y, x, y, H, W);
        v[1 % VEC] = remap_px(src, n_rs, f01.z, f01.w, x + 1, y, H, W);
        v[2 % VEC] = remap_px(src, n_rs, f23.x, f23.y, x + 2, y, H, W);
        v[3 % VEC] = remap_px(src, n_rs, f23.z, f23.w, x + 3, y, H, W);
    } else {
        const float2 f = __ldg(reinterpret_cast<const float2*>(flow) + fpx);
        v[0] = remap_px(src, n_rs, f.x, f.y, x, y, H, W);
    }
    if (b < n) {
        float* ap = acc + (int64_t)b * a_ss + (int64_t)y * a_rs + x;
        if (VEC == 4) {
            float4 a = make_float4(0.f, 0.f, 0.f, 0.f);
            if (!first) a = *reinterpret_cast<float4*>(ap);
            const float av[4] = {a.x, a.y, a.z, a.w};
            float o[4];
#pragma unroll
            for (int i = 0; i < 4; i++)
                o[i] = (float)__dadd_rn(first ? 0.0 : (double)av[i], __dmul_rn((double)v[i % VEC], weight));
            *reinterpret_cast<float4*>(ap) = make_float4(o[0], o[1], o[2], o[3]);
        } else {
            *ap = (float)__dadd_rn(first ? 0.0 : (double)*ap, __dmul_rn((double)v[0], weight));
        }
    } else {
        float* sp = stash + ((int64_t)(b - n) * H + y) * W + x;
        if (VEC == 4) *reinterpret_cast<float4*>(sp) = make_float4(v[0], v[1 % VEC], v[2 % VEC], v[3 % VEC]);
        else *sp = v[0];
    }
}

int launch_warp_pair(const float* neigh, int64_t n_ss, int64_t n_rs, SlotMap n_map, const float* flow, double weight,
                     float* acc, int64_t a_ss, int64_t a_rs, float* stash, int n, int H, int W, int first,
                     cudaStream_t st)
{
    const bool vec4 = W % 4 == 0 && n_ss % 4 == 0 && n_rs % 4 == 0 && a_ss % 4 == 0 && a_rs % 4 == 0 &&
                      (reinterpret_cast<uintptr_t>(neigh) & 15) == 0 && (reinterpret_cast<uintptr_t>(acc) & 15) == 0 &&
                      (reinterpret_cast<uintptr_t>(flow) & 15) == 0 && (reinterpret_cast<uintptr_t>(stash) & 15) == 0;
    FDN_CHECK_ARG(H <= 65535, "image too tall for one launch");
    for (int b0 = 0; b0 < 2 * n; b0 += 65535) {
        const int nb = 2 * n - b0 < 65535 ? 2 * n - b0 : 65535;
        // remap: flow 8 + neighbour 4; backward half: accumulator 4 (+4 unless first), forward half: stash 4
        ProfScope ps(K_WARP_ACC, (double)nb * H * W * (12.0 + (first ? 4.0 : 6.0)), st);
        if (vec4) {
            dim3 grid((unsigned)cdiv(W, 512), (unsigned)H, (unsigned)nb);
            k_warp_pair<4><<<grid, 128, 0, st>>>(neigh, n_ss, n_rs, n_map, flow, weight, acc, a_ss, a_rs, stash, n, b0, H, W,
                                                 first);
        } else {
            dim3 grid((unsigned)cdiv(W, 128), (unsigned)H, (unsigned)nb);
            k_warp_pair<1><<<grid, 128, 0, st>>>(neigh, n_ss, n_rs, n_map, flow, weight, acc, a_ss, a_rs, stash, n, b0, H, W,
                                                 first);
        }
        FDN_LAUNCHED("k_warp_pair");
    }
    return FDN_OK;
}

#define FDN_MAX_R 128
struct FinishTaps {
    double k[FDN_MAX_R + 1];   // centre, +1, ..., +r
};

template <int VEC>
__global__ void __launch_bounds__(128)
k_acc_finish(const float* __restrict__ centre, int64_t c_ss, int64_t c_rs, SlotMap c_map, const float* __restrict__ stash,
             int64_t stash_stride, int r, FinishTaps taps, float* __restrict__ acc, int64_t a_ss, int64_t a_rs, int b0,
             int H, int W, int have_acc)
{
    const int x = (blockIdx.x * 128 + threadIdx.x) * VEC;
    const int y = blockIdx.y;
    const int b = b0 + blockIdx.z;
    if (x >= W) return;
    const float* cp = centre + (int64_t)c_map.slot(b) * c_ss + (int64_t)y * c_rs + x;
    float* ap = acc + (int64_t)b * a_ss + (int64_t)y * a_rs + x;
    const float* sp = stash + ((int64_t)b * H + y) * W + x;
    float a[VEC], v[VEC];
    auto load = [&](const float* p, float* dst) {
        if (VEC == 4) {
            const float4 q = __ldg(reinterpret_cast<const float4*>(p));
            dst[0] = q.x; dst[1 % VEC] = q.y; dst[2 % VEC] = q.z; dst[3 % VEC] = q.w;
        } else {
            dst[0] = __ldg(p);
        }
    };
#pragma unroll
    for (int i = 0; i < VEC; i++) a[i] = 0.f;
    if (have_acc) {
        if (VEC == 4) {
            const float4 q = *reinterpret_cast<const float4*>(ap);
            a[0] = q.x; a[1 % VEC] = q.y; a[2 % VEC] = q.z; a[3 % VEC] = q.w;
        } else {
            a[0] = *ap;
        }
    }
    load(cp, v);   // tmp_slice += vol[z] * kernel[ks2]  (src/flowdenoising.py:317)
#pragma unroll
    for (int i = 0; i < VEC; i++) a[i] = (float)__dadd_rn((double)a[i], __dmul_rn((double)v[i], taps.k[0]));
    for (int d = 0; d < r; d++) {   // forward chain, nearest neighbour first (:319-324)
        load(sp + (int64_t)d * stash_stride, v);
#pragma unroll
        for (int i = 0; i < VEC; i++) a[i] = (float)__dadd_rn((double)a[i], __dmul_rn((double)v[i], taps.k[d + 1]));
    }
    if (VEC == 4) *reinterpret_cast<float4*>(ap) = make_float4(a[0], a[1 % VEC], a[2 % VEC], a[3 % VEC]);
    else *ap = a[0];
}

int launch_acc_finish(const float* centre, int64_t c_ss, int64_t c_rs, SlotMap c_map, const float* stash, int r,
                      const double* k, float* acc, int64_t a_ss, int64_t a_rs, int n, int H, int W, int have_acc,
                      cudaStream_t st)
{
    FDN_CHECK_ARG(r >= 0 && r <= FDN_MAX_R, "kernel radius %d unsupported (<= %d)", r, FDN_MAX_R);
    FDN_CHECK_ARG(H <= 65535, "image too tall for one launch");
    FinishTaps taps;
    for (int i = 0; i <= r; i++) taps.k[i] = k[i];
    const bool vec4 = W % 4 == 0 && c_ss % 4 == 0 && c_rs % 4 == 0 && a_ss % 4 == 0 && a_rs % 4 == 0 &&
                      (reinterpret_cast<uintptr_t>(centre) & 15) == 0 && (reinterpret_cast<uintptr_t>(acc) & 15) == 0 &&
                      (reinterpret_cast<uintptr_t>(stash) & 15) == 0;
    const int64_t stash_stride = (int64_t)n * H * W;
    for (int b0 = 0; b0 < n; b0 += 65535) {
        const int nb = n - b0 < 65535 ? n - b0 : 65535;
        ProfScope ps(K_WARP_ACC, (double)nb * H * W * (4.0 * (r + 2) + (have_acc ? 4.0 : 0.0)), st);
        if (vec4) {
            dim3 grid((unsigned)cdiv(W, 512), (unsigned)H, (unsigned)nb);
            k_acc_finish<4><<<grid, 128, 0, st>>>(centre, c_ss, c_rs, c_map, stash, stash_stride, r, taps, acc, a_ss, a_rs, b0,
                                                  H, W, have_acc);
        } else {
            dim3 grid((unsigned)cdiv(W, 128), (unsigned)H, (unsigned)nb);
            k_acc_finish<1><<<grid, 128, 0, st>>>(centre, c_ss, c_rs, c_map, stash, stash_stride, r, taps, acc, a_ss, a_rs, b0,
                                                  H, W, have_acc);
        }
        FDN_LAUNCHED("k_acc_finish");
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
