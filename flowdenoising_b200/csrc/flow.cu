// Stage 3 kernels: displacement update (FarnebackUpdateMatrices) fused with the flow blur + 2x2 solve
// (FarnebackUpdateFlow_Blur), and the flow resampling between pyramid levels. sm_100a.
//
// Reference behaviour: the inner loop of cv2.calcOpticalFlowFarneback as called from
// /root/reference/src/flowdenoising.py:69-79.
//
// k_flow_iter design (one launch = one Farneback iteration of one level for a batch of image pairs):
//   * a block owns a strip of CW columns of one image pair and MARCHES down the rows, exactly like OpenCV's
//     sliding vertical sum: vsum(y) = vsum(y-1) + float32(M[y+m] - M[y-m-1]) in float64. The float32
//     subtraction makes OpenCV's box sums history dependent, so a tiled box filter cannot reproduce them
//     bit for bit; the march does, and it needs no vertical halo (every R0/flow row is read once).
//   * the five M channels of a pixel are never written to HBM: each thread computes M for its own column from
//     R0, flow and a bilinear gather of R1, keeps the last 2m+2 rows in a shared-memory ring, updates its
//     float64 column sums, publishes them to shared memory, and the strip's core threads sum 2m+1 neighbours
//     (replicate border = clamped column) and solve.
//   Algorithmic HBM bytes per pixel: R0 20 + R1 20 + flow in 8 + flow out 8 = 56.
// Compiled with -fmad=false: OpenCV's scalar code has no fused multiply-adds here.
#include "fdn_internal.cuh"

namespace fdn {

struct FlowIterArgs {
    const float* R;
    int64_t R_stride;
    SlotMap map0, map1;
    const float* flow_in;
    float* flow_out;
    int h, w, m;
    double scale;  // 1 / winsize^2
};

__device__ __forceinline__ void update_matrices_px(const float* __restrict__ R0, const float* __restrict__ R1,
                                                   const float2* __restrict__ flow, int x, int y, int h, int w,
                                                   float M[5])
{
    const float2 f = flow[(int64_t)y * w + x];
    const float dx = f.x, dy = f.y;
    float fx = __fadd_rn((float)x, dx), fy = __fadd_rn((float)y, dy);
    const int x1 = (int)floorf(fx), y1 = (int)floorf(fy);
    fx = __fsub_rn(fx, (float)x1);
    fy = __fsub_rn(fy, (float)y1);
    const float* r0p = R0 + (int64_t)y * 5 * w + x;
    const float c0 = r0p[0], c1 = r0p[w], c2 = r0p[2 * w], c3 = r0p[3 * w], c4 = r0p[4 * w];
    float r2, r3, r4, r5, r6;
    if ((unsigned)x1 < (unsigned)(w - 1) && (unsigned)y1 < (unsigned)(h - 1)) {
        const float ofx = __fsub_rn(1.f, fx), ofy = __fsub_rn(1.f, fy);
        const float a00 = __fmul_rn(ofx, ofy), a01 = __fmul_rn(fx, ofy), a10 = __fmul_rn(ofx, fy),
                    a11 = __fmul_rn(fx, fy);
        const float* q0 = R1 + (int64_t)y1 * 5 * w + x1;
        const float* q1 = q0 + 5 * w;
#define FDN_BILIN(c)                                                                                         \
    __fadd_rn(__fadd_rn(__fadd_rn(__fmul_rn(a00, q0[(c) * w]), __fmul_rn(a01, q0[(c) * w + 1])),             \
                        __fmul_rn(a10, q1[(c) * w])),                                                        \
              __fmul_rn(a11, q1[(c) * w + 1]))
        r2 = FDN_BILIN(0);
        r3 = FDN_BILIN(1);
        r4 = FDN_BILIN(2);
        r5 = FDN_BILIN(3);
        r6 = FDN_BILIN(4);
#undef FDN_BILIN
        r4 = __fmul_rn(__fadd_rn(c2, r4), 0.5f);
        r5 = __fmul_rn(__fadd_rn(c3, r5), 0.5f);
        r6 = __fmul_rn(__fadd_rn(c4, r6), 0.25f);
    } else {
        r2 = r3 = 0.f;
        r4 = c2;
        r5 = c3;
        r6 = __fmul_rn(c4, 0.5f);
    }
    r2 = __fmul_rn(__fsub_rn(c0, r2), 0.5f);
    r3 = __fmul_rn(__fsub_rn(c1, r3), 0.5f);
    r2 = __fadd_rn(r2, __fadd_rn(__fmul_rn(r4, dy), __fmul_rn(r6, dx)));
    r3 = __fadd_rn(r3, __fadd_rn(__fmul_rn(r6, dy), __fmul_rn(r5, dx)));
    if ((unsigned)(x - 5) >= (unsigned)(w - 10) || (unsigned)(y - 5) >= (unsigned)(h - 10)) {
        // border[5] = {0.14, 0.14, 0.4472, 0.4472, 0.4472}
        float s = x < 5 ? (x < 2 ? 0.14f : 0.4472f) : 1.f;
        int xr = w - x - 1, yr = h - y - 1;
        s = __fmul_rn(s, x >= w - 5 ? (xr < 2 ? 0.14f : 0.4472f) : 1.f);
        s = __fmul_rn(s, y < 5 ? (y < 2 ? 0.14f : 0.4472f) : 1.f);
        s = __fmul_rn(s, y >= h - 5 ? (yr < 2 ? 0.14f : 0.4472f) : 1.f);
        r2 = __fmul_rn(r2, s); r3 = __fmul_rn(r3, s); r4 = __fmul_rn(r4, s);
        r5 = __fmul_rn(r5, s); r6 = __fmul_rn(r6, s);
    }
    M[0] = __fadd_rn(__fmul_rn(r4, r4), __fmul_rn(r6, r6));
    M[1] = __fmul_rn(__fadd_rn(r4, r5), r6);
    M[2] = __fadd_rn(__fmul_rn(r5, r5), __fmul_rn(r6, r6));
    M[3] = __fadd_rn(__fmul_rn(r4, r2), __fmul_rn(r6, r3));
    M[4] = __fadd_rn(__fmul_rn(r6, r2), __fmul_rn(r5, r3));
}

template <int CW>
__global__ void __launch_bounds__(CW + 32)
k_flow_iter(FlowIterArgs a)
{
    constexpr int NT = CW + 32;
    extern __shared__ __align__(16) unsigned char smem_raw[];
    const int m = a.m, h = a.h, w = a.w;
    const int RR = 2 * m + 2;
    double* vbuf = reinterpret_cast<double*>(smem_raw);                              // [2][5][NT]
    float* ring = reinterpret_cast<float*>(smem_raw + sizeof(double) * 2 * 5 * NT);  // [RR][5][NT]

    const int t = threadIdx.x;
    const int b = blockIdx.y;
    const int x0 = blockIdx.x * CW;
    // position p in the strip's extended window [x0 - m, x0 + CW + m)
    int p;
    bool active = true;
    if (t < CW) {
        p = t + m;
    } else {
        const int hd = t - CW;
        if (hd < m) p = hd;
        else if (hd < 2 * m) p = CW + hd;
        else { p = 0; active = false; }
    }
    const int xcl = min(max(x0 - m + p, 0), w - 1);
    const bool core = (t < CW) && (x0 + t < w);

    const float* R0 = a.R + (int64_t)a.map0.slot(b) * a.R_stride;
    const float* R1 = a.R + (int64_t)a.map1.slot(b) * a.R_stride;
    const float2* fin = reinterpret_cast<const float2*>(a.flow_in) + (int64_t)b * h * w;
    float2* fout = reinterpret_cast<float2*>(a.flow_out) + (int64_t)b * h * w;

    double vs[5] = {0, 0, 0, 0, 0};
    int next_row = 0;   // next M row to compute
    int next_slot = 0;  // its ring slot (= next_row mod RR)
    float Mv[5];

    if (active) {
        // rows 0 .. m-1 (clamped to h-1) seed the column sums: vsum = M[0]*(m+2) + sum_{y=1}^{m-1} M[min(y,h-1)]
        const int last_init = min(max(m - 1, 0), h - 1);
        for (; next_row <= last_init; next_row++) {
            update_matrices_px(R0, R1, fin, xcl, next_row, h, w, Mv);
#pragma unroll
            for (int c = 0; c < 5; c++) ring[(next_slot * 5 + c) * NT + t] = Mv[c];
            next_slot = next_slot + 1 == RR ? 0 : next_slot + 1;
        }
        const float mp2 = (float)(m + 2);
#pragma unroll
        for (int c = 0; c < 5; c++) vs[c] = (double)__fmul_rn(ring[c * NT + t], mp2);
        for (int y = 1; y < m; y++) {
            const int sl = min(y, h - 1);  // < RR always
#pragma unroll
            for (int c = 0; c < 5; c++) vs[c] = __dadd_rn(vs[c], (double)ring[(sl * 5 + c) * NT + t]);
        }
    }

    for (int y = 0; y < h; y++) {
        double* vb = vbuf + (y & 1) * 5 * NT;
        if (active) {
            const int j1 = min(y + m, h - 1);
            if (next_row <= j1) {  // exactly one new row per step while y + m < h
                update_matrices_px(R0, R1, fin, xcl, next_row, h, w, Mv);
#pragma unroll
                for (int c = 0; c < 5; c++) ring[(next_slot * 5 + c) * NT + t] = Mv[c];
                next_slot = next_slot + 1 == RR ? 0 : next_slot + 1;
                next_row++;
            }
            const int j0 = max(y - m - 1, 0);
            const int s1 = j1 % RR, s0 = j0 % RR;
#pragma unroll
            for (int c = 0; c < 5; c++) {
                const float d = __fsub_rn(ring[(s1 * 5 + c) * NT + t], ring[(s0 * 5 + c) * NT + t]);
                vs[c] = __dadd_rn(vs[c], (double)d);
                vb[c * NT + p] = vs[c];
            }
        }
        __syncthreads();
        if (core) {
            double g[5];
#pragma unroll
            for (int c = 0; c < 5; c++) {
                const double* q = vb + c * NT + t;  // positions t .. t + 2m  <->  columns x-m .. x+m
                double s = q[0];
                for (int i = 1; i <= 2 * m; i++) s = __dadd_rn(s, q[i]);
                g[c] = __dmul_rn(s, a.scale);
            }
            const double det = __dadd_rn(__dsub_rn(__dmul_rn(g[0], g[2]), __dmul_rn(g[1], g[1])), 1e-3);
            const double idet = __ddiv_rn(1., det);
            float2 o;
            o.x = (float)__dmul_rn(__dsub_rn(__dmul_rn(g[0], g[4]), __dmul_rn(g[1], g[3])), idet);
            o.y = (float)__dmul_rn(__dsub_rn(__dmul_rn(g[2], g[3]), __dmul_rn(g[1], g[4])), idet);
            fout[(int64_t)y * w + x0 + t] = o;
        }
    }
}

int launch_flow_iter(const float* R, int64_t R_stride, SlotMap map0, SlotMap map1, const float* flow_in,
                     float* flow_out, int n, int h, int w, int winsize, cudaStream_t st)
{
    FDN_CHECK_ARG(winsize >= 1 && winsize <= 33, "winsize %d unsupported (1..33)", winsize);
    FDN_CHECK_ARG(flow_in != flow_out, "flow_in and flow_out must not alias");
    FlowIterArgs a;
    a.R = R; a.R_stride = R_stride; a.flow_in = flow_in; a.flow_out = flow_out;
    a.h = h; a.w = w; a.m = winsize / 2;
    a.scale = 1. / ((double)winsize * winsize);
    const int RR = 2 * a.m + 2;
    const bool wide = w > 96;
    const int NT = (wide ? 128 : 32) + 32;
    const size_t smem = sizeof(double) * 2 * 5 * NT + sizeof(float) * RR * 5 * NT;
    static bool attr_set = false;
    if (!attr_set) {
        FDN_CUDA(cudaFuncSetAttribute(k_flow_iter<128>, cudaFuncAttributeMaxDynamicSharedMemorySize, 160 * 1024));
        FDN_CUDA(cudaFuncSetAttribute(k_flow_iter<32>, cudaFuncAttributeMaxDynamicSharedMemorySize, 160 * 1024));
        attr_set = true;
    }
    for (int b0 = 0; b0 < n; b0 += 65535) {
        const int nb = n - b0 < 65535 ? n - b0 : 65535;
        a.map0 = map0; a.map0.base += b0;
        a.map1 = map1; a.map1.base += b0;
        a.flow_in = flow_in + (int64_t)b0 * h * w * 2;
        a.flow_out = flow_out + (int64_t)b0 * h * w * 2;
        ProfScope ps(K_FLOW_ITER, 56.0 * nb * h * w, st);
        if (wide) {
            dim3 grid((unsigned)cdiv(w, 128), (unsigned)nb);
            k_flow_iter<128><<<grid, 160, smem, st>>>(a);
        } else {
            dim3 grid((unsigned)cdiv(w, 32), (unsigned)nb);
            k_flow_iter<32><<<grid, 64, smem, st>>>(a);
        }
        FDN_LAUNCHED("k_flow_iter");
    }
    return FDN_OK;
}

// ------------------------------------------------------------------------------------------------
// Initial flow of the coarsest level: resize(flow0, INTER_AREA) * scale  
// ------------------------------------------------------------------------------------------------
struct AreaAxis {
    int S, D, iscale, fast;
    double scale;
};

// fractional-coverage taps of destination index d (OpenCV computeResizeAreaTab): [first, last] source range with
// alpha weights; returns count and fills si/alpha (at most 3 + interior entries; interior alpha = 1/cellWidth)
__device__ __forceinline__ void area_entry_range(const AreaAxis& ax, int d, int& sx1, int& sx2, float& a_first,
                                                 float& a_mid, float& a_last, bool& has_first, bool& has_last)
{
    const double fsx1 = __dmul_rn((double)d, ax.scale);
    const double fsx2 = __dadd_rn(fsx1, ax.scale);
    const double cell = fmin(ax.scale, __dsub_rn((double)ax.S, fsx1));
    sx1 = (int)ceil(fsx1);
    sx2 = (int)floor(fsx2);
    sx2 = min(sx2, ax.S - 1);
    sx1 = min(sx1, sx2);
    has_first = __dsub_rn((double)sx1, fsx1) > 1e-3;
    a_first = (float)__ddiv_rn(__dsub_rn((double)sx1, fsx1), cell);
    a_mid = (float)__ddiv_rn(1.0, cell);
    has_last = __dsub_rn(fsx2, (double)sx2) > 1e-3;
    a_last = (float)__ddiv_rn(fmin(fmin(__dsub_rn(fsx2, (double)sx2), 1.), cell), cell);
}

__global__ void __launch_bounds__(128)
k_flow_area_down(const float2* __restrict__ in, int H, int W, float2* __restrict__ out, int h, int w, AreaAxis ax,
                 AreaAxis ay, float scale)
{
    const int dx = blockIdx.x * 128 + threadIdx.x;
    const int dy = blockIdx.y;
    const int b = blockIdx.z;
    if (dx >= w) return;
    const float2* src = in + (int64_t)b * H * W;
    float rx, ry;
    if (ax.fast && ay.fast) {
        // resizeAreaFast_: block sum in row-major order, unrolled by four, times 1/area (float32)
        const int area = ax.iscale * ay.iscale;
        float sx_ = 0.f, sy_ = 0.f;
        int k = 0;
        float vx[4], vy[4];
        for (int sy = 0; sy < ay.iscale; sy++) {
            const float2* row = src + (int64_t)(dy * ay.iscale + sy) * W + dx * ax.iscale;
            for (int sx = 0; sx < ax.iscale; sx++) {
                const float2 v = row[sx];
                if (k < (area & ~3)) {
                    vx[k & 3] = v.x; vy[k & 3] = v.y;
                    if ((k & 3) == 3) {
                        sx_ = __fadd_rn(sx_, __fadd_rn(__fadd_rn(__fadd_rn(vx[0], vx[1]), vx[2]), vx[3]));
                        sy_ = __fadd_rn(sy_, __fadd_rn(__fadd_rn(__fadd_rn(vy[0], vy[1]), vy[2]), vy[3]));
                    }
                } else {
                    sx_ = __fadd_rn(sx_, v.x);
                    sy_ = __fadd_rn(sy_, v.y);
                }
                k++;
            }
        }
        const float inv = __fdiv_rn(1.f, (float)area);
        rx = __fmul_rn(sx_, inv);
        ry = __fmul_rn(sy_, inv);
    } else {
        // resizeArea_: per source row, horizontal weighted sum in table order; rows combined with beta weights
        int x1, x2, y1, y2;
        float axf, axm, axl, ayf, aym, ayl;
        bool hxf, hxl, hyf, hyl;
        area_entry_range(ax, dx, x1, x2, axf, axm, axl, hxf, hxl);
        area_entry_range(ay, dy, y1, y2, ayf, aym, ayl, hyf, hyl);
        float sumx = 0.f, sumy = 0.f;
        bool firstrow = true;
        const int ya = hyf ? y1 - 1 : y1, yb = hyl ? y2 : y2 - 1;
        for (int sy = ya; sy <= yb; sy++) {
            const float beta = (sy == y1 - 1) ? ayf : ((sy == y2 && hyl) ? ayl : aym);
            const float2* row = src + (int64_t)sy * W;
            float bx = 0.f, by = 0.f;
            if (hxf) {
                const float2 v = row[x1 - 1];
                bx = __fadd_rn(bx, __fmul_rn(v.x, axf)); by = __fadd_rn(by, __fmul_rn(v.y, axf));
            }
            for (int sx = x1; sx < x2; sx++) {
                const float2 v = row[sx];
                bx = __fadd_rn(bx, __fmul_rn(v.x, axm)); by = __fadd_rn(by, __fmul_rn(v.y, axm));
            }
            if (hxl) {
                const float2 v = row[x2];
                bx = __fadd_rn(bx, __fmul_rn(v.x, axl)); by = __fadd_rn(by, __fmul_rn(v.y, axl));
            }
            if (firstrow) {
                sumx = __fmul_rn(beta, bx); sumy = __fmul_rn(beta, by);
                firstrow = false;
            } else {
                sumx = __fadd_rn(sumx, __fmul_rn(beta, bx)); sumy = __fadd_rn(sumy, __fmul_rn(beta, by));
            }
        }
        rx = sumx; ry = sumy;
    }
    float2 o;
    o.x = __fmul_rn(rx, scale);
    o.y = __fmul_rn(ry, scale);
    out[((int64_t)b * h + dy) * w + dx] = o;
}

static AreaAxis make_area_axis(int S, int D)
{
    AreaAxis a;
    a.S = S; a.D = D;
    a.scale = 1. / ((double)D / S);
    a.iscale = (int)lrint(a.scale);
    a.fast = fabs(a.scale - a.iscale) < 2.220446049250313e-16;
    return a;
}

int launch_flow_area_down(const float* flow, int n, int H, int W, float* out, int h, int w, float scale,
                          cudaStream_t st)
{
    AreaAxis ax = make_area_axis(W, w), ay = make_area_axis(H, h);
    FDN_CHECK_ARG(h <= H && w <= W, "area resize only shrinks");
    for (int b0 = 0; b0 < n; b0 += 65535) {
        const int nb = n - b0 < 65535 ? n - b0 : 65535;
        dim3 grid((unsigned)cdiv(w, 128), (unsigned)h, (unsigned)nb);
        ProfScope ps(K_FLOW_AREA, 8.0 * nb * ((double)H * W + (double)h * w), st);
        k_flow_area_down<<<grid, 128, 0, st>>>(reinterpret_cast<const float2*>(flow) + (int64_t)b0 * H * W, H, W,
                                               reinterpret_cast<float2*>(out) + (int64_t)b0 * h * w, h, w, ax, ay,
                                               scale);
        FDN_LAUNCHED("k_flow_area_down");
    }
    return FDN_OK;
}

// ------------------------------------------------------------------------------------------------
// Flow to the next finer level: resize(prevFlow, INTER_LINEAR) * 2  (OpenCV's own 2-channel code path:
// frac = float(coord) - floor, out = a*(1-f) + b*f, fractions zeroed at the borders horizontally only)
// ------------------------------------------------------------------------------------------------
__global__ void __launch_bounds__(128)
k_flow_upsample(const float2* __restrict__ in, int hin, int win, float2* __restrict__ out, int h, int w,
                double scale_x, double scale_y)
{
    const int dx = blockIdx.x * 128 + threadIdx.x;
    const int dy = blockIdx.y;
    const int b = blockIdx.z;
    if (dx >= w) return;
    float fx = (float)__dsub_rn(__dmul_rn((double)dx + 0.5, scale_x), 0.5);
    int sx = (int)floorf(fx);
    fx = __fsub_rn(fx, (float)sx);
    if (sx < 0) { fx = 0.f; sx = 0; }
    if (sx >= win - 1) { fx = 0.f; sx = win - 1; }
    float fy = (float)__dsub_rn(__dmul_rn((double)dy + 0.5, scale_y), 0.5);
    const int sy = (int)floorf(fy);
    fy = __fsub_rn(fy, (float)sy);
    const int sy0 = min(max(sy, 0), hin - 1), sy1 = min(max(sy + 1, 0), hin - 1);
    const int sx1 = min(sx + 1, win - 1);
    const float2* S0 = in + ((int64_t)b * hin + sy0) * win;
    const float2* S1 = in + ((int64_t)b * hin + sy1) * win;
    const float a1 = fx, a0 = __fsub_rn(1.f, fx), b1 = fy, b0 = __fsub_rn(1.f, fy);
    const float2 p00 = S0[sx], p01 = S0[sx1], p10 = S1[sx], p11 = S1[sx1];
    const float r0x = __fadd_rn(__fmul_rn(p00.x, a0), __fmul_rn(p01.x, a1));
    const float r0y = __fadd_rn(__fmul_rn(p00.y, a0), __fmul_rn(p01.y, a1));
    const float r1x = __fadd_rn(__fmul_rn(p10.x, a0), __fmul_rn(p11.x, a1));
    const float r1y = __fadd_rn(__fmul_rn(p10.y, a0), __fmul_rn(p11.y, a1));
    float2 o;
    o.x = __fmul_rn(__fadd_rn(__fmul_rn(r0x, b0), __fmul_rn(r1x, b1)), 2.f);
    o.y = __fmul_rn(__fadd_rn(__fmul_rn(r0y, b0), __fmul_rn(r1y, b1)), 2.f);
    out[((int64_t)b * h + dy) * w + dx] = o;
}

int launch_flow_upsample(const float* flow, int n, int hin, int win, float* out, int h, int w, cudaStream_t st)
{
    const double scale_x = 1. / ((double)w / win), scale_y = 1. / ((double)h / hin);
    for (int b0 = 0; b0 < n; b0 += 65535) {
        const int nb = n - b0 < 65535 ? n - b0 : 65535;
        dim3 grid((unsigned)cdiv(w, 128), (unsigned)h, (unsigned)nb);
        ProfScope ps(K_FLOW_UP, 8.0 * nb * ((double)hin * win + (double)h * w), st);
        k_flow_upsample<<<grid, 128, 0, st>>>(reinterpret_cast<const float2*>(flow) + (int64_t)b0 * hin * win, hin,
                                              win, reinterpret_cast<float2*>(out) + (int64_t)b0 * h * w, h, w,
                                              scale_x, scale_y);
        FDN_LAUNCHED("k_flow_upsample");
    }
    return FDN_OK;
}

}  // namespace fdn
