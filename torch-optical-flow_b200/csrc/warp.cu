// warp.cu -- K1: flow-based backward warp (bilinear / nearest) + in-kernel validity mask.
//
// Replaces optical_flow.warp + warp_grid (reference optical_flow/operator/operator.py:8-56):
//   gx = linspace(-1,1,W)[j] + flow[b,0,i,j] ; gy = linspace(-1,1,H)[i] + flow[b,1,i,j]
//   out[b,c,i,j] = grid_sample(frame[b,c], (gx,gy), mode, padding_mode, align_corners)
// Neither the base grid nor grid+flow is ever materialised.
//
// Bilinear NCHW kernels (variant argument of ofb_warp_f32; all produce the same bits):
//   * rows (default): lane <-> column, a thread owns 2 consecutive rows of its column, coordinates / weights /
//     tap offsets computed once per pixel and reused per channel; interior pixels take a predicate-free channel
//     loop with every gather in flight before the first FMA.  32 registers -> 64 resident warps per SM: the
//     kernel is bound by latency and instruction issue, not by HBM.
//   * direct: one thread per output pixel, grid-stride; also serves nearest mode and NHWC frames.
//   * staged: one CTA per 64x16 output tile, bounding box of the taps copied to shared memory with cp.async.
//   * TMA window (warp_tma.cu): persistent CTAs, the tile's neighbourhood arrives as 3-D TMA boxes.
// HBM roofline: 4*(C + 2 + C) bytes per pixel (+1 for the u8 mask).
#include <cstdlib>

#include "common.cuh"

namespace {

using namespace ofb;

struct SrcPos {
    float ix, iy;
    bool valid;
};

// flow multipliers: 1 for a normalised flow (the reference's contract); 2/(W-1), 2/(H-1) fuse
// optical_flow.normalize (operator.py:117-130) into the warp -- one separately rounded multiply, as there
struct FlowMul {
    float x, y;
    float step_x, step_y;   // linspace(-1, 1, W / H) steps, divided once on the host (same IEEE fp32 division)
};

template <int PAD, bool AC>
__device__ __forceinline__ SrcPos source_from_flow(float fx, float fy, int i, int j, int H, int W, float step_x,
                                                   float step_y, FlowMul fm) {
    const float gx = __fadd_rn(linspace_m1_p1(j, W, step_x), __fmul_rn(fx, fm.x));
    const float gy = __fadd_rn(linspace_m1_p1(i, H, step_y), __fmul_rn(fy, fm.y));
    SrcPos s;
    s.ix = source_index<PAD, AC>(gx, W);
    s.iy = source_index<PAD, AC>(gy, H);
    s.valid = (gx > -1.0f) && (gy > -1.0f) && (gx < 1.0f) && (gy < 1.0f);
    return s;
}

template <int PAD, bool AC>
__device__ __forceinline__ SrcPos source_position(const float* __restrict__ flow, int b, int i, int j, int H, int W,
                                                  float step_x, float step_y, FlowMul fm = FlowMul{1.0f, 1.0f}) {
    const size_t HW = (size_t)H * W;
    const size_t p = (size_t)i * W + j;
    float fx = __fmul_rn(__ldg(flow + ((size_t)b * 2 + 0) * HW + p), fm.x);
    float fy = __fmul_rn(__ldg(flow + ((size_t)b * 2 + 1) * HW + p), fm.y);
    float gx = __fadd_rn(linspace_m1_p1(j, W, step_x), fx);
    float gy = __fadd_rn(linspace_m1_p1(i, H, step_y), fy);
    SrcPos s;
    s.ix = source_index<PAD, AC>(gx, W);
    s.iy = source_index<PAD, AC>(gy, H);
    s.valid = (gx > -1.0f) && (gy > -1.0f) && (gx < 1.0f) && (gy < 1.0f);
    return s;
}

// ------------------------------------------------------------------------------ direct kernel
template <int MODE, int PAD, bool AC, bool NHWC>
__global__ void __launch_bounds__(256) warp_direct_kernel(const float* __restrict__ frame, const float* __restrict__ flow,
                                                          float* __restrict__ out, uint8_t* __restrict__ valid, int B,
                                                          int C, int H, int W, FlowMul fm) {
    const size_t HW = (size_t)H * W;
    const size_t total = (size_t)B * HW;
    const float step_x = linspace_step(W), step_y = linspace_step(H);
    for (size_t q = (size_t)blockIdx.x * blockDim.x + threadIdx.x; q < total; q += (size_t)gridDim.x * blockDim.x) {
        const int b = (int)(q / HW);
        const size_t p = q - (size_t)b * HW;
        const int i = (int)(p / W), j = (int)(p - (size_t)i * W);
        SrcPos s = source_position<PAD, AC>(flow, b, i, j, H, W, step_x, step_y, fm);
        if (valid) valid[q] = s.valid ? 1 : 0;
        if (MODE == OFB_MODE_NEAREST) {
            int x = (int)rintf(s.ix), y = (int)rintf(s.iy);
            bool in = x >= 0 && x < W && y >= 0 && y < H;
            for (int c = 0; c < C; ++c) {
                float v = 0.0f;
                if (in) v = NHWC ? __ldg(frame + ((size_t)b * HW + (size_t)y * W + x) * C + c)
                                 : __ldg(frame + ((size_t)b * C + c) * HW + (size_t)y * W + x);
                out[((size_t)b * C + c) * HW + p] = v;
            }
            continue;
        }
        const float x0f = floorf(s.ix), y0f = floorf(s.iy);
        const int x0 = (int)x0f, y0 = (int)y0f, x1 = x0 + 1, y1 = y0 + 1;
        const float wx1 = s.ix - x0f, wx0 = (x0f + 1.0f) - s.ix;
        const float wy1 = s.iy - y0f, wy0 = (y0f + 1.0f) - s.iy;
        const float w00 = wx0 * wy0, w01 = wx1 * wy0, w10 = wx0 * wy1, w11 = wx1 * wy1;
        const bool inx0 = x0 >= 0 && x0 < W, inx1 = x1 >= 0 && x1 < W;
        const bool iny0 = y0 >= 0 && y0 < H, iny1 = y1 >= 0 && y1 < H;
        if (NHWC) {
            const float* base = frame + (size_t)b * HW * C;
            const float* p00 = base + ((size_t)y0 * W + x0) * C;
            const float* p01 = p00 + C;
            const float* p10 = p00 + (size_t)W * C;
            const float* p11 = p10 + C;
            if ((C & 3) == 0) {
                for (int c = 0; c < C; c += 4) {
                    float4 a = make_float4(0, 0, 0, 0), bb = a, cc = a, dd = a;
                    if (iny0 && inx0) a = __ldg(reinterpret_cast<const float4*>(p00 + c));
                    if (iny0 && inx1) bb = __ldg(reinterpret_cast<const float4*>(p01 + c));
                    if (iny1 && inx0) cc = __ldg(reinterpret_cast<const float4*>(p10 + c));
                    if (iny1 && inx1) dd = __ldg(reinterpret_cast<const float4*>(p11 + c));
                    float r0 = __fmaf_rn(dd.x, w11, __fmaf_rn(cc.x, w10, __fmaf_rn(bb.x, w01, a.x * w00)));
                    float r1 = __fmaf_rn(dd.y, w11, __fmaf_rn(cc.y, w10, __fmaf_rn(bb.y, w01, a.y * w00)));
                    float r2 = __fmaf_rn(dd.z, w11, __fmaf_rn(cc.z, w10, __fmaf_rn(bb.z, w01, a.z * w00)));
                    float r3 = __fmaf_rn(dd.w, w11, __fmaf_rn(cc.w, w10, __fmaf_rn(bb.w, w01, a.w * w00)));
                    out[((size_t)b * C + c + 0) * HW + p] = r0;
                    out[((size_t)b * C + c + 1) * HW + p] = r1;
                    out[((size_t)b * C + c + 2) * HW + p] = r2;
                    out[((size_t)b * C + c + 3) * HW + p] = r3;
                }
            } else {
                for (int c = 0; c < C; ++c) {
                    float acc = 0.0f;
                    if (iny0 && inx0) acc = __fmaf_rn(__ldg(p00 + c), w00, acc);
                    if (iny0 && inx1) acc = __fmaf_rn(__ldg(p01 + c), w01, acc);
                    if (iny1 && inx0) acc = __fmaf_rn(__ldg(p10 + c), w10, acc);
                    if (iny1 && inx1) acc = __fmaf_rn(__ldg(p11 + c), w11, acc);
                    out[((size_t)b * C + c) * HW + p] = acc;
                }
            }
        } else {
            const size_t o00 = (size_t)y0 * W + x0;
            for (int c = 0; c < C; ++c) {
                const float* plane = frame + ((size_t)b * C + c) * HW;
                float acc = 0.0f;
                if (iny0 && inx0) acc = __fmaf_rn(__ldg(plane + o00), w00, acc);
                if (iny0 && inx1) acc = __fmaf_rn(__ldg(plane + o00 + 1), w01, acc);
                if (iny1 && inx0) acc = __fmaf_rn(__ldg(plane + o00 + W), w10, acc);
                if (iny1 && inx1) acc = __fmaf_rn(__ldg(plane + o00 + W + 1), w11, acc);
                out[((size_t)b * C + c) * HW + p] = acc;
            }
        }
    }
}

// ------------------------------------------------------------------------------ staged kernel
constexpr int TW = 64, TH = 16, NT = 256, PPT = (TW * TH) / NT;  // 4 pixels per thread
constexpr int MARGIN = 16;                                        // staged window = tile +- MARGIN
constexpr int WIN_W = TW + 2 * MARGIN + 8;                        // +8: 16-byte alignment slack on both sides
constexpr int WIN_H = TH + 2 * MARGIN + 2;
constexpr int CG_MAX = 4;                                         // channels staged per pass

__device__ __forceinline__ void cp_async16(void* smem, const void* gmem) {
    unsigned s = (unsigned)__cvta_generic_to_shared(smem);
    asm volatile("cp.async.cg.shared.global [%0], [%1], 16;\n" ::"r"(s), "l"(gmem));
}
__device__ __forceinline__ void cp_async_wait_all() { asm volatile("cp.async.wait_all;\n" ::: "memory"); }

template <int PAD, bool AC>
__global__ void __launch_bounds__(NT) warp_staged_kernel(const float* __restrict__ frame, const float* __restrict__ flow,
                                                         float* __restrict__ out, uint8_t* __restrict__ valid, int B,
                                                         int C, int H, int W, int vec4, FlowMul fm) {
    extern __shared__ __align__(16) float win[];  // [cg][WIN_H][pitch]
    __shared__ int s_box[4];                      // min x, max x, min y, max y of the taps
    const int b = blockIdx.z;
    const int tile_x = blockIdx.x * TW, tile_y = blockIdx.y * TH;
    const int tx = threadIdx.x % TW, ty = threadIdx.x / TW;  // ty in [0, NT/TW)
    const size_t HW = (size_t)H * W;
    const float step_x = linspace_step(W), step_y = linspace_step(H);

    if (threadIdx.x == 0) { s_box[0] = INT_MAX; s_box[1] = INT_MIN; s_box[2] = INT_MAX; s_box[3] = INT_MIN; }
    __syncthreads();

    float ix[PPT], iy[PPT];
    int bx0 = INT_MAX, bx1 = INT_MIN, by0 = INT_MAX, by1 = INT_MIN;
#pragma unroll
    for (int k = 0; k < PPT; ++k) {
        const int i = tile_y + ty + k * (NT / TW), j = tile_x + tx;
        ix[k] = 0.0f; iy[k] = 0.0f;
        if (i < H && j < W) {
            SrcPos s = source_position<PAD, AC>(flow, b, i, j, H, W, step_x, step_y, fm);
            ix[k] = s.ix; iy[k] = s.iy;
            if (valid) valid[(size_t)b * HW + (size_t)i * W + j] = s.valid ? 1 : 0;
            // clamp before the int conversion: zeros padding can leave coordinates far outside
            int x0 = (int)floorf(fminf(fmaxf(s.ix, -2.0f), (float)W + 1.0f));
            int y0 = (int)floorf(fminf(fmaxf(s.iy, -2.0f), (float)H + 1.0f));
            bx0 = min(bx0, x0); bx1 = max(bx1, x0 + 1);
            by0 = min(by0, y0); by1 = max(by1, y0 + 1);
        }
    }
    bx0 = warp_min(bx0); bx1 = warp_max(bx1); by0 = warp_min(by0); by1 = warp_max(by1);
    if ((threadIdx.x & 31) == 0) {
        atomicMin(&s_box[0], bx0); atomicMax(&s_box[1], bx1);
        atomicMin(&s_box[2], by0); atomicMax(&s_box[3], by1);
    }
    __syncthreads();
    // staged window: tap bounding box, clamped to tile +- MARGIN and to the frame
    int wx_lo = max(max(s_box[0], tile_x - MARGIN), 0);
    int wx_hi = min(min(s_box[1], tile_x + TW - 1 + MARGIN), W - 1);
    int wy_lo = max(max(s_box[2], tile_y - MARGIN), 0);
    int wy_hi = min(min(s_box[3], tile_y + TH - 1 + MARGIN), H - 1);
    if (vec4) { wx_lo &= ~3; wx_hi |= 3; if (wx_hi > W - 1) wx_hi = W - 1; }
    const int ww = wx_hi - wx_lo + 1, wh = wy_hi - wy_lo + 1;  // may be <= 0: nothing staged
    const int pitch = vec4 ? ww : (ww | 1);
    const int plane_sz = pitch * max(wh, 0);

    for (int c0 = 0; c0 < C; c0 += CG_MAX) {
        const int cg = min(CG_MAX, C - c0);
        if (ww > 0 && wh > 0) {
            if (vec4) {
                const int w4 = ww >> 2, per_c = w4 * wh;
                for (int e = threadIdx.x; e < per_c * cg; e += NT) {
                    int c = e / per_c, r = e - c * per_c;
                    int yy = r / w4, x4 = r - yy * w4;
                    const float* src = frame + ((size_t)b * C + c0 + c) * HW + (size_t)(wy_lo + yy) * W + wx_lo + x4 * 4;
                    cp_async16(win + c * plane_sz + yy * pitch + x4 * 4, src);
                }
                cp_async_wait_all();
            } else {
                const int per_c = ww * wh;
                for (int e = threadIdx.x; e < per_c * cg; e += NT) {
                    int c = e / per_c, r = e - c * per_c;
                    int yy = r / ww, xx = r - yy * ww;
                    win[c * plane_sz + yy * pitch + xx] =
                        __ldg(frame + ((size_t)b * C + c0 + c) * HW + (size_t)(wy_lo + yy) * W + wx_lo + xx);
                }
            }
        }
        __syncthreads();
#pragma unroll
        for (int k = 0; k < PPT; ++k) {
            const int i = tile_y + ty + k * (NT / TW), j = tile_x + tx;
            if (i >= H || j >= W) continue;
            const float x0f = floorf(ix[k]), y0f = floorf(iy[k]);
            const float wx1 = ix[k] - x0f, wx0 = (x0f + 1.0f) - ix[k];
            const float wy1 = iy[k] - y0f, wy0 = (y0f + 1.0f) - iy[k];
            const float w00 = wx0 * wy0, w01 = wx1 * wy0, w10 = wx0 * wy1, w11 = wx1 * wy1;
            const int x0 = (int)fminf(fmaxf(x0f, -2.0f), (float)W + 1.0f);
            const int y0 = (int)fminf(fmaxf(y0f, -2.0f), (float)H + 1.0f);
            const int x1 = x0 + 1, y1 = y0 + 1;
            const bool fast = x0 >= wx_lo && x1 <= wx_hi && y0 >= wy_lo && y1 <= wy_hi;
            const size_t p = (size_t)i * W + j;
            if (fast) {
                const float* s00 = win + (y0 - wy_lo) * pitch + (x0 - wx_lo);
                for (int c = 0; c < cg; ++c) {
                    const float* s = s00 + c * plane_sz;
                    float acc = s[0] * w00;
                    acc = __fmaf_rn(s[1], w01, acc);
                    acc = __fmaf_rn(s[pitch], w10, acc);
                    acc = __fmaf_rn(s[pitch + 1], w11, acc);
                    out[((size_t)b * C + c0 + c) * HW + p] = acc;
                }
            } else {
                const bool inx0 = x0 >= 0 && x0 < W, inx1 = x1 >= 0 && x1 < W;
                const bool iny0 = y0 >= 0 && y0 < H, iny1 = y1 >= 0 && y1 < H;
                for (int c = 0; c < cg; ++c) {
                    const float* plane = frame + ((size_t)b * C + c0 + c) * HW;
                    float acc = 0.0f;
                    if (iny0 && inx0) acc = __fmaf_rn(__ldg(plane + (size_t)y0 * W + x0), w00, acc);
                    if (iny0 && inx1) acc = __fmaf_rn(__ldg(plane + (size_t)y0 * W + x1), w01, acc);
                    if (iny1 && inx0) acc = __fmaf_rn(__ldg(plane + (size_t)y1 * W + x0), w10, acc);
                    if (iny1 && inx1) acc = __fmaf_rn(__ldg(plane + (size_t)y1 * W + x1), w11, acc);
                    out[((size_t)b * C + c0 + c) * HW + p] = acc;
                }
            }
        }
        __syncthreads();
    }
}

// ------------------------------------------------------------------------------ row kernel
// The default NCHW bilinear path.  grid = (column groups, row groups, batch): no per-pixel integer division,
// 32-bit in-plane offsets.  Lane <-> column (every load / store of a warp is one or two 128-byte lines), and a
// thread walks RPT consecutive rows of its column: the lower taps of row i are the upper taps of row i+1 for
// smooth flows, so they hit L1.  The coordinate pipeline and the 4 tap offsets are computed once per pixel
// and reused for every channel.
constexpr int RPT_DEFAULT = 2;   // rows per thread: 32 registers, 64 resident warps per SM (see the sweep in DESIGN.md)

template <int PAD, bool AC, int RPT>
__global__ void __launch_bounds__(128) warp_rows_kernel(const float* __restrict__ frame, const float* __restrict__ flow,
                                                        float* __restrict__ out, uint8_t* __restrict__ valid, int C,
                                                        int H, int W, FlowMul fm) {
    const int j = blockIdx.x * blockDim.x + threadIdx.x, b = blockIdx.z;
    const int i0 = blockIdx.y * RPT;
    if (j >= W) return;
    const int HW = H * W;                                     // launch guard: H*W < 2^30
    const float step_x = fm.step_x, step_y = fm.step_y;
    const float* fxp = flow + (size_t)(b * 2) * HW;
    const float* fyp = fxp + HW;
    int o00[RPT];
    float w00[RPT], w01[RPT], w10[RPT], w11[RPT];
    unsigned inb = 0;
#pragma unroll
    for (int k = 0; k < RPT; ++k) {
        const int i = i0 + k;
        o00[k] = 0; w00[k] = w01[k] = w10[k] = w11[k] = 0.0f;
        if (i >= H) continue;
        const int off = i * W + j;
        const SrcPos s = source_from_flow<PAD, AC>(__ldg(fxp + off), __ldg(fyp + off), i, j, H, W, step_x, step_y, fm);
        if (valid) valid[(size_t)b * HW + off] = s.valid ? 1 : 0;
        const float x0f = floorf(s.ix), y0f = floorf(s.iy);
        const float wx1 = s.ix - x0f, wx0 = (x0f + 1.0f) - s.ix;
        const float wy1 = s.iy - y0f, wy0 = (y0f + 1.0f) - s.iy;
        w00[k] = wx0 * wy0; w01[k] = wx1 * wy0; w10[k] = wx0 * wy1; w11[k] = wx1 * wy1;
        int x0, y0;
        bool inx0, inx1, iny0, iny1;
        if (PAD != OFB_PAD_ZEROS) {
            // border / reflection clip the coordinate into [0, size-1] (NaN clips to 0): the upper-left tap is
            // always inside, only the +1 taps can fall off the far edge (their weight is 0 there)
            x0 = (int)x0f; y0 = (int)y0f;
            inx0 = true; iny0 = true; inx1 = x0 + 1 < W; iny1 = y0 + 1 < H;
        } else {
            // clamp before the int conversion: zeros padding can leave coordinates far outside; NaN samples nothing
            x0 = (int)fminf(fmaxf(x0f, -2.0f), (float)W + 1.0f); y0 = (int)fminf(fmaxf(y0f, -2.0f), (float)H + 1.0f);
            const bool fin = x0f == x0f && y0f == y0f;
            inx0 = fin && x0 >= 0 && x0 < W; inx1 = fin && x0 + 1 >= 0 && x0 + 1 < W;
            iny0 = y0 >= 0 && y0 < H; iny1 = y0 + 1 >= 0 && y0 + 1 < H;
        }
        inb |= (((iny0 && inx0) ? 1u : 0u) | ((iny0 && inx1) ? 2u : 0u) | ((iny1 && inx0) ? 4u : 0u) |
                ((iny1 && inx1) ? 8u : 0u)) << (4 * k);
        o00[k] = y0 * W + x0;
    }
    // interior pixels (all 4 x RPT taps inside the frame -- everything but the last row / column and, under zeros
    // padding, samples that leave the frame): no predicates, no per-tap address arithmetic.  The kernel is bound
    // by instruction issue, not by HBM, so the channel loop is kept to 4 loads + 4 FMAs + 1 store per pixel.
    if (inb == (RPT == 8 ? 0xFFFFFFFFu : (1u << (4 * RPT)) - 1u)) {
        const float* plane = frame + (size_t)(b * C) * HW;
        float* op = out + (size_t)(b * C) * HW + i0 * W + j;
#pragma unroll 1
        for (int c = 0; c < C; ++c, plane += HW, op += HW) {
            float v[RPT][4];                              // all 4 x RPT gathers in flight before the first FMA
#pragma unroll
            for (int k = 0; k < RPT; ++k) {
                const float* p = plane + o00[k];
                const float* q = p + W;
                v[k][0] = __ldg(p); v[k][1] = __ldg(p + 1); v[k][2] = __ldg(q); v[k][3] = __ldg(q + 1);
            }
#pragma unroll
            for (int k = 0; k < RPT; ++k) {
                float acc = __fmaf_rn(v[k][0], w00[k], 0.0f);
                acc = __fmaf_rn(v[k][1], w01[k], acc);
                acc = __fmaf_rn(v[k][2], w10[k], acc);
                acc = __fmaf_rn(v[k][3], w11[k], acc);
                op[k * W] = acc;
            }
        }
        return;
    }
    for (int c = 0; c < C; ++c) {
        const float* plane = frame + (size_t)(b * C + c) * HW;
        float* op = out + (size_t)(b * C + c) * HW + i0 * W + j;
#pragma unroll
        for (int k = 0; k < RPT; ++k) {
            if (i0 + k >= H) continue;
            const unsigned m = inb >> (4 * k);
            const float* p = plane + o00[k];
            float acc = 0.0f;
            if (m & 1u) acc = __fmaf_rn(__ldg(p), w00[k], acc);
            if (m & 2u) acc = __fmaf_rn(__ldg(p + 1), w01[k], acc);
            if (m & 4u) acc = __fmaf_rn(__ldg(p + W), w10[k], acc);
            if (m & 8u) acc = __fmaf_rn(__ldg(p + W + 1), w11[k], acc);
            op[k * W] = acc;
        }
    }
}

__global__ void __launch_bounds__(256) warp_grid_kernel(const float* __restrict__ flow, float* __restrict__ grid, int B,
                                                        int H, int W) {
    const size_t total = (size_t)B * H * W;
    const float step_x = linspace_step(W), step_y = linspace_step(H);
    for (size_t q = (size_t)blockIdx.x * blockDim.x + threadIdx.x; q < total; q += (size_t)gridDim.x * blockDim.x) {
        const int j = (int)(q % W), i = (int)((q / W) % H);
        float2 f = __ldg(reinterpret_cast<const float2*>(flow) + q);
        float2 g;
        g.x = __fadd_rn(linspace_m1_p1(j, W, step_x), f.x);
        g.y = __fadd_rn(linspace_m1_p1(i, H, step_y), f.y);
        reinterpret_cast<float2*>(grid)[q] = g;
    }
}

template <int MODE, int PAD, bool AC>
int launch_direct(const float* frame, const float* flow, float* out, uint8_t* valid, int B, int C, int H, int W,
                  int channels_last, FlowMul fm, cudaStream_t st) {
    const size_t total = (size_t)B * H * W;
    int blocks = (int)((total + 255) / 256);
    const int cap = ofb_num_sms() * 32;
    if (blocks > cap) blocks = cap;
    if (channels_last)
        warp_direct_kernel<MODE, PAD, AC, true><<<blocks, 256, 0, st>>>(frame, flow, out, valid, B, C, H, W, fm);
    else
        warp_direct_kernel<MODE, PAD, AC, false><<<blocks, 256, 0, st>>>(frame, flow, out, valid, B, C, H, W, fm);
    OFB_LAUNCH_CHECK();
    return OFB_OK;
}

template <int PAD, bool AC>
int launch_rows(const float* frame, const float* flow, float* out, uint8_t* valid, int B, int C, int H, int W,
                FlowMul fm, cudaStream_t st) {
    static int rpt = 0;
    if (!rpt) {
        const char* e = getenv("OFB_WARP_RPT");           // tuning override
        rpt = e ? atoi(e) : RPT_DEFAULT;
        if (rpt != 1 && rpt != 2 && rpt != 4 && rpt != 8) rpt = RPT_DEFAULT;
    }
    const dim3 block(128);
    const dim3 grid((W + 127) / 128, (H + rpt - 1) / rpt, B);
    if (grid.y > 65535) return OFB_EUNSUPPORTED;
    if (rpt == 1) warp_rows_kernel<PAD, AC, 1><<<grid, block, 0, st>>>(frame, flow, out, valid, C, H, W, fm);
    else if (rpt == 2) warp_rows_kernel<PAD, AC, 2><<<grid, block, 0, st>>>(frame, flow, out, valid, C, H, W, fm);
    else if (rpt == 8) warp_rows_kernel<PAD, AC, 8><<<grid, block, 0, st>>>(frame, flow, out, valid, C, H, W, fm);
    else warp_rows_kernel<PAD, AC, 4><<<grid, block, 0, st>>>(frame, flow, out, valid, C, H, W, fm);
    OFB_LAUNCH_CHECK();
    return OFB_OK;
}

template <int PAD, bool AC>
int launch_staged(const float* frame, const float* flow, float* out, uint8_t* valid, int B, int C, int H, int W,
                  FlowMul fm, cudaStream_t st) {
    const int cg = C < CG_MAX ? C : CG_MAX;
    const size_t smem = (size_t)cg * WIN_H * WIN_W * sizeof(float);
    static bool configured[OFB_MAX_DEVICES] = {false};
    const int dev = ofb_device();
    if (!configured[dev]) {
        OFB_CUDA(cudaFuncSetAttribute(warp_staged_kernel<PAD, AC>, cudaFuncAttributeMaxDynamicSharedMemorySize,
                                      (int)(CG_MAX * WIN_H * WIN_W * sizeof(float))));
        configured[dev] = true;
    }
    const int vec4 = (W % 4 == 0) && ((reinterpret_cast<uintptr_t>(frame) & 15) == 0);
    dim3 grid((W + TW - 1) / TW, (H + TH - 1) / TH, B);
    warp_staged_kernel<PAD, AC><<<grid, NT, smem, st>>>(frame, flow, out, valid, B, C, H, W, vec4, fm);
    OFB_LAUNCH_CHECK();
    return OFB_OK;
}

template <int MODE>
int dispatch_direct(int pad, int ac, const float* frame, const float* flow, float* out, uint8_t* valid, int B, int C,
                    int H, int W, int cl, FlowMul fm, cudaStream_t st) {
#define OFB_CASE(P, A)                                                                                 \
    if (pad == P && ac == (A ? 1 : 0))                                                                  \
        return launch_direct<MODE, P, A>(frame, flow, out, valid, B, C, H, W, cl, fm, st);
    OFB_CASE(OFB_PAD_ZEROS, false)
    OFB_CASE(OFB_PAD_ZEROS, true)
    OFB_CASE(OFB_PAD_BORDER, false)
    OFB_CASE(OFB_PAD_BORDER, true)
    OFB_CASE(OFB_PAD_REFLECTION, false)
    OFB_CASE(OFB_PAD_REFLECTION, true)
#undef OFB_CASE
    return OFB_EINVAL;
}

int dispatch_staged(int pad, int ac, int rows, const float* frame, const float* flow, float* out, uint8_t* valid, int B,
                    int C, int H, int W, FlowMul fm, cudaStream_t st) {
#define OFB_CASE(P, A)                                                                                   \
    if (pad == P && ac == (A ? 1 : 0))                                                                    \
        return rows ? launch_rows<P, A>(frame, flow, out, valid, B, C, H, W, fm, st)                      \
                    : launch_staged<P, A>(frame, flow, out, valid, B, C, H, W, fm, st);
    OFB_CASE(OFB_PAD_ZEROS, false)
    OFB_CASE(OFB_PAD_ZEROS, true)
    OFB_CASE(OFB_PAD_BORDER, false)
    OFB_CASE(OFB_PAD_BORDER, true)
    OFB_CASE(OFB_PAD_REFLECTION, false)
    OFB_CASE(OFB_PAD_REFLECTION, true)
#undef OFB_CASE
    return OFB_EINVAL;
}

}  // namespace

// warp_tma.cu
int ofb_warp_tma_launch(const float* frame, const float* flow, float* out, uint8_t* valid, int B, int C, int H, int W,
                        int pad, int ac, float fmx, float fmy, cudaStream_t st);

OFB_API int ofb_warp_f32(const float* frame, const float* flow, float* out, uint8_t* valid_or_null, int B, int C, int H,
                         int W, int mode, int padding_mode, int align_corners, int channels_last, int variant,
                         float flow_mul_x, float flow_mul_y, void* stream) {
    if (B == 0 || C == 0 || H == 0 || W == 0) return OFB_OK;   // nothing to do: empty tensors have no storage, their pointers may be null
    if (!frame || !flow || !out || B < 0 || C < 0 || H < 0 || W < 0) return OFB_EINVAL;
    if (mode != OFB_MODE_BILINEAR && mode != OFB_MODE_NEAREST) return OFB_EINVAL;
    if (padding_mode < 0 || padding_mode > 2 || variant < 0 || variant > 4) return OFB_EINVAL;
    if ((size_t)B * C * H * W == 0) return OFB_OK;
    if (B > 65535) return OFB_EUNSUPPORTED;
    cudaStream_t st = (cudaStream_t)stream;
    const FlowMul fm{flow_mul_x, flow_mul_y, ofb::linspace_step(W), ofb::linspace_step(H)};
    const bool nchw_bilinear = mode == OFB_MODE_BILINEAR && !channels_last;
    const bool can_rows = nchw_bilinear && H <= 65535 * RPT_DEFAULT && (long long)H * W < (1LL << 30) && (long long)B * C * H * W < (1LL << 40);
    // TMA needs 16-byte global strides
    const bool can_tma = can_rows && W % 4 == 0 && (reinterpret_cast<uintptr_t>(frame) & 15) == 0 &&
                         (long long)B * C < (1LL << 31);
    if ((variant == 2 && !nchw_bilinear) || (variant == 3 && !can_rows) || (variant == 4 && !can_tma))
        return OFB_EUNSUPPORTED;
    // auto: the row kernel; 1 = direct gather (also nearest / NHWC), 2 = cp.async-staged, 3 = row kernel,
    // 4 = TMA-staged window (warp_tma.cu)
    static int auto_v = -1;
    if (auto_v < 0) {
        const char* e = getenv("OFB_WARP_VARIANT");       // tuning override for variant 0
        auto_v = e ? atoi(e) : 0;
    }
    int v = variant != 0 ? variant : (auto_v == 4 && can_tma ? 4 : (can_rows ? 3 : 1));
    if (v == 4)
        return ofb_warp_tma_launch(frame, flow, out, valid_or_null, B, C, H, W, padding_mode, align_corners, fm.x, fm.y, st);
    if (v == 2 || v == 3)
        return dispatch_staged(padding_mode, align_corners, v == 3, frame, flow, out, valid_or_null, B, C, H, W, fm, st);
    if (mode == OFB_MODE_BILINEAR)
        return dispatch_direct<OFB_MODE_BILINEAR>(padding_mode, align_corners, frame, flow, out, valid_or_null, B, C, H,
                                                  W, channels_last, fm, st);
    return dispatch_direct<OFB_MODE_NEAREST>(padding_mode, align_corners, frame, flow, out, valid_or_null, B, C, H, W,
                                             channels_last, fm, st);
}

OFB_API int ofb_warp_grid_f32(const float* flow_bhw2, float* grid_bhw2, int B, int H, int W, void* stream) {
    if (B == 0 || H == 0 || W == 0) return OFB_OK;   // nothing to do: empty tensors have no storage, their pointers may be null
    if (!flow_bhw2 || !grid_bhw2 || B < 0 || H < 0 || W < 0) return OFB_EINVAL;
    const size_t total = (size_t)B * H * W;
    if (total == 0) return OFB_OK;
    int blocks = (int)((total + 255) / 256);
    const int cap = ofb_num_sms() * 32;
    if (blocks > cap) blocks = cap;
    warp_grid_kernel<<<blocks, 256, 0, (cudaStream_t)stream>>>(flow_bhw2, grid_bhw2, B, H, W);
    OFB_LAUNCH_CHECK();
    return OFB_OK;
}
