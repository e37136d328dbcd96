/*
 * oracle.c -- CPU restatement of the reference hot path.  TEST INFRASTRUCTURE ONLY.
 *
 * Only tests/, __graft_entry__.smoke() and bench.py's cpu_baseline / --impl reference
 * legs may load this library; the product (torch-optical-flow_b200/) never does.
 *
 * The reference (awaelchli/torch-optical-flow) is pure Python; the arithmetic of the
 * path lives in a third-party dependency that is NOT under /root/reference:
 *     torch (ATen)   requirement `torch>=1.5.1` (requirements/base.txt:2, unpinned;
 *                    2.11.0+cu128 installed here and used to pin this file).
 * Each function below restates the published ATen algorithm behind one reference call
 * site and cites that call site (paths relative to /root/reference) plus the ATen
 * header that states the formula (paths relative to torch/include/ATen/native).
 *
 * Pinning: oracle outputs are checked against golden vectors produced by importing the
 * reference itself in the authoring container (tests/golden/make_golden.py ->
 * tests/golden/<case>.npz) and against the known-answer vectors of the reference's own
 * tests/operator/test_operator.py.  See tests/test_oracle_golden.py.
 *
 * Build: see oracle/Makefile (-ffp-contract=off: every fused multiply-add below is an
 * explicit fmaf(), every other operation is separately rounded).
 */
#include <math.h>
#include <stdint.h>
#include <stdlib.h>
#include <string.h>

#define ORC_API __attribute__((visibility("default")))

/* ------------------------------------------------------------------------------------
 * torch.linspace(start, end, n) for fp32 on CPU (operator.py:49-50, corr.py:64-65).
 * ATen RangeFactories: step = (end-start)/(n-1); symmetric halves, each element one
 * fused op (bit-exact vs torch 2.11 CPU for n in {2..2048}, see tests).
 * ---------------------------------------------------------------------------------- */
ORC_API void orc_linspace_f32(float start, float end, int n, float* out) {
    if (n <= 0) return;
    if (n == 1) { out[0] = start; return; }
    float step = (end - start) / (float)(n - 1);
    int half = n / 2;
    for (int i = 0; i < n; ++i) {
        if (i < half) out[i] = fmaf(step, (float)i, start);
        else          out[i] = fmaf(-step, (float)(n - 1 - i), end);
    }
}

/* ------------------------------------------------------------------------------------
 * F.grid_sample coordinate pipeline (GridSampler.h:26-36 unnormalise, :57-59 clip,
 * :88-107 reflect, :143-160 padding dispatch).
 * padding: 0 = zeros, 1 = border, 2 = reflection.
 * ---------------------------------------------------------------------------------- */
static inline float orc_unnormalize(float g, int size, int align_corners) {
    if (align_corners) {
        /* ((g + 1) / 2) * (size - 1) */
        return ((g + 1.0f) / 2.0f) * (float)(size - 1);
    }
    /* ((g + 1) * size - 1) / 2  ==  fma(g + 1, size/2, -0.5)  (one rounding; the CPU
     * vector kernel and the nvcc-contracted CUDA kernel both evaluate it fused) */
    return fmaf(g + 1.0f, 0.5f * (float)size, -0.5f);
}

static inline float orc_clip(float x, int size) {
    return fminf((float)(size - 1), fmaxf(x, 0.0f));
}

static inline float orc_reflect(float in, int twice_low, int twice_high) {
    if (twice_low == twice_high) return 0.0f;
    float mn = (float)twice_low / 2.0f;
    float span = (float)(twice_high - twice_low) / 2.0f;
    in = fabsf(in - mn);
    float extra = fmodf(in, span);
    int flips = (int)floorf(in / span);
    return (flips % 2 == 0) ? (extra + mn) : (span - extra + mn);
}

static inline float orc_source_index(float g, int size, int padding, int align_corners) {
    float x = orc_unnormalize(g, size, align_corners);
    if (padding == 1) {
        x = orc_clip(x, size);
    } else if (padding == 2) {
        x = align_corners ? orc_reflect(x, 0, 2 * (size - 1)) : orc_reflect(x, -1, 2 * size - 1);
        x = orc_clip(x, size);
    }
    return x;
}

/* One bilinear sample of a single-channel HxW plane; taps outside the plane contribute
 * zero (grid_sampler_2d, bilinear).  Taps are accumulated nw, ne, sw, se, each with one
 * fused multiply-add -- bit-exact vs ATen's CPU kernel on the golden vectors. */
static inline float orc_bilinear_tap(const float* plane, int H, int W, float ix, float iy) {
    float x0f = floorf(ix), y0f = floorf(iy);
    int x0 = (int)x0f, y0 = (int)y0f, x1 = x0 + 1, y1 = y0 + 1;
    float wx1 = ix - x0f, wx0 = (x0f + 1.0f) - ix;
    float wy1 = iy - y0f, wy0 = (y0f + 1.0f) - iy;
    float acc = 0.0f;
    if (y0 >= 0 && y0 < H) {
        if (x0 >= 0 && x0 < W) acc = fmaf(plane[(size_t)y0 * W + x0], wx0 * wy0, acc);
        if (x1 >= 0 && x1 < W) acc = fmaf(plane[(size_t)y0 * W + x1], wx1 * wy0, acc);
    }
    if (y1 >= 0 && y1 < H) {
        if (x0 >= 0 && x0 < W) acc = fmaf(plane[(size_t)y1 * W + x0], wx0 * wy1, acc);
        if (x1 >= 0 && x1 < W) acc = fmaf(plane[(size_t)y1 * W + x1], wx1 * wy1, acc);
    }
    return acc;
}

static inline float orc_nearest_tap(const float* plane, int H, int W, float ix, float iy) {
    int x = (int)nearbyintf(ix), y = (int)nearbyintf(iy);
    if (x >= 0 && x < W && y >= 0 && y < H) return plane[(size_t)y * W + x];
    return 0.0f;
}

/* ------------------------------------------------------------------------------------
 * F.grid_sample(img (N,C,H,W), grid (N,Ho,Wo,2)) -> (N,C,Ho,Wo)
 * mode: 0 bilinear, 1 nearest.
 * ---------------------------------------------------------------------------------- */
ORC_API void orc_grid_sample_f32(const float* img, const float* grid, float* out,
                                 int N, int C, int H, int W, int Ho, int Wo,
                                 int mode, int padding, int align_corners) {
#pragma omp parallel for collapse(2) schedule(static)
    for (int n = 0; n < N; ++n) {
        for (int i = 0; i < Ho; ++i) {
            for (int j = 0; j < Wo; ++j) {
                const float* g = grid + (((size_t)n * Ho + i) * Wo + j) * 2;
                float ix = orc_source_index(g[0], W, padding, align_corners);
                float iy = orc_source_index(g[1], H, padding, align_corners);
                for (int c = 0; c < C; ++c) {
                    const float* plane = img + ((size_t)n * C + c) * H * W;
                    float v = mode == 0 ? orc_bilinear_tap(plane, H, W, ix, iy)
                                        : orc_nearest_tap(plane, H, W, ix, iy);
                    out[(((size_t)n * C + c) * Ho + i) * Wo + j] = v;
                }
            }
        }
    }
}

/* ------------------------------------------------------------------------------------
 * optical_flow.warp (operator.py:8-33) with warp_grid (operator.py:36-56) folded in:
 *   grid[b,i,j] = (linspace(-1,1,W)[j] + flow[b,0,i,j], linspace(-1,1,H)[i] + flow[b,1,i,j])
 *   out = grid_sample(frame, grid, mode, padding_mode, align_corners)
 * valid_or_null (B,H,W) u8: 1 iff the *unclamped* source position lies inside the frame,
 * i.e. -1 < g < 1 on both axes -- the predicate of bilinear_sampler's mask
 * (methods/raft/model/utils.py:76-78) applied to warp's normalised grid.
 * ---------------------------------------------------------------------------------- */
ORC_API void orc_warp_f32(const float* frame, const float* flow, float* out, uint8_t* valid_or_null,
                          int B, int C, int H, int W, int mode, int padding, int align_corners) {
    float* lx = (float*)malloc(sizeof(float) * (size_t)W);
    float* ly = (float*)malloc(sizeof(float) * (size_t)H);
    orc_linspace_f32(-1.0f, 1.0f, W, lx);
    orc_linspace_f32(-1.0f, 1.0f, H, ly);
    size_t HW = (size_t)H * W;
#pragma omp parallel for collapse(2) schedule(static)
    for (int b = 0; b < B; ++b) {
        for (int i = 0; i < H; ++i) {
            for (int j = 0; j < W; ++j) {
                float gx = lx[j] + flow[((size_t)b * 2 + 0) * HW + (size_t)i * W + j];
                float gy = ly[i] + flow[((size_t)b * 2 + 1) * HW + (size_t)i * W + j];
                float ix = orc_source_index(gx, W, padding, align_corners);
                float iy = orc_source_index(gy, H, padding, align_corners);
                for (int c = 0; c < C; ++c) {
                    const float* plane = frame + ((size_t)b * C + c) * HW;
                    float v = mode == 0 ? orc_bilinear_tap(plane, H, W, ix, iy)
                                        : orc_nearest_tap(plane, H, W, ix, iy);
                    out[((size_t)b * C + c) * HW + (size_t)i * W + j] = v;
                }
                if (valid_or_null)
                    valid_or_null[(size_t)b * HW + (size_t)i * W + j] =
                        (gx > -1.0f) && (gy > -1.0f) && (gx < 1.0f) && (gy < 1.0f);
            }
        }
    }
    free(lx);
    free(ly);
}

/* ------------------------------------------------------------------------------------
 * F.interpolate(x, size, mode="bilinear", align_corners) (operator.py:112, utils.py:91)
 * UpSample.h:259-313 (scale, source index), :442-476 (index / lambda guard).
 * Then an optional per-channel multiply (scale(), operator.py:59-82 / the "8 *" of
 * utils.py:91): channel c of the C channels is multiplied by mul[c % 2].
 * ---------------------------------------------------------------------------------- */
static inline void orc_src_index(float ratio, int dst, int in_size, int out_size, int align_corners,
                                 int* i0, int* i1, float* l0, float* l1) {
    if (out_size == in_size) { *i0 = dst; *i1 = dst; *l0 = 1.0f; *l1 = 0.0f; return; }
    float r;
    if (align_corners) {
        r = ratio * (float)dst;
    } else {
        r = fmaf(ratio, (float)dst + 0.5f, -0.5f);
        if (r < 0.0f) r = 0.0f;
    }
    int idx = (int)floorf(r);
    if (idx > in_size - 1) idx = in_size - 1;
    float lam = r - (float)idx;
    lam = fminf(fmaxf(lam, 0.0f), 1.0f);
    *i0 = idx;
    *i1 = idx + (idx < in_size - 1 ? 1 : 0);
    *l1 = lam;
    *l0 = 1.0f - lam;
}

ORC_API void orc_resize_bilinear_f32(const float* in, float* out, int N, int C, int H, int W,
                                     int Ho, int Wo, int align_corners, float mul_x, float mul_y) {
    float rh, rw;
    if (align_corners) {
        rh = Ho > 1 ? (float)(H - 1) / (float)(Ho - 1) : 0.0f;
        rw = Wo > 1 ? (float)(W - 1) / (float)(Wo - 1) : 0.0f;
    } else {
        rh = (float)H / (float)Ho;
        rw = (float)W / (float)Wo;
    }
#pragma omp parallel for collapse(2) schedule(static)
    for (int n = 0; n < N; ++n) {
        for (int c = 0; c < C; ++c) {
            const float* src = in + ((size_t)n * C + c) * H * W;
            float* dst = out + ((size_t)n * C + c) * Ho * Wo;
            float mul = (c % 2 == 0) ? mul_x : mul_y;
            for (int oy = 0; oy < Ho; ++oy) {
                int y0, y1; float ly0, ly1;
                orc_src_index(rh, oy, H, Ho, align_corners, &y0, &y1, &ly0, &ly1);
                for (int ox = 0; ox < Wo; ++ox) {
                    int x0, x1; float lx0, lx1;
                    orc_src_index(rw, ox, W, Wo, align_corners, &x0, &x1, &lx0, &lx1);
                    /* h0*(w0*a + w1*b) + h1*(w0*c + w1*d), each sum contracted as
                     * fma(first product's factors, second product) -- bit-exact vs ATen CPU */
                    float top = fmaf(lx0, src[(size_t)y0 * W + x0], lx1 * src[(size_t)y0 * W + x1]);
                    float bot = fmaf(lx0, src[(size_t)y1 * W + x0], lx1 * src[(size_t)y1 * W + x1]);
                    dst[(size_t)oy * Wo + ox] = fmaf(ly0, top, ly1 * bot) * mul;
                }
            }
        }
    }
}

/* ------------------------------------------------------------------------------------
 * F.avg_pool2d(x, 2, stride=2) on (N,1,H,W) (corr.py:52-54): floor output size, sum in
 * raster order then divide by the window size.
 * ---------------------------------------------------------------------------------- */
ORC_API void orc_avg_pool2_f32(const float* in, float* out, int64_t N, int H, int W) {
    int Ho = H / 2, Wo = W / 2;
#pragma omp parallel for schedule(static)
    for (int64_t n = 0; n < N; ++n) {
        const float* src = in + (size_t)n * H * W;
        float* dst = out + (size_t)n * Ho * Wo;
        for (int y = 0; y < Ho; ++y)
            for (int x = 0; x < Wo; ++x) {
                float s = src[(size_t)(2 * y) * W + 2 * x];
                s += src[(size_t)(2 * y) * W + 2 * x + 1];
                s += src[(size_t)(2 * y + 1) * W + 2 * x];
                s += src[(size_t)(2 * y + 1) * W + 2 * x + 1];
                dst[(size_t)y * Wo + x] = s / 4.0f;
            }
    }
}

/* ------------------------------------------------------------------------------------
 * CorrBlock.__call__ (corr.py:56-77) + bilinear_sampler (utils.py:64-80):
 * for every query q = (b, y, x), level l and window cell (i, j):
 *     cx = coords[b,0,y,x] / 2^l + (i - r)        <- delta[i,j] = (dy[i], dx[j]) is added
 *     cy = coords[b,1,y,x] / 2^l + (j - r)           to (x, y): the window is transposed
 *     gx = 2*cx/(W_l-1) - 1 ; gy = 2*cy/(H_l-1) - 1                      (utils.py:70-71)
 *     ix = ((gx+1)/2)*(W_l-1) ; iy likewise                              (GridSampler.h:30)
 *     out[b, l*(2r+1)^2 + i*(2r+1) + j, y, x] = bilinear(pyr_l[q], ix, iy), zeros outside
 * lvl_ptrs[l] -> slice of query q starts at lvl_ptrs[l] + q*q_stride[l], rows row_pitch[l]
 * apart (elements).  Optional extra outputs (the bit-exact part of the contract):
 *     idx   (Q, L, 2, 2r+1) int32 : floor(ix) per i (axis 0) and floor(iy) per j (axis 1)
 *     valid (Q, L, (2r+1)^2) u8   : utils.py:77 predicate on the normalised (gx, gy)
 * ---------------------------------------------------------------------------------- */
ORC_API void orc_corr_lookup_f32(const float* const* lvl_ptrs, const int64_t* q_stride,
                                 const int* row_pitch, const int* lvl_h, const int* lvl_w,
                                 const float* coords, float* out, int32_t* idx_or_null,
                                 uint8_t* valid_or_null, int B, int h, int w, int levels, int radius) {
    const int D = 2 * radius + 1;
    const int64_t HW = (int64_t)h * w;
    const int64_t Q = (int64_t)B * HW;
#pragma omp parallel for schedule(static)
    for (int64_t q = 0; q < Q; ++q) {
        int64_t b = q / HW, p = q % HW;
        float cx0 = coords[(b * 2 + 0) * HW + p];
        float cy0 = coords[(b * 2 + 1) * HW + p];
        float ixs[64], iys[64], gxs[64], gys[64];
        for (int l = 0; l < levels; ++l) {
            const int Hl = lvl_h[l], Wl = lvl_w[l];
            const float* slice = lvl_ptrs[l] + q * q_stride[l];
            const float div = (float)(1 << l);
            const float cx = cx0 / div, cy = cy0 / div;
            for (int t = 0; t < D; ++t) {
                float d = (float)(t - radius);
                float x = cx + d, y = cy + d;
                float gx = (2.0f * x) / (float)(Wl - 1) - 1.0f;
                float gy = (2.0f * y) / (float)(Hl - 1) - 1.0f;
                gxs[t] = gx; gys[t] = gy;
                ixs[t] = ((gx + 1.0f) / 2.0f) * (float)(Wl - 1);
                iys[t] = ((gy + 1.0f) / 2.0f) * (float)(Hl - 1);
                if (idx_or_null) {
                    idx_or_null[((q * levels + l) * 2 + 0) * D + t] = (int32_t)floorf(ixs[t]);
                    idx_or_null[((q * levels + l) * 2 + 1) * D + t] = (int32_t)floorf(iys[t]);
                }
            }
            for (int i = 0; i < D; ++i) {
                for (int j = 0; j < D; ++j) {
                    float ix = ixs[i], iy = iys[j];
                    float x0f = floorf(ix), y0f = floorf(iy);
                    int x0 = (int)x0f, y0 = (int)y0f, x1 = x0 + 1, y1 = y0 + 1;
                    float wx1 = ix - x0f, wx0 = (x0f + 1.0f) - ix;
                    float wy1 = iy - y0f, wy0 = (y0f + 1.0f) - iy;
                    float acc = 0.0f;
                    if (y0 >= 0 && y0 < Hl) {
                        if (x0 >= 0 && x0 < Wl) acc = fmaf(slice[(size_t)y0 * row_pitch[l] + x0], wx0 * wy0, acc);
                        if (x1 >= 0 && x1 < Wl) acc = fmaf(slice[(size_t)y0 * row_pitch[l] + x1], wx1 * wy0, acc);
                    }
                    if (y1 >= 0 && y1 < Hl) {
                        if (x0 >= 0 && x0 < Wl) acc = fmaf(slice[(size_t)y1 * row_pitch[l] + x0], wx0 * wy1, acc);
                        if (x1 >= 0 && x1 < Wl) acc = fmaf(slice[(size_t)y1 * row_pitch[l] + x1], wx1 * wy1, acc);
                    }
                    int ch = l * D * D + i * D + j;
                    out[(b * (int64_t)(levels * D * D) + ch) * HW + p] = acc;
                    if (valid_or_null)
                        valid_or_null[(q * levels + l) * D * D + i * D + j] =
                            (gxs[i] > -1.0f) && (gys[j] > -1.0f) && (gxs[i] < 1.0f) && (gys[j] < 1.0f);
                }
            }
        }
    }
}

/* ------------------------------------------------------------------------------------
 * RAFT.upsample_flow (methods/raft/model/raft.py:73-85): convex 8x upsampling.
 *   out[n,c,8y+i,8x+j] = sum_k softmax_k(mask[n, k*64+i*8+j, y, x]) * 8*flow_pad[n,c,y+ky-1,x+kx-1]
 * k = ky*3 + kx over the zero-padded 3x3 neighbourhood (F.unfold, padding=1).
 * softmax as ATen: exp(v - max) / sum(exp(v - max)).
 * ---------------------------------------------------------------------------------- */
ORC_API void orc_convex_upsample_f32(const float* flow, const float* mask, float* out, int N, int h, int w) {
    const size_t hw = (size_t)h * w;
#pragma omp parallel for collapse(2) schedule(static)
    for (int n = 0; n < N; ++n) {
        for (int y = 0; y < h; ++y) {
            for (int x = 0; x < w; ++x) {
                float nb[2][9];
                for (int c = 0; c < 2; ++c)
                    for (int ky = 0; ky < 3; ++ky)
                        for (int kx = 0; kx < 3; ++kx) {
                            int yy = y + ky - 1, xx = x + kx - 1;
                            float v = 0.0f;
                            if (yy >= 0 && yy < h && xx >= 0 && xx < w)
                                v = 8.0f * flow[((size_t)n * 2 + c) * hw + (size_t)yy * w + xx];
                            nb[c][ky * 3 + kx] = v;
                        }
                for (int i = 0; i < 8; ++i)
                    for (int j = 0; j < 8; ++j) {
                        float m[9], mx = -INFINITY;
                        for (int k = 0; k < 9; ++k) {
                            m[k] = mask[((size_t)n * 576 + k * 64 + i * 8 + j) * hw + (size_t)y * w + x];
                            mx = fmaxf(mx, m[k]);
                        }
                        float s = 0.0f;
                        for (int k = 0; k < 9; ++k) { m[k] = expf(m[k] - mx); s += m[k]; }
                        for (int c = 0; c < 2; ++c) {
                            float acc = 0.0f;
                            for (int k = 0; k < 9; ++k) acc += (m[k] / s) * nb[c][k];
                            out[(((size_t)n * 2 + c) * (8 * h) + (8 * y + i)) * (size_t)(8 * w) + 8 * x + j] = acc;
                        }
                    }
            }
        }
    }
}

/* ------------------------------------------------------------------------------------
 * AverageEndPointError.update (optical_flow/metrics/epe.py:25-35) over dim=1 of
 * (B,2,H,W): epe = sqrt(dx^2 + dy^2) (torch.norm p=2, epe.py:58); pixels with
 * valid >= 0.5 (or all, when valid is NULL) are summed and counted.
 * sum is returned in double (the reference accumulates an fp32 tensor; comparisons use
 * a relative tolerance), count is exact.
 * ---------------------------------------------------------------------------------- */
ORC_API void orc_epe_f32(const float* pred, const float* target, const float* valid_or_null,
                         double* sum_out, int64_t* count_out, int B, int H, int W) {
    const size_t HW = (size_t)H * W;
    double total = 0.0;
    int64_t cnt = 0;
#pragma omp parallel for reduction(+ : total, cnt) schedule(static)
    for (int64_t q = 0; q < (int64_t)B * (int64_t)HW; ++q) {
        size_t b = (size_t)q / HW, p = (size_t)q % HW;
        if (valid_or_null && !(valid_or_null[q] >= 0.5f)) continue;
        float dx = pred[(b * 2 + 0) * HW + p] - target[(b * 2 + 0) * HW + p];
        float dy = pred[(b * 2 + 1) * HW + p] - target[(b * 2 + 1) * HW + p];
        total += (double)sqrtf(dx * dx + dy * dy);
        cnt += 1;
    }
    *sum_out = total;
    *count_out = cnt;
}

/* OutlierRatio.update (optical_flow/metrics/f1.py:33-48): outlier = epe > abs_thr and epe / |target| > rel_thr;
 * masked count of outliers and of selected pixels. */
ORC_API void orc_outlier_f32(const float* pred, const float* target, const float* valid_or_null, double* sum_out,
                             int64_t* count_out, int B, int H, int W, float abs_thr, float rel_thr) {
    const size_t HW = (size_t)H * W;
    double total = 0.0;
    int64_t cnt = 0;
#pragma omp parallel for reduction(+ : total, cnt) schedule(static)
    for (int64_t q = 0; q < (int64_t)B * (int64_t)HW; ++q) {
        size_t b = (size_t)q / HW, p = (size_t)q % HW;
        if (valid_or_null && !(valid_or_null[q] >= 0.5f)) continue;
        float tx = target[(b * 2 + 0) * HW + p], ty = target[(b * 2 + 1) * HW + p];
        float dx = pred[(b * 2 + 0) * HW + p] - tx, dy = pred[(b * 2 + 1) * HW + p] - ty;
        float e = sqrtf(dx * dx + dy * dy);
        float mag = sqrtf(tx * tx + ty * ty);
        if (e > abs_thr && (e / mag) > rel_thr) total += 1.0;
        cnt += 1;
    }
    *sum_out = total;
    *count_out = cnt;
}

/* sequence_loss (methods/raft/model/raft.py:231-260): keep = (valid >= 0.5) & (|gt| < max_flow);
 * loss = sum_i gamma^(n-1-i) * mean(keep * |pred_i - gt|) over all B*2*H*W elements; the metrics are the fractions
 * of kept pixels whose end-point error of the last prediction is below 1 / 3 / 5 px.
 * out[0] = loss, out[1] = sum of kept epe, out[2] = kept count, out[3..5] = counts below 1 / 3 / 5 px. */
ORC_API void orc_sequence_loss_f32(const float* const* preds, int n, const float* gt, const float* valid, double* out,
                                   int B, int H, int W, double gamma, float max_flow) {
    const size_t HW = (size_t)H * W;
    double loss = 0.0, epe = 0.0;
    int64_t keep = 0, n1 = 0, n3 = 0, n5 = 0;
    for (int i = 0; i < n; ++i) {
        double w = 1.0;
        for (int k = 0; k < n - 1 - i; ++k) w *= gamma;
        double total = 0.0;
        const float* pred = preds[i];
        const int last = (i == n - 1);
#pragma omp parallel for reduction(+ : total, epe, keep, n1, n3, n5) schedule(static)
        for (int64_t q = 0; q < (int64_t)B * (int64_t)HW; ++q) {
            size_t b = (size_t)q / HW, p = (size_t)q % HW;
            float gx = gt[(b * 2 + 0) * HW + p], gy = gt[(b * 2 + 1) * HW + p];
            float mag = sqrtf(gx * gx + gy * gy);
            float m = (valid[q] >= 0.5f && mag < max_flow) ? 1.0f : 0.0f;
            float dx = pred[(b * 2 + 0) * HW + p] - gx, dy = pred[(b * 2 + 1) * HW + p] - gy;
            total += (double)(m * fabsf(dx)) + (double)(m * fabsf(dy));
            if (last && m != 0.0f) {
                float e = sqrtf(dx * dx + dy * dy);
                epe += (double)e;
                keep += 1; n1 += e < 1.0f; n3 += e < 3.0f; n5 += e < 5.0f;
            }
        }
        loss += w * (total / ((double)B * 2.0 * (double)HW));
    }
    out[0] = loss; out[1] = epe; out[2] = (double)keep; out[3] = (double)n1; out[4] = (double)n3; out[5] = (double)n5;
}

/* Per-pixel EPE map (end_point_error(reduce=False), epe.py:41-61). */
ORC_API void orc_epe_map_f32(const float* pred, const float* target, float* out, int B, int H, int W) {
    const size_t HW = (size_t)H * W;
#pragma omp parallel for schedule(static)
    for (int64_t q = 0; q < (int64_t)B * (int64_t)HW; ++q) {
        size_t b = (size_t)q / HW, p = (size_t)q % HW;
        float dx = pred[(b * 2 + 0) * HW + p] - target[(b * 2 + 0) * HW + p];
        float dy = pred[(b * 2 + 1) * HW + p] - target[(b * 2 + 1) * HW + p];
        out[q] = sqrtf(dx * dx + dy * dy);
    }
}

/* ------------------------------------------------------------------------------------
 * Round-to-nearest-even fp32 -> bf16 -> fp32 (what the CUDA path does to the feature
 * maps before the tensor-core contraction; lets tests separate quantisation error from
 * kernel error).
 * ---------------------------------------------------------------------------------- */
ORC_API void orc_round_bf16_f32(const float* in, float* out, int64_t n) {
#pragma omp parallel for schedule(static)
    for (int64_t i = 0; i < n; ++i) {
        uint32_t u;
        memcpy(&u, &in[i], 4);
        if ((u & 0x7fffffffu) > 0x7f800000u) { u |= 0x00400000u; u &= 0xffff0000u; }
        else { u += 0x7fffu + ((u >> 16) & 1u); u &= 0xffff0000u; }
        memcpy(&out[i], &u, 4);
    }
}

ORC_API int orc_version(void) { return 1; }
