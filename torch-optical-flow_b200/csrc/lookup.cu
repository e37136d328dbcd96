// lookup.cu -- K3: correlation-pyramid lookup, all levels in one pass.
//
// Replaces CorrBlock.__call__ + bilinear_sampler (reference methods/raft/model/corr.py:56-77,
// methods/raft/model/utils.py:64-80).  The reference runs ~70 small ATen launches per
// refinement iteration (delta grid + H2D copy, 3 materialised (N,9,9,2) coordinate tensors,
// grid_sample, 5 concatenations); here one kernel reads the coordinates and the pyramid and
// writes the (B, L*(2r+1)^2, h, w) fp32 output.
//
// Bit-exactness contract: the fp32 coordinate sequence is replayed operation
// by operation with round-to-nearest intrinsics so nvcc cannot contract it:
//     c  = coord / 2^l                       (corr.py:68, exact)
//     x  = c + (i - r)                       (corr.py:70; window TRANSPOSED: i moves x, j moves y)
//     g  = 2*x/(W_l-1) - 1                   (utils.py:70-71)
//     ix = ((g + 1)/2) * (W_l-1)             (ATen GridSampler.h:30, align_corners=True)
//     x0 = floor(ix); valid = -1 < g < 1     (utils.py:77)
//
// Work decomposition: a CTA owns 32 consecutive queries; each of its 8 warps walks 4 of them.
// Per (query, level) the warp
//   1. computes the 2*(2r+1) tap coordinates, one per lane (lanes 0..8: x taps, 16..24: y taps),
//   2. loads the <=12x12 patch of the query's level slice that covers all taps into a private
//      shared-memory patch (4-byte words; out-of-image elements become zeros = zeros padding),
//   3. produces the (2r+1)^2 samples, 3 per lane, fetching the per-axis floor index / weight of
//      its cell from the owning lanes with warp shuffles.
// Results are staged in a [channel][query] shared tile so the global writes are 128-byte rows.
// HBM roofline per query and iteration: L*(2r+2)^2*esize + 8 read, 4*L*(2r+1)^2 written.
#include "common.cuh"

namespace {

constexpr int QT = 32;        // queries per CTA
constexpr int NW = 8;         // warps per CTA
constexpr int PD = 12;        // patch rows / cols
constexpr int OT_PITCH = QT + 1;
constexpr int MAX_D = 9;      // 2*radius+1, radius <= 4

struct LookupParams {
    const void* base[OFB_MAX_LEVELS];
    long long q_stride[OFB_MAX_LEVELS];
    int pitch[OFB_MAX_LEVELS];
    int lh[OFB_MAX_LEVELS];
    int lw[OFB_MAX_LEVELS];
    int levels, radius, B, h, w;
};

template <typename T> struct Elem;
template <> struct Elem<float> {
    static __device__ __forceinline__ float get(const float* row, int x) { return __ldg(row + x); }
};
template <> struct Elem<__nv_bfloat16> {
    static __device__ __forceinline__ float get(const __nv_bfloat16* row, int x) {
        return __bfloat162float(__ldg(row + x));
    }
};

// patch load: fp32 -> one element per word; bf16 -> two elements per 4-byte word
template <typename T>
__device__ __forceinline__ void load_patch(float* patch, const T* slice, int pitch, int Hl, int Wl, int px0, int py0,
                                           int rows, int lane);

template <>
__device__ __forceinline__ void load_patch<float>(float* patch, const float* slice, int pitch, int Hl, int Wl, int px0,
                                                  int py0, int rows, int lane) {
#pragma unroll
    for (int e = lane; e < PD * PD; e += 32) {
        const int r = e / PD, c = e - r * PD;
        const int y = py0 + r, x = px0 + c;
        float v = 0.0f;
        if (r < rows && y >= 0 && y < Hl && x >= 0 && x < Wl) v = __ldg(slice + (size_t)y * pitch + x);
        patch[e] = v;
    }
}

template <>
__device__ __forceinline__ void load_patch<__nv_bfloat16>(float* patch, const __nv_bfloat16* slice, int pitch, int Hl,
                                                          int Wl, int px0, int py0, int rows, int lane) {
    // px0 is even, pitch is even, the slice base is 4-byte aligned
#pragma unroll
    for (int e = lane; e < PD * (PD / 2); e += 32) {
        const int r = e / (PD / 2), cw = e - r * (PD / 2);
        const int y = py0 + r, x = px0 + 2 * cw;
        float v0 = 0.0f, v1 = 0.0f;
        if (r < rows && y >= 0 && y < Hl && x >= 0 && x < Wl) {
            const unsigned wd = __ldg(reinterpret_cast<const unsigned*>(slice + (size_t)y * pitch + x));
            v0 = __uint_as_float(wd << 16);
            if (x + 1 < Wl) v1 = __uint_as_float(wd & 0xffff0000u);
        }
        patch[r * PD + 2 * cw] = v0;
        patch[r * PD + 2 * cw + 1] = v1;
    }
}

template <typename T>
__global__ void __launch_bounds__(NW * 32) lookup_kernel(const LookupParams P, const float* __restrict__ coords,
                                                         float* __restrict__ out, int32_t* __restrict__ idx_out,
                                                         uint8_t* __restrict__ valid_out) {
    extern __shared__ float smem[];
    const int D = 2 * P.radius + 1, DD = D * D, CH = P.levels * DD;
    float* otile = smem;                              // [CH][OT_PITCH]
    float* patches = smem + (size_t)CH * OT_PITCH;    // [NW][PD*PD]
    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    float* patch = patches + warp * PD * PD;
    const long long HW = (long long)P.h * P.w, Q = (long long)P.B * HW;
    const long long q0 = (long long)blockIdx.x * QT;
    constexpr bool kBf16 = sizeof(T) == 2;

    for (int qi = warp; qi < QT; qi += NW) {
        const long long q = q0 + qi;
        if (q >= Q) break;
        const long long b = q / HW, p = q - b * HW;
        const float cx0 = __ldg(coords + (b * 2 + 0) * HW + p);
        const float cy0 = __ldg(coords + (b * 2 + 1) * HW + p);
        for (int l = 0; l < P.levels; ++l) {
            const int Wl = P.lw[l], Hl = P.lh[l], pitch = P.pitch[l];
            const T* slice = reinterpret_cast<const T*>(P.base[l]) + q * P.q_stride[l];
            const float inv = 1.0f / (float)(1 << l);          // exact power of two
            // ---- 1. tap coordinates: lane t (x taps) and lane 16+t (y taps)
            const int t = lane & 15, isy = lane >> 4;
            const float cen = __fmul_rn(isy ? cy0 : cx0, inv);
            const int size = isy ? Hl : Wl;
            const float pos = __fadd_rn(cen, (float)(t - P.radius));
            const float sm1 = (float)(size - 1);
            const float g = __fsub_rn(__fdiv_rn(__fmul_rn(2.0f, pos), sm1), 1.0f);
            const float ic = __fmul_rn(__fdiv_rn(__fadd_rn(g, 1.0f), 2.0f), sm1);
            const float fl = floorf(ic);
            float w1 = __fsub_rn(ic, fl), w0 = __fsub_rn(__fadd_rn(fl, 1.0f), ic);
            // int conversion guarded: non-finite / far-away coordinates map to "everything out of range"
            int i0 = (fl >= -8.0f && fl <= 70000.0f) ? (int)fl : -1000000;
            if (i0 == -1000000) { w0 = 0.0f; w1 = 0.0f; }
            const int gv = (g > -1.0f) && (g < 1.0f);
            if (idx_out && t < D)
                idx_out[((q * P.levels + l) * 2 + isy) * D + t] = (int32_t)fl;
            const int x_first = __shfl_sync(0xffffffffu, i0, 0), x_last = __shfl_sync(0xffffffffu, i0, D - 1);
            const int y_first = __shfl_sync(0xffffffffu, i0, 16), y_last = __shfl_sync(0xffffffffu, i0, 16 + D - 1);
            const int px0 = kBf16 ? (x_first & ~1) : x_first;
            const int py0 = y_first;
            const int cols = x_last + 2 - px0, rows = y_last + 2 - py0;
            const bool patch_ok = cols <= PD && rows <= PD && cols > 0 && rows > 0;   // warp-uniform
            // ---- 2. stage the patch
            if (patch_ok) load_patch<T>(patch, slice, pitch, Hl, Wl, px0, py0, rows, lane);
            __syncwarp();
            // ---- 3. (2r+1)^2 samples, 3 per lane
#pragma unroll
            for (int kk = 0; kk < (MAX_D * MAX_D + 31) / 32; ++kk) {
                const int k = kk * 32 + lane;
                const bool active = k < DD;
                const int i = active ? k / D : 0, j = active ? k - (k / D) * D : 0;
                const int x0 = __shfl_sync(0xffffffffu, i0, i), y0 = __shfl_sync(0xffffffffu, i0, 16 + j);
                const float wx0 = __shfl_sync(0xffffffffu, w0, i), wx1 = __shfl_sync(0xffffffffu, w1, i);
                const float wy0 = __shfl_sync(0xffffffffu, w0, 16 + j), wy1 = __shfl_sync(0xffffffffu, w1, 16 + j);
                const int vx = __shfl_sync(0xffffffffu, gv, i), vy = __shfl_sync(0xffffffffu, gv, 16 + j);
                if (!active) continue;
                float v00, v01, v10, v11;
                if (patch_ok) {
                    const float* s = patch + (y0 - py0) * PD + (x0 - px0);
                    v00 = s[0]; v01 = s[1]; v10 = s[PD]; v11 = s[PD + 1];
                } else {   // generic path: taps straight from global memory
                    const bool inx0 = x0 >= 0 && x0 < Wl, inx1 = x0 + 1 >= 0 && x0 + 1 < Wl;
                    const bool iny0 = y0 >= 0 && y0 < Hl, iny1 = y0 + 1 >= 0 && y0 + 1 < Hl;
                    v00 = (iny0 && inx0) ? Elem<T>::get(slice + (size_t)y0 * pitch, x0) : 0.0f;
                    v01 = (iny0 && inx1) ? Elem<T>::get(slice + (size_t)y0 * pitch, x0 + 1) : 0.0f;
                    v10 = (iny1 && inx0) ? Elem<T>::get(slice + (size_t)(y0 + 1) * pitch, x0) : 0.0f;
                    v11 = (iny1 && inx1) ? Elem<T>::get(slice + (size_t)(y0 + 1) * pitch, x0 + 1) : 0.0f;
                }
                float acc = __fmul_rn(v00, __fmul_rn(wx0, wy0));
                acc = __fmaf_rn(v01, __fmul_rn(wx1, wy0), acc);
                acc = __fmaf_rn(v10, __fmul_rn(wx0, wy1), acc);
                acc = __fmaf_rn(v11, __fmul_rn(wx1, wy1), acc);
                otile[(l * DD + k) * OT_PITCH + qi] = acc;
                if (valid_out) valid_out[(q * P.levels + l) * DD + k] = (uint8_t)(vx & vy);
            }
            __syncwarp();
        }
    }
    __syncthreads();
    // ---- coalesced write-out: one channel row (32 queries) per warp instruction
    const long long q = q0 + lane;
    if (q < Q) {
        const long long b = q / HW, p = q - b * HW;
        float* dst = out + b * CH * HW + p;
        for (int ch = warp; ch < CH; ch += NW) dst[(long long)ch * HW] = otile[ch * OT_PITCH + lane];
    }
}

__global__ void __launch_bounds__(256) bilinear_sampler_kernel(const float* __restrict__ img,
                                                               const float* __restrict__ coords,
                                                               float* __restrict__ out, float* __restrict__ mask, int N,
                                                               int C, int H, int W, int Ho, int Wo) {
    const size_t HWo = (size_t)Ho * Wo, total = (size_t)N * HWo;
    for (size_t t = (size_t)blockIdx.x * blockDim.x + threadIdx.x; t < total; t += (size_t)gridDim.x * blockDim.x) {
        const size_t n = t / HWo, p = t - n * HWo;
        const float2 c = __ldg(reinterpret_cast<const float2*>(coords) + t);
        const float gx = __fsub_rn(__fdiv_rn(__fmul_rn(2.0f, c.x), (float)(W - 1)), 1.0f);
        const float gy = __fsub_rn(__fdiv_rn(__fmul_rn(2.0f, c.y), (float)(H - 1)), 1.0f);
        const float ix = ofb::unnormalize<true>(gx, W), iy = ofb::unnormalize<true>(gy, H);
        if (mask) mask[t] = ((gx > -1.0f) && (gy > -1.0f) && (gx < 1.0f) && (gy < 1.0f)) ? 1.0f : 0.0f;
        const float x0f = floorf(ix), y0f = floorf(iy);
        const bool fin = x0f >= -8.0f && x0f <= 1.0e6f && y0f >= -8.0f && y0f <= 1.0e6f;
        const int x0 = fin ? (int)x0f : -100, y0 = fin ? (int)y0f : -100;
        const float wx1 = ix - x0f, wx0 = (x0f + 1.0f) - ix, wy1 = iy - y0f, wy0 = (y0f + 1.0f) - iy;
        const bool inx0 = x0 >= 0 && x0 < W, inx1 = x0 + 1 >= 0 && x0 + 1 < W;
        const bool iny0 = y0 >= 0 && y0 < H, iny1 = y0 + 1 >= 0 && y0 + 1 < H;
        for (int ch = 0; ch < C; ++ch) {
            const float* plane = img + ((size_t)n * C + ch) * H * W;
            float acc = 0.0f;
            if (iny0 && inx0) acc = __fmaf_rn(__ldg(plane + (size_t)y0 * W + x0), wx0 * wy0, acc);
            if (iny0 && inx1) acc = __fmaf_rn(__ldg(plane + (size_t)y0 * W + x0 + 1), wx1 * wy0, acc);
            if (iny1 && inx0) acc = __fmaf_rn(__ldg(plane + (size_t)(y0 + 1) * W + x0), wx0 * wy1, acc);
            if (iny1 && inx1) acc = __fmaf_rn(__ldg(plane + (size_t)(y0 + 1) * W + x0 + 1), wx1 * wy1, acc);
            out[((size_t)n * C + ch) * HWo + p] = acc;
        }
    }
}

}  // namespace

OFB_API int ofb_corr_lookup(const ofb_pyramid* pyr, const float* coords, float* out, int32_t* idx_or_null,
                            uint8_t* valid_or_null, int B, int h, int w, int radius, void* stream) {
    if (!pyr || !coords || !out || B < 0 || h <= 0 || w <= 0) return OFB_EINVAL;
    if (pyr->levels < 1 || pyr->levels > OFB_MAX_LEVELS) return OFB_EINVAL;
    if (radius < 0 || 2 * radius + 1 > MAX_D) return OFB_EUNSUPPORTED;
    if (pyr->dtype != OFB_DTYPE_F32 && pyr->dtype != OFB_DTYPE_BF16) return OFB_EINVAL;
    if (B == 0) return OFB_OK;
    LookupParams P;
    P.levels = pyr->levels; P.radius = radius; P.B = B; P.h = h; P.w = w;
    for (int l = 0; l < OFB_MAX_LEVELS; ++l) {
        const bool on = l < pyr->levels;
        P.base[l] = on ? pyr->base[l] : nullptr;
        P.q_stride[l] = on ? pyr->q_stride[l] : 0;
        P.pitch[l] = on ? pyr->row_pitch[l] : 0;
        P.lh[l] = on ? pyr->lvl_h[l] : 0;
        P.lw[l] = on ? pyr->lvl_w[l] : 0;
        if (on) {
            if (!P.base[l] || P.lh[l] <= 0 || P.lw[l] <= 0 || P.pitch[l] < P.lw[l]) return OFB_EINVAL;
            if (P.lh[l] > 65536 || P.lw[l] > 65536) return OFB_EUNSUPPORTED;
            if (pyr->dtype == OFB_DTYPE_BF16 &&
                ((P.pitch[l] & 1) || (P.q_stride[l] & 1) || (reinterpret_cast<uintptr_t>(P.base[l]) & 3)))
                return OFB_EALIGN;
        }
    }
    const int D = 2 * radius + 1, CH = pyr->levels * D * D;
    const size_t smem = ((size_t)CH * OT_PITCH + (size_t)NW * PD * PD) * sizeof(float);
    const long long Q = (long long)B * h * w;
    const long long blocks = (Q + QT - 1) / QT;
    if (blocks > 0x7fffffffLL) return OFB_EUNSUPPORTED;
    cudaStream_t st = (cudaStream_t)stream;
    if (pyr->dtype == OFB_DTYPE_BF16) {
        OFB_CUDA(cudaFuncSetAttribute(lookup_kernel<__nv_bfloat16>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
        lookup_kernel<__nv_bfloat16><<<(int)blocks, NW * 32, smem, st>>>(P, coords, out, idx_or_null, valid_or_null);
    } else {
        OFB_CUDA(cudaFuncSetAttribute(lookup_kernel<float>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
        lookup_kernel<float><<<(int)blocks, NW * 32, smem, st>>>(P, coords, out, idx_or_null, valid_or_null);
    }
    OFB_LAUNCH_CHECK();
    return OFB_OK;
}

OFB_API int ofb_bilinear_sampler_f32(const float* img, const float* coords, float* out, float* mask_or_null, int N, int C,
                                     int H, int W, int Ho, int Wo, void* stream) {
    if (!img || !coords || !out || N < 0 || C < 0 || H <= 0 || W <= 0 || Ho < 0 || Wo < 0) return OFB_EINVAL;
    const size_t total = (size_t)N * Ho * Wo;
    if (total == 0) return OFB_OK;
    int blocks = (int)((total + 255) / 256);
    const int cap = ofb_num_sms() * 32;
    if (blocks > cap) blocks = cap;
    bilinear_sampler_kernel<<<blocks, 256, 0, (cudaStream_t)stream>>>(img, coords, out, mask_or_null, N, C, H, W, Ho, Wo);
    OFB_LAUNCH_CHECK();
    return OFB_OK;
}
