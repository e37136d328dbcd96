// epe.cu -- K4c: end-point-error map, masked EPE sum/count and outlier (F1) count reductions.
//
// Replaces torch.norm(pred-target, p=2, dim=1) + boolean select + sum/numel in
// AverageEndPointError.update (reference optical_flow/metrics/epe.py:25-35,58).
// One streaming pass: float4 loads of both flow components, per-thread fp64 partial sums,
// warp shuffle -> shared -> ONE atomicAdd(double) pair per CTA into acc[2] = {sum, count}.
// The cross-rank reduction (dist_reduce_fx="sum", epe.py:22-23) is a single all-reduce of
// that 16-byte buffer, issued by the host (see optical_flow/metrics/epe.py in this repo).
// HBM roofline: 16 bytes per pixel (+4 with a validity map).
#include "common.cuh"

namespace {

constexpr int NT = 256;

__device__ __forceinline__ float epe1(float px, float py, float tx, float ty) {
    const float dx = __fsub_rn(px, tx), dy = __fsub_rn(py, ty);
    return sqrtf(__fadd_rn(__fmul_rn(dx, dx), __fmul_rn(dy, dy)));
}

// per-pixel contribution: MODE 0 = the end-point error (epe.py:28-35), MODE 1 = 1 if the pixel is an outlier,
// epe > abs_thr and epe / |target| > rel_thr (reference optical_flow/metrics/f1.py:36-41), else 0
template <int MODE>
__device__ __forceinline__ float contribution(float px, float py, float tx, float ty, float abs_thr, float rel_thr) {
    const float e = epe1(px, py, tx, ty);
    if (MODE == 0) return e;
    const float mag = sqrtf(__fadd_rn(__fmul_rn(tx, tx), __fmul_rn(ty, ty)));
    return (e > abs_thr && __fdiv_rn(e, mag) > rel_thr) ? 1.0f : 0.0f;
}

template <int MODE>
__global__ void __launch_bounds__(NT) epe_reduce_kernel(const float* __restrict__ pred, const float* __restrict__ target,
                                                        const float* __restrict__ valid, double* __restrict__ acc,
                                                        int B, int64_t HW, int vec, float abs_thr, float rel_thr) {
    double sum = 0.0;
    unsigned long long cnt = 0;
    if (vec) {
        const int64_t HW4 = HW >> 2;
        const int64_t total = (int64_t)B * HW4;
        for (int64_t t = (int64_t)blockIdx.x * NT + threadIdx.x; t < total; t += (int64_t)gridDim.x * NT) {
            const int64_t b = t / HW4, q = t - b * HW4;
            const float4 px = __ldg(reinterpret_cast<const float4*>(pred + (b * 2 + 0) * HW) + q);
            const float4 py = __ldg(reinterpret_cast<const float4*>(pred + (b * 2 + 1) * HW) + q);
            const float4 tx = __ldg(reinterpret_cast<const float4*>(target + (b * 2 + 0) * HW) + q);
            const float4 ty = __ldg(reinterpret_cast<const float4*>(target + (b * 2 + 1) * HW) + q);
            float4 v = make_float4(1.f, 1.f, 1.f, 1.f);
            if (valid) v = __ldg(reinterpret_cast<const float4*>(valid + b * HW) + q);
            if (v.x >= 0.5f) { sum += (double)contribution<MODE>(px.x, py.x, tx.x, ty.x, abs_thr, rel_thr); ++cnt; }
            if (v.y >= 0.5f) { sum += (double)contribution<MODE>(px.y, py.y, tx.y, ty.y, abs_thr, rel_thr); ++cnt; }
            if (v.z >= 0.5f) { sum += (double)contribution<MODE>(px.z, py.z, tx.z, ty.z, abs_thr, rel_thr); ++cnt; }
            if (v.w >= 0.5f) { sum += (double)contribution<MODE>(px.w, py.w, tx.w, ty.w, abs_thr, rel_thr); ++cnt; }
        }
    } else {
        const int64_t total = (int64_t)B * HW;
        for (int64_t t = (int64_t)blockIdx.x * NT + threadIdx.x; t < total; t += (int64_t)gridDim.x * NT) {
            const int64_t b = t / HW, q = t - b * HW;
            if (valid && !(__ldg(valid + t) >= 0.5f)) continue;
            sum += (double)contribution<MODE>(__ldg(pred + (b * 2 + 0) * HW + q), __ldg(pred + (b * 2 + 1) * HW + q),
                                              __ldg(target + (b * 2 + 0) * HW + q), __ldg(target + (b * 2 + 1) * HW + q),
                                              abs_thr, rel_thr);
            ++cnt;
        }
    }
    __shared__ double s_sum[NT / 32];
    __shared__ double s_cnt[NT / 32];
    double c = (double)cnt;
    sum = ofb::warp_sum(sum);
    c = ofb::warp_sum(c);
    if ((threadIdx.x & 31) == 0) { s_sum[threadIdx.x >> 5] = sum; s_cnt[threadIdx.x >> 5] = c; }
    __syncthreads();
    if (threadIdx.x == 0) {
        double ts = 0.0, tc = 0.0;
#pragma unroll
        for (int k = 0; k < NT / 32; ++k) { ts += s_sum[k]; tc += s_cnt[k]; }
        atomicAdd(acc + 0, ts);
        atomicAdd(acc + 1, tc);
    }
}

__global__ void __launch_bounds__(NT) epe_map_kernel(const float* __restrict__ pred, const float* __restrict__ target,
                                                     float* __restrict__ out, int B, int64_t HW) {
    const int64_t total = (int64_t)B * HW;
    for (int64_t t = (int64_t)blockIdx.x * NT + threadIdx.x; t < total; t += (int64_t)gridDim.x * NT) {
        const int64_t b = t / HW, q = t - b * HW;
        out[t] = epe1(__ldg(pred + (b * 2 + 0) * HW + q), __ldg(pred + (b * 2 + 1) * HW + q),
                      __ldg(target + (b * 2 + 0) * HW + q), __ldg(target + (b * 2 + 1) * HW + q));
    }
}

}  // namespace

namespace {
template <int MODE>
int launch_reduce(const float* pred, const float* target, const float* valid_or_null, double* acc, int B, int H, int W,
                  float abs_thr, float rel_thr, void* stream) {
    if (!pred || !target || !acc || B < 0 || H < 0 || W < 0) return OFB_EINVAL;
    const int64_t HW = (int64_t)H * W;
    if ((int64_t)B * HW == 0) return OFB_OK;
    const uintptr_t al = reinterpret_cast<uintptr_t>(pred) | reinterpret_cast<uintptr_t>(target) |
                         reinterpret_cast<uintptr_t>(valid_or_null);
    const int vec = (HW % 4 == 0) && ((al & 15) == 0);
    const int64_t work = vec ? (int64_t)B * (HW / 4) : (int64_t)B * HW;
    int64_t blocks = (work + NT - 1) / NT;
    const int cap = ofb_num_sms() * 8;   // few CTAs -> few atomics; 8 x 256 threads fill an SM
    if (blocks > cap) blocks = cap;
    epe_reduce_kernel<MODE><<<(int)blocks, NT, 0, (cudaStream_t)stream>>>(pred, target, valid_or_null, acc, B, HW, vec,
                                                                         abs_thr, rel_thr);
    OFB_LAUNCH_CHECK();
    return OFB_OK;
}
}  // namespace

OFB_API int ofb_epe_reduce_f32(const float* pred, const float* target, const float* valid_or_null, double* acc, int B,
                               int H, int W, void* stream) {
    return launch_reduce<0>(pred, target, valid_or_null, acc, B, H, W, 0.0f, 0.0f, stream);
}

OFB_API int ofb_outlier_reduce_f32(const float* pred, const float* target, const float* valid_or_null, double* acc, int B,
                                   int H, int W, float abs_threshold, float rel_threshold, void* stream) {
    return launch_reduce<1>(pred, target, valid_or_null, acc, B, H, W, abs_threshold, rel_threshold, stream);
}

OFB_API int ofb_epe_map_f32(const float* pred, const float* target, float* out, int B, int H, int W, void* stream) {
    if (!pred || !target || !out || B < 0 || H < 0 || W < 0) return OFB_EINVAL;
    const int64_t HW = (int64_t)H * W;
    if ((int64_t)B * HW == 0) return OFB_OK;
    int64_t blocks = ((int64_t)B * HW + NT - 1) / NT;
    const int cap = ofb_num_sms() * 32;
    if (blocks > cap) blocks = cap;
    epe_map_kernel<<<(int)blocks, NT, 0, (cudaStream_t)stream>>>(pred, target, out, B, HW);
    OFB_LAUNCH_CHECK();
    return OFB_OK;
}
