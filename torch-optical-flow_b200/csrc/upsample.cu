// upsample.cu -- K4b: RAFT convex 8x flow upsampling in one pass.
//
// Replaces RAFT.upsample_flow (reference methods/raft/model/raft.py:73-85):
//   out[n,c,8y+i,8x+j] = sum_k softmax_k(mask[n, k*64 + i*8 + j, y, x]) * 8*flow_pad[n,c,y+ky-1,x+kx-1]
// with k = ky*3+kx over the zero-padded 3x3 neighbourhood.  The reference materialises the
// softmax (B,1,9,8,8,h,w), the unfolded flow (9x) and their product (9x the output); here
// nothing but the 576-channel mask is read and the upsampled flow written.
//
// Mapping: thread = (coarse pixel p = y*w+x, sub-row i).  Mask planes are contiguous over p,
// so every mask load of a warp is a full 128-byte line; each thread produces the eight
// consecutive outputs j = 0..7 of fine row 8y+i (two 16-byte stores per flow component).
// HBM roofline: 4*(576 + 2) bytes read + 4*128 bytes written per coarse pixel.
#include "common.cuh"

namespace {

constexpr int PX = 32;   // coarse pixels per CTA
constexpr int NT = PX * 8;

__global__ void __launch_bounds__(NT) convex_upsample_kernel(const float* __restrict__ flow,
                                                             const float* __restrict__ mask,
                                                             float* __restrict__ out, int N, int h, int w) {
    const int hw = h * w;
    const int lane_p = threadIdx.x % PX;
    const int i = threadIdx.x / PX;                     // sub-row 0..7 (warp-uniform)
    const int n = blockIdx.y;
    const int p = blockIdx.x * PX + lane_p;
    if (p >= hw) return;
    const int y = p / w, x = p - y * w;

    // 3x3 neighbourhood of 8*flow, zero padded (F.unfold(8*flow, 3, padding=1))
    float nbx[9], nby[9];
    const float* fx = flow + (size_t)n * 2 * hw;
    const float* fy = fx + hw;
#pragma unroll
    for (int ky = 0; ky < 3; ++ky)
#pragma unroll
        for (int kx = 0; kx < 3; ++kx) {
            const int yy = y + ky - 1, xx = x + kx - 1;
            const bool in = yy >= 0 && yy < h && xx >= 0 && xx < w;
            nbx[ky * 3 + kx] = in ? 8.0f * __ldg(fx + yy * w + xx) : 0.0f;
            nby[ky * 3 + kx] = in ? 8.0f * __ldg(fy + yy * w + xx) : 0.0f;
        }

    const float* mp = mask + (size_t)n * 576 * hw + (size_t)(i * 8) * hw + p;
    float ox[8], oy[8];
    // JG sub-columns at a time: 9*JG independent 128-byte-coalesced loads are in flight per warp before
    // the first softmax starts (the kernel is a pure stream of the 576-channel mask: latency, not math)
    constexpr int JG = 4;
#pragma unroll
    for (int jg = 0; jg < 8; jg += JG) {
        float m[JG][9];
#pragma unroll
        for (int jj = 0; jj < JG; ++jj)
#pragma unroll
            for (int k = 0; k < 9; ++k) m[jj][k] = __ldg(mp + (size_t)(k * 64 + jg + jj) * hw);
#pragma unroll
        for (int jj = 0; jj < JG; ++jj) {
            float mx = m[jj][0];
#pragma unroll
            for (int k = 1; k < 9; ++k) mx = fmaxf(mx, m[jj][k]);
            float s = 0.0f;
#pragma unroll
            for (int k = 0; k < 9; ++k) { m[jj][k] = expf(m[jj][k] - mx); s += m[jj][k]; }
            // softmax-weighted sums, normalised once at the end: sum_k e_k v_k / sum_k e_k (one division per
            // component instead of nine; differs from the reference's per-weight division by <= 2 ulp)
            float ax = 0.0f, ay = 0.0f;
#pragma unroll
            for (int k = 0; k < 9; ++k) {
                ax = __fmaf_rn(m[jj][k], nbx[k], ax);
                ay = __fmaf_rn(m[jj][k], nby[k], ay);
            }
            const float inv = __fdiv_rn(1.0f, s);
            ox[jg + jj] = ax * inv; oy[jg + jj] = ay * inv;
        }
    }
    const size_t W8 = (size_t)8 * w, H8 = (size_t)8 * h;
    float* dx = out + ((size_t)n * 2 + 0) * H8 * W8 + (size_t)(8 * y + i) * W8 + 8 * x;
    float* dy = dx + H8 * W8;
    // rows are 8*w floats and x offsets multiples of 8 -> 32-byte aligned
    reinterpret_cast<float4*>(dx)[0] = make_float4(ox[0], ox[1], ox[2], ox[3]);
    reinterpret_cast<float4*>(dx)[1] = make_float4(ox[4], ox[5], ox[6], ox[7]);
    reinterpret_cast<float4*>(dy)[0] = make_float4(oy[0], oy[1], oy[2], oy[3]);
    reinterpret_cast<float4*>(dy)[1] = make_float4(oy[4], oy[5], oy[6], oy[7]);
}

}  // namespace

OFB_API int ofb_convex_upsample_f32(const float* flow, const float* mask, float* out, int N, int h, int w, void* stream) {
    if (!flow || !mask || !out || N < 0 || h < 0 || w < 0) return OFB_EINVAL;
    if ((size_t)N * h * w == 0) return OFB_OK;
    if (N > 65535) return OFB_EUNSUPPORTED;
    if (reinterpret_cast<uintptr_t>(out) & 15) return OFB_EALIGN;
    dim3 grid((h * w + PX - 1) / PX, N);
    convex_upsample_kernel<<<grid, NT, 0, (cudaStream_t)stream>>>(flow, mask, out, N, h, w);
    OFB_LAUNCH_CHECK();
    return OFB_OK;
}
