// seg_probe.cu -- write bandwidth when every warp store covers several SEGMENTS of `seg` contiguous bytes that lie
// 64 KiB apart (the K2 epilogue pattern: per query row, a short run of the target image).
#include <cstdio>
#include <cstdint>
#include <cuda_runtime.h>

// grid = 148*k CTAs of 256 threads; CTA owns `nq` consecutive "queries" (slices of `slice` bytes) and walks the slice
// in steps of `seg` bytes: at each step all 256 threads together write seg bytes for each of nq queries.
template <int W>
__global__ void seg_write(uint8_t* base, size_t slice, int nq, int seg, int total_q) {
    const int lanes_per_seg = seg / W;                   // threads covering one segment
    const int segs_per_pass = blockDim.x / lanes_per_seg;
    for (int q0 = blockIdx.x * nq; q0 < total_q; q0 += gridDim.x * nq)
        for (size_t off = 0; off + seg <= slice; off += seg)
            for (int s = threadIdx.x / lanes_per_seg; s < nq; s += segs_per_pass) {
                uint8_t* d = base + (size_t)(q0 + s) * slice + off + (size_t)(threadIdx.x % lanes_per_seg) * W;
                if (W == 16) asm volatile("st.global.cs.v4.b32 [%0], {%1, %2, %3, %4};" ::"l"(d), "r"(1), "r"(2), "r"(3), "r"(4) : "memory");
                else asm volatile("st.global.cs.v8.b32 [%0], {%1, %2, %3, %4, %5, %6, %7, %8};" ::"l"(d), "r"(1), "r"(2), "r"(3), "r"(4), "r"(5), "r"(6), "r"(7), "r"(8) : "memory");
            }
}
template <typename F> float timeit(F f, int reps = 5) {
    cudaEvent_t e0, e1; cudaEventCreate(&e0); cudaEventCreate(&e1);
    for (int i = 0; i < 2; ++i) f();
    cudaEventRecord(e0);
    for (int i = 0; i < reps; ++i) f();
    cudaEventRecord(e1); cudaEventSynchronize(e1);
    float ms; cudaEventElapsedTime(&ms, e0, e1); return ms / reps;
}
int main() {
    const size_t slice = 65536;                 // one query's level-0 slice at 1080p (136 x 240 bf16)
    const int total_q = 65536;                  // 4 GiB
    uint8_t* buf; cudaMalloc(&buf, slice * total_q);
    const double gb = (double)slice * total_q / 1e9;
    for (int nq : {128})
        for (int seg : {64, 128, 256, 1024})
            for (int per_sm : {1}) {
                float ms = timeit([&] { seg_write<16><<<148 * per_sm, 256>>>(buf, slice, nq, seg, total_q); });
                float ms8 = timeit([&] { seg_write<32><<<148 * per_sm, 256>>>(buf, slice, nq, seg, total_q); });
                printf("{\"queries_per_cta\": %d, \"seg_bytes\": %d, \"ctas_per_sm\": %d, \"v4_write_gbs\": %.0f, \"v8_write_gbs\": %.0f}\n", nq, seg, per_sm, gb / ms * 1e3, gb / ms8 * 1e3);
            }
    printf("%s\n", cudaGetErrorString(cudaDeviceSynchronize()));
    return 0;
}
