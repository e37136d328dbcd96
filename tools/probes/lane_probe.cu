// lane_probe.cu -- write bandwidth when every LANE owns its own slice (the tcgen05.ld register layout: thread = query
// row) and writes `run` contiguous bytes per visit with 16-byte or 32-byte stores -- i.e. stores straight from the
// accumulator registers without a shared-memory transpose.
#include <cstdio>
#include <cstdint>
#include <cuda_runtime.h>

template <int W>   // W = 16: st.v4.b32, W = 32: st.v8.b32
__global__ void lane_write(uint8_t* base, size_t slice, int run, int total_q) {
    for (int q = blockIdx.x * blockDim.x + threadIdx.x; q < total_q; q += gridDim.x * blockDim.x) {
        uint8_t* s = base + (size_t)q * slice;
        for (size_t off = 0; off + run <= slice; off += run)
#pragma unroll 4
            for (int b = 0; b < run; b += W) {
                uint8_t* d = s + off + b;
                if (W == 16) asm volatile("st.global.cs.v4.b32 [%0], {%1, %2, %3, %4};" ::"l"(d), "r"(1), "r"(2), "r"(3), "r"(4) : "memory");
                else asm volatile("st.global.cs.v8.b32 [%0], {%1, %2, %3, %4, %5, %6, %7, %8};" ::"l"(d), "r"(1), "r"(2), "r"(3), "r"(4), "r"(5), "r"(6), "r"(7), "r"(8) : "memory");
            }
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
    const size_t slice = 65536; const int total_q = 65536;
    uint8_t* buf; cudaMalloc(&buf, slice * total_q);
    const double gb = (double)slice * total_q / 1e9;
    for (int run : {64, 128, 256})
        for (int threads : {128, 256}) {
            float a = timeit([&] { lane_write<16><<<148, threads>>>(buf, slice, run, total_q); });
            float b = timeit([&] { lane_write<32><<<148, threads>>>(buf, slice, run, total_q); });
            printf("{\"run_bytes\": %d, \"threads_per_sm\": %d, \"v4_gbs\": %.0f, \"v8_gbs\": %.0f}\n", run, threads, gb / a * 1e3, gb / b * 1e3);
        }
    printf("%s\n", cudaGetErrorString(cudaDeviceSynchronize()));
    return 0;
}
