// bulk_probe.cu -- TMA bulk stores (cp.async.bulk.global.shared::cta) of SMALL pieces: one piece per thread, pieces
// 64 KiB apart in global memory (one query slice each) -- would the K2 epilogue be better off handing each query's
// run to the TMA engine instead of the LSU?
#include <cstdio>
#include <cstdint>
#include <cuda_runtime.h>

__global__ void bulk_write(uint8_t* base, size_t slice, int piece, int total_q, int issuers) {
    extern __shared__ __align__(128) uint8_t sm[];
    for (int i = threadIdx.x; i < issuers * piece / 4; i += blockDim.x) reinterpret_cast<uint32_t*>(sm)[i] = i;
    asm volatile("fence.proxy.async.shared::cta;" ::: "memory");
    __syncthreads();
    if (threadIdx.x >= issuers) return;
    const uint32_t s = (uint32_t)__cvta_generic_to_shared(sm) + threadIdx.x * piece;
    int k = 0;
    for (int q0 = blockIdx.x * issuers; q0 < total_q; q0 += gridDim.x * issuers) {
        uint8_t* dq = base + (size_t)(q0 + threadIdx.x) * slice;
        for (size_t off = 0; off + piece <= slice; off += piece) {
            asm volatile("cp.async.bulk.global.shared::cta.bulk_group [%0], [%1], %2;" ::"l"(dq + off), "r"(s), "r"(piece) : "memory");
            asm volatile("cp.async.bulk.commit_group;" ::: "memory");
            if (++k % 4 == 0) asm volatile("cp.async.bulk.wait_group.read 2;" ::: "memory");
        }
    }
    asm volatile("cp.async.bulk.wait_group 0;" ::: "memory");
}
template <typename F> float timeit(F f, int reps = 3) {
    cudaEvent_t e0, e1; cudaEventCreate(&e0); cudaEventCreate(&e1);
    f();
    cudaEventRecord(e0);
    for (int i = 0; i < reps; ++i) f();
    cudaEventRecord(e1); cudaEventSynchronize(e1);
    float ms; cudaEventElapsedTime(&ms, e0, e1); return ms / reps;
}
int main() {
    const size_t slice = 65536; const int total_q = 32768;       // 2 GiB
    uint8_t* buf; cudaMalloc(&buf, slice * total_q);
    const double gb = (double)slice * total_q / 1e9;
    cudaFuncSetAttribute(bulk_write, cudaFuncAttributeMaxDynamicSharedMemorySize, 128 * 1024);
    for (int piece : {128, 256, 512, 1024, 4096})
        for (int issuers : {32, 128}) {
            if ((size_t)issuers * piece > 128 * 1024) continue;
            float ms = timeit([&] { bulk_write<<<148, 128, issuers * piece>>>(buf, slice, piece, total_q, issuers); });
            printf("{\"piece_bytes\": %d, \"issuing_threads_per_sm\": %d, \"write_gbs\": %.0f}\n", piece, issuers, gb / ms * 1e3);
        }
    printf("%s\n", cudaGetErrorString(cudaDeviceSynchronize()));
    return 0;
}
