// write_probe.cu -- how fast can one B200 stream WRITES to HBM, by store flavour?
// Build: nvcc -O3 -gencode arch=compute_100a,code=sm_100a -o write_probe write_probe.cu ; run: ./write_probe
#include <cstdio>
#include <cstdint>
#include <cuda_runtime.h>

__global__ void st_plain(uint4* p, size_t n) {
    for (size_t i = (size_t)blockIdx.x * blockDim.x + threadIdx.x; i < n; i += (size_t)gridDim.x * blockDim.x)
        p[i] = make_uint4(1, 2, 3, 4);
}
__global__ void st_cs(uint4* p, size_t n) {
    for (size_t i = (size_t)blockIdx.x * blockDim.x + threadIdx.x; i < n; i += (size_t)gridDim.x * blockDim.x)
        asm volatile("st.global.cs.v4.b32 [%0], {%1, %2, %3, %4};" ::"l"(p + i), "r"(1), "r"(2), "r"(3), "r"(4) : "memory");
}
__global__ void st_wt(uint4* p, size_t n) {
    for (size_t i = (size_t)blockIdx.x * blockDim.x + threadIdx.x; i < n; i += (size_t)gridDim.x * blockDim.x)
        asm volatile("st.global.wt.v4.b32 [%0], {%1, %2, %3, %4};" ::"l"(p + i), "r"(1), "r"(2), "r"(3), "r"(4) : "memory");
}
// each CTA owns contiguous 64 KiB blocks (DRAM-page friendly) instead of a grid-stride interleave
__global__ void st_blocked(uint4* p, size_t n, int cs) {
    const size_t per = 4096;  // uint4 per block = 64 KiB
    for (size_t blk = blockIdx.x; blk * per < n; blk += gridDim.x)
        for (size_t i = threadIdx.x; i < per && blk * per + i < n; i += blockDim.x) {
            uint4* d = p + blk * per + i;
            if (cs) asm volatile("st.global.cs.v4.b32 [%0], {%1, %2, %3, %4};" ::"l"(d), "r"(1), "r"(2), "r"(3), "r"(4) : "memory");
            else *d = make_uint4(1, 2, 3, 4);
        }
}
// TMA bulk store: smem -> global in 16 KiB pieces
__global__ void st_bulk(uint8_t* p, size_t bytes) {
    extern __shared__ __align__(128) uint8_t sm[];
    const int PIECE = 16384;
    for (int i = threadIdx.x; i < PIECE / 4; i += blockDim.x) reinterpret_cast<uint32_t*>(sm)[i] = i;
    asm volatile("fence.proxy.async.shared::cta;" ::: "memory");
    __syncthreads();
    if (threadIdx.x == 0) {
        uint32_t s = (uint32_t)__cvta_generic_to_shared(sm);
        int k = 0;
        for (size_t off = (size_t)blockIdx.x * PIECE; off + PIECE <= bytes; off += (size_t)gridDim.x * PIECE) {
            asm volatile("cp.async.bulk.global.shared::cta.bulk_group [%0], [%1], %2;" ::"l"(p + off), "r"(s), "r"(PIECE) : "memory");
            asm volatile("cp.async.bulk.commit_group;" ::: "memory");
            if (++k % 8 == 0) asm volatile("cp.async.bulk.wait_group.read 4;" ::: "memory");
        }
        asm volatile("cp.async.bulk.wait_group 0;" ::: "memory");
    }
}
__global__ void rd_plain(const uint4* p, size_t n, uint4* sink) {
    uint4 a = make_uint4(0, 0, 0, 0);
    for (size_t i = (size_t)blockIdx.x * blockDim.x + threadIdx.x; i < n; i += (size_t)gridDim.x * blockDim.x) {
        uint4 v = __ldg(p + i);
        a.x ^= v.x; a.y ^= v.y; a.z ^= v.z; a.w ^= v.w;
    }
    if (a.x == 0x12345678u) *sink = a;
}

template <typename F> float timeit(F f, int reps = 10) {
    cudaEvent_t e0, e1;
    cudaEventCreate(&e0); cudaEventCreate(&e1);
    for (int i = 0; i < 3; ++i) f();
    cudaEventRecord(e0);
    for (int i = 0; i < reps; ++i) f();
    cudaEventRecord(e1);
    cudaEventSynchronize(e1);
    float ms; cudaEventElapsedTime(&ms, e0, e1);
    return ms / reps;
}

int main() {
    const size_t bytes = (size_t)4 << 30;
    uint8_t* buf; cudaMalloc(&buf, bytes);
    uint4* sink; cudaMalloc(&sink, 16);
    const size_t n = bytes / 16;
    const double gb = bytes / 1e9;
    int sms = 148;
    for (int mult : {8, 32}) {
        printf("{\"grid_per_sm\": %d, \"st_plain_gbs\": %.0f", mult, gb / timeit([&] { st_plain<<<sms * mult, 256>>>((uint4*)buf, n); }) * 1e3);
        printf(", \"st_cs_gbs\": %.0f", gb / timeit([&] { st_cs<<<sms * mult, 256>>>((uint4*)buf, n); }) * 1e3);
        printf(", \"st_wt_gbs\": %.0f", gb / timeit([&] { st_wt<<<sms * mult, 256>>>((uint4*)buf, n); }) * 1e3);
        printf(", \"st_blocked_gbs\": %.0f", gb / timeit([&] { st_blocked<<<sms * mult, 256>>>((uint4*)buf, n, 0); }) * 1e3);
        printf(", \"st_blocked_cs_gbs\": %.0f", gb / timeit([&] { st_blocked<<<sms * mult, 256>>>((uint4*)buf, n, 1); }) * 1e3);
        printf(", \"rd_plain_gbs\": %.0f}\n", gb / timeit([&] { rd_plain<<<sms * mult, 256>>>((const uint4*)buf, n, sink); }) * 1e3);
    }
    for (int mult : {1, 2, 4}) {
        cudaFuncSetAttribute(st_bulk, cudaFuncAttributeMaxDynamicSharedMemorySize, 16384);
        printf("{\"grid_per_sm\": %d, \"st_tma_bulk_16k_gbs\": %.0f}\n", mult, gb / timeit([&] { st_bulk<<<sms * mult, 128, 16384>>>(buf, bytes); }) * 1e3);
    }
    cudaMemset(buf, 0, bytes);
    printf("{\"memset_gbs\": %.0f}\n", gb / timeit([&] { cudaMemsetAsync(buf, 0, bytes); }) * 1e3);
    printf("%s\n", cudaGetErrorString(cudaDeviceSynchronize()));
    return 0;
}
