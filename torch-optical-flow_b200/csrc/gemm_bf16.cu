// gemm_bf16.cu -- tcgen05 GEMMs for the backward of the correlation block, and the cast that feeds them.
//
// With dP_l the gradient of pyramid level l (queries x targets, accumulated in fp32 by lookup_bwd.cu), autograd
// through the reference's CorrBlock (methods/raft/model/corr.py:45-54,79-87: matmul, / sqrt(C), avg_pool2d chain) is
//     d fmap1[q, :]          = sum_l  sum_t dP_l[q, t] * pool_l(fmap2)[t, :] / sqrt(C)
//     d pool_l(fmap2)[t, :]  =        sum_q dP_l[q, t] * fmap1[q, :]         / sqrt(C)
// Both are D[M x N] = A[M x K] . B[N x K]^T with N = C <= 256 and a LONG K (the other spatial axis), the opposite of
// the forward builder (short K = C, huge M x N): here both operands stream through a TMA ring, the 128 x N
// accumulator lives in TMEM for the whole K loop, and the epilogue is a small fp32 store.
//   ofb_cast_bf16        fp32 (rows x cols) -> bf16 copy (A of the first product) and bf16 transpose (A of the second),
//                        row pitches padded to 8 elements (TMA strides are multiples of 16 bytes), padding zeroed
//   ofb_gemm_nt_bf16     batched D = alpha * A . B^T (+ D), bf16 K-major operands, fp32 accumulate / output
// Warp roles as in corr_gemm.cu: warp 0 = TMA producer, warp 1 = MMA issuer + TMEM owner, warps 2-5 = epilogue
// (one TMEM lane quarter each); accumulators are double-buffered so the epilogue of one tile overlaps the K loop of
// the next.  Every mbarrier wait is bounded (trap instead of hang).
#include <cuda.h>
#include <cudaTypedefs.h>

#include "common.cuh"

namespace {

constexpr int BM = 128, BK = 64, UK = 16;
constexpr int A_STAGE = BM * BK * 2;                 // 16 KiB
constexpr int B_STAGE_MAX = 256 * BK * 2;            // 32 KiB
constexpr int STAGES = 4;
constexpr int NT = 64 + 4 * 32;                      // TMA warp, MMA warp, 4 epilogue warps
constexpr int OFF_BAR = STAGES * (A_STAGE + B_STAGE_MAX);
constexpr int NUM_BARS = 2 * STAGES + 4;
constexpr int SMEM_ALLOC = OFF_BAR + NUM_BARS * 8 + 16 + 1024;
static_assert(SMEM_ALLOC <= 232448, "shared memory budget");

struct GemmArgs {
    float* D;
    long long ldd, strideD;
    int batch, M, N, K, mtiles, kblocks, items;
    float alpha;
    int accumulate;
};

__device__ __forceinline__ uint32_t smem_u32(const void* p) { return (uint32_t)__cvta_generic_to_shared(p); }
__device__ __forceinline__ void mbar_init(uint32_t bar, uint32_t count) {
    asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(bar), "r"(count));
}
__device__ __forceinline__ void mbar_expect_tx(uint32_t bar, uint32_t bytes) {
    asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(bar), "r"(bytes) : "memory");
}
__device__ __forceinline__ void mbar_arrive(uint32_t bar) {
    asm volatile("mbarrier.arrive.shared::cta.b64 _, [%0];" ::"r"(bar) : "memory");
}
__device__ __forceinline__ bool mbar_try_wait(uint32_t bar, uint32_t parity) {
    uint32_t ok;
    asm volatile(
        "{\n\t.reg .pred p;\n\t"
        "mbarrier.try_wait.parity.shared::cta.b64 p, [%1], %2;\n\t"
        "selp.u32 %0, 1, 0, p;\n\t}"
        : "=r"(ok) : "r"(bar), "r"(parity) : "memory");
    return ok != 0;
}
__device__ __forceinline__ void mbar_wait(uint32_t bar, uint32_t parity) {
    uint32_t spins = 0;
    while (!mbar_try_wait(bar, parity)) {
        if (++spins > (1u << 27)) __trap();          // a protocol bug becomes a trapped launch, never a hung GPU
    }
}
__device__ __forceinline__ void tc_fence_before() { asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory"); }
__device__ __forceinline__ void tc_fence_after() { asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory"); }
__device__ __forceinline__ void tma_load_3d(uint32_t dst, const CUtensorMap* map, uint32_t bar, int c0, int c1, int c2) {
    asm volatile(
        "cp.async.bulk.tensor.3d.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1, {%3, %4, %5}], [%2];"
        ::"r"(dst), "l"(map), "r"(bar), "r"(c0), "r"(c1), "r"(c2) : "memory");
}
__device__ __forceinline__ void umma_bf16(uint32_t d_tmem, uint64_t adesc, uint64_t bdesc, uint32_t idesc, uint32_t accumulate) {
    asm volatile(
        "{\n\t.reg .pred p;\n\tsetp.ne.b32 p, %4, 0;\n\t"
        "tcgen05.mma.cta_group::1.kind::f16 [%0], %1, %2, %3, p;\n\t}"
        ::"r"(d_tmem), "l"(adesc), "l"(bdesc), "r"(idesc), "r"(accumulate) : "memory");
}
__device__ __forceinline__ void umma_commit(uint32_t bar) {
    asm volatile("tcgen05.commit.cta_group::1.mbarrier::arrive::one.shared::cluster.b64 [%0];" ::"r"(bar) : "memory");
}
__device__ __forceinline__ void tmem_ld32(uint32_t taddr, uint32_t (&v)[32]) {
    asm volatile(
        "tcgen05.ld.sync.aligned.32x32b.x32.b32 "
        "{%0, %1, %2, %3, %4, %5, %6, %7, %8, %9, %10, %11, %12, %13, %14, %15, "
        "%16, %17, %18, %19, %20, %21, %22, %23, %24, %25, %26, %27, %28, %29, %30, %31}, [%32];"
        : "=r"(v[0]), "=r"(v[1]), "=r"(v[2]), "=r"(v[3]), "=r"(v[4]), "=r"(v[5]), "=r"(v[6]), "=r"(v[7]),
          "=r"(v[8]), "=r"(v[9]), "=r"(v[10]), "=r"(v[11]), "=r"(v[12]), "=r"(v[13]), "=r"(v[14]), "=r"(v[15]),
          "=r"(v[16]), "=r"(v[17]), "=r"(v[18]), "=r"(v[19]), "=r"(v[20]), "=r"(v[21]), "=r"(v[22]), "=r"(v[23]),
          "=r"(v[24]), "=r"(v[25]), "=r"(v[26]), "=r"(v[27]), "=r"(v[28]), "=r"(v[29]), "=r"(v[30]), "=r"(v[31])
        : "r"(taddr) : "memory");
}
// K-major, 128-byte swizzled operand tile (rows of 128 bytes, 8-row groups 1024 bytes apart): as corr_gemm.cu
__device__ __forceinline__ uint64_t make_smem_desc(uint32_t saddr) {
    uint64_t d = 0;
    d |= (uint64_t)((saddr & 0x3FFFFu) >> 4);
    d |= (uint64_t)1 << 16;
    d |= (uint64_t)(1024 >> 4) << 32;
    d |= (uint64_t)1 << 46;
    d |= (uint64_t)2 << 61;
    return d;
}
__device__ __forceinline__ uint32_t make_idesc(int M, int N) {
    return (1u << 4) | (1u << 7) | (1u << 10) | ((uint32_t)(N >> 3) << 17) | ((uint32_t)(M >> 4) << 24);
}

__global__ void __launch_bounds__(NT, 1) gemm_nt_kernel(const __grid_constant__ CUtensorMap map_a,
                                                        const __grid_constant__ CUtensorMap map_b, const GemmArgs G) {
    extern __shared__ uint8_t smem_raw[];
    const uint32_t raw = smem_u32(smem_raw);
    const uint32_t sbase = (raw + 1023u) & ~1023u;       // the 128-byte swizzle is a function of the absolute address
    uint8_t* sgen = smem_raw + (sbase - raw);
    const uint32_t bar_full = sbase + OFF_BAR, bar_empty = bar_full + 8 * STAGES;
    const uint32_t bar_tfull = bar_empty + 8 * STAGES, bar_tempty = bar_tfull + 16;
    const uint32_t tmem_slot = sbase + OFF_BAR + NUM_BARS * 8;
    volatile uint32_t* tmem_slot_gen = reinterpret_cast<volatile uint32_t*>(sgen + OFF_BAR + NUM_BARS * 8);
    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    const uint32_t b_stage = (uint32_t)G.N * BK * 2;
    const uint32_t off_b = STAGES * A_STAGE;

    if (warp == 0 && lane == 0) {
        asm volatile("prefetch.tensormap [%0];" ::"l"(&map_a) : "memory");
        asm volatile("prefetch.tensormap [%0];" ::"l"(&map_b) : "memory");
        for (int s = 0; s < STAGES; ++s) { mbar_init(bar_full + 8 * s, 1); mbar_init(bar_empty + 8 * s, 1); }
        for (int s = 0; s < 2; ++s) { mbar_init(bar_tfull + 8 * s, 1); mbar_init(bar_tempty + 8 * s, 4); }
        asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
    }
    if (warp == 1) {
        asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(tmem_slot), "r"(512) : "memory");
        asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;" ::: "memory");
    }
    tc_fence_before();
    __syncthreads();
    tc_fence_after();
    const uint32_t tmem_base = *tmem_slot_gen;

    if (warp == 0) {
        if (lane == 0) {
            uint32_t stage = 0, phase = 0;
            for (int item = blockIdx.x; item < G.items; item += gridDim.x) {
                const int b = item / G.mtiles, m0 = (item - b * G.mtiles) * BM;
                for (int kb = 0; kb < G.kblocks; ++kb) {
                    mbar_wait(bar_empty + 8 * stage, phase ^ 1);
                    mbar_expect_tx(bar_full + 8 * stage, (uint32_t)A_STAGE + b_stage);
                    tma_load_3d(sbase + stage * A_STAGE, &map_a, bar_full + 8 * stage, kb * BK, m0, b);
                    tma_load_3d(sbase + off_b + stage * B_STAGE_MAX, &map_b, bar_full + 8 * stage, kb * BK, 0, b);
                    if (++stage == STAGES) { stage = 0; phase ^= 1; }
                }
            }
        }
        __syncwarp();
    } else if (warp == 1) {
        if (lane == 0) {
            const uint32_t idesc = make_idesc(BM, G.N);
            uint32_t stage = 0, phase = 0, acc = 0, tphase = 0;
            for (int item = blockIdx.x; item < G.items; item += gridDim.x) {
                mbar_wait(bar_tempty + 8 * acc, tphase ^ 1);
                tc_fence_after();
                const uint32_t d_tmem = tmem_base + acc * 256;
                for (int kb = 0; kb < G.kblocks; ++kb) {
                    mbar_wait(bar_full + 8 * stage, phase);
                    tc_fence_after();
                    const uint64_t adesc = make_smem_desc(sbase + stage * A_STAGE);
                    const uint64_t bdesc = make_smem_desc(sbase + off_b + stage * B_STAGE_MAX);
#pragma unroll
                    for (int k = 0; k < BK / UK; ++k)
                        umma_bf16(d_tmem, adesc + (uint64_t)(2 * k), bdesc + (uint64_t)(2 * k), idesc, (uint32_t)((kb | k) != 0));
                    umma_commit(bar_empty + 8 * stage);
                    if (++stage == STAGES) { stage = 0; phase ^= 1; }
                }
                umma_commit(bar_tfull + 8 * acc);
                if (++acc == 2) { acc = 0; tphase ^= 1; }
            }
        }
        __syncwarp();
    } else {
        const int q4 = warp & 3;                          // TMEM lane quarter this warp may read (warp_id % 4)
        uint32_t acc = 0, tphase = 0;
        for (int item = blockIdx.x; item < G.items; item += gridDim.x) {
            const int b = item / G.mtiles, m0 = (item - b * G.mtiles) * BM;
            const int row = m0 + q4 * 32 + lane;
            float* drow = G.D + (long long)b * G.strideD + (long long)row * G.ldd;
            mbar_wait(bar_tfull + 8 * acc, tphase);
            tc_fence_after();
            const uint32_t taddr = tmem_base + ((uint32_t)(q4 * 32) << 16) + acc * 256;
            for (int c0 = 0; c0 < G.N; c0 += 32) {
                uint32_t v[32];
                tmem_ld32(taddr + c0, v);
                asm volatile("tcgen05.wait::ld.sync.aligned;" ::: "memory");
                if (row < G.M) {
#pragma unroll
                    for (int c = 0; c < 32; c += 4) {
                        float4 o = make_float4(__uint_as_float(v[c]) * G.alpha, __uint_as_float(v[c + 1]) * G.alpha,
                                               __uint_as_float(v[c + 2]) * G.alpha, __uint_as_float(v[c + 3]) * G.alpha);
                        float4* dst = reinterpret_cast<float4*>(drow + c0 + c);
                        if (G.accumulate) { const float4 p = *dst; o.x += p.x; o.y += p.y; o.z += p.z; o.w += p.w; }
                        *dst = o;
                    }
                }
            }
            tc_fence_before();
            __syncwarp();
            if (lane == 0) mbar_arrive(bar_tempty + 8 * acc);
            if (++acc == 2) { acc = 0; tphase ^= 1; }
        }
    }
    tc_fence_before();
    __syncthreads();
    if (warp == 1) asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" ::"r"(tmem_base), "r"(512) : "memory");
}

// fp32 (rows x cols, tight) -> bf16 copy (pitch pc) and / or bf16 transpose (pitch pr); 32 x 32 tiles through smem
__global__ void __launch_bounds__(256) cast_bf16_kernel(const float* __restrict__ src, __nv_bfloat16* __restrict__ dst,
                                                        __nv_bfloat16* __restrict__ dst_t, int rows, int cols, long long pc,
                                                        long long pr) {
    __shared__ float tile[32][33];
    const int b = blockIdx.z, r0 = blockIdx.y * 32, c0 = blockIdx.x * 32;
    const int tx = threadIdx.x & 31, ty = threadIdx.x >> 5;   // ty 0..7
    const float* s = src + (size_t)b * rows * cols;
#pragma unroll
    for (int k = 0; k < 4; ++k) {
        const int r = r0 + ty + 8 * k, c = c0 + tx;
        const float v = (r < rows && c < cols) ? __ldg(s + (size_t)r * cols + c) : 0.0f;
        tile[ty + 8 * k][tx] = v;
        if (dst && r < rows && c < pc) dst[((size_t)b * rows + r) * pc + c] = __float2bfloat16_rn(v);
    }
    if (!dst_t) return;
    __syncthreads();
#pragma unroll
    for (int k = 0; k < 4; ++k) {
        const int c = c0 + ty + 8 * k, r = r0 + tx;           // transposed: row index of dst_t = column of src
        if (c < cols && r < pr) dst_t[((size_t)b * cols + c) * pr + r] = __float2bfloat16_rn(tile[tx][ty + 8 * k]);
    }
}

PFN_cuTensorMapEncodeTiled_v12000 get_encode() {
    static PFN_cuTensorMapEncodeTiled_v12000 fn = nullptr;
    static bool tried = false;
    if (!tried) {
        tried = true;
        void* p = nullptr;
        cudaDriverEntryPointQueryResult qres;
        if (cudaGetDriverEntryPoint("cuTensorMapEncodeTiled", &p, cudaEnableDefault, &qres) == cudaSuccess &&
            qres == cudaDriverEntryPointSuccess)
            fn = reinterpret_cast<PFN_cuTensorMapEncodeTiled_v12000>(p);
    }
    return fn;
}

bool encode_operand(CUtensorMap* m, const void* base, int K, int rows, int batch, long long ld, long long stride, int box_rows) {
    PFN_cuTensorMapEncodeTiled_v12000 enc = get_encode();
    if (!enc) return false;
    cuuint64_t gdim[3] = {(cuuint64_t)K, (cuuint64_t)rows, (cuuint64_t)batch};
    cuuint64_t gstr[2] = {(cuuint64_t)ld * 2, (cuuint64_t)stride * 2};
    cuuint32_t box[3] = {BK, (cuuint32_t)box_rows, 1};
    cuuint32_t estr[3] = {1, 1, 1};
    return enc(m, CU_TENSOR_MAP_DATA_TYPE_BFLOAT16, 3, const_cast<void*>(base), gdim, gstr, box, estr,
               CU_TENSOR_MAP_INTERLEAVE_NONE, CU_TENSOR_MAP_SWIZZLE_128B, CU_TENSOR_MAP_L2_PROMOTION_L2_256B,
               CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE) == CUDA_SUCCESS;
}

}  // namespace

OFB_API int ofb_gemm_nt_bf16(const void* A, const void* B, float* D, int batch, int M, int N, int K, long long lda,
                             long long ldb, long long ldd, long long strideA, long long strideB, long long strideD,
                             float alpha, int accumulate, void* stream) {
    if (batch == 0 || M == 0) return OFB_OK;
    if (!A || !B || !D || batch < 0 || M < 0 || N <= 0 || K <= 0) return OFB_EINVAL;
    if (N % 32 != 0 || N > 256 || lda % 8 || ldb % 8 || strideA % 8 || strideB % 8 || ldd % 4 || strideD % 4 || lda < K || ldb < K ||
        ldd < N)
        return OFB_EUNSUPPORTED;
    if ((reinterpret_cast<uintptr_t>(A) | reinterpret_cast<uintptr_t>(B) | reinterpret_cast<uintptr_t>(D)) & 15) return OFB_EALIGN;
    CUtensorMap ma, mb;
    // batch == 1: the batch stride is unused but must still be a legal (non-zero, 16-byte multiple) stride
    const long long sa = batch > 1 ? strideA : (long long)M * lda, sb = batch > 1 ? strideB : (long long)N * ldb;
    if (!encode_operand(&ma, A, K, M, batch, lda, sa, BM) || !encode_operand(&mb, B, K, N, batch, ldb, sb, N)) return OFB_EDRIVER;
    GemmArgs G;
    G.D = D; G.ldd = ldd; G.strideD = strideD;
    G.batch = batch; G.M = M; G.N = N; G.K = K;
    G.mtiles = (M + BM - 1) / BM;
    G.kblocks = (K + BK - 1) / BK;
    const long long items = (long long)batch * G.mtiles;
    if (items > 0x7fffffffLL) return OFB_EUNSUPPORTED;
    G.items = (int)items;
    G.alpha = alpha; G.accumulate = accumulate;
    static bool configured[OFB_MAX_DEVICES] = {false};
    const int dev = ofb_device();
    if (!configured[dev]) {
        OFB_CUDA(cudaFuncSetAttribute(gemm_nt_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, SMEM_ALLOC));
        configured[dev] = true;
    }
    const int grid = (int)(items < ofb_num_sms() ? items : ofb_num_sms());
    gemm_nt_kernel<<<grid, NT, SMEM_ALLOC, (cudaStream_t)stream>>>(ma, mb, G);
    OFB_LAUNCH_CHECK();
    return OFB_OK;
}

OFB_API int ofb_cast_bf16(const float* src, void* dst_or_null, void* dst_t_or_null, int batch, int rows, int cols,
                          long long pitch_dst, long long pitch_dst_t, void* stream) {
    if (batch == 0 || rows == 0 || cols == 0) return OFB_OK;
    if (!src || (!dst_or_null && !dst_t_or_null) || batch < 0 || rows < 0 || cols < 0) return OFB_EINVAL;
    if ((dst_or_null && pitch_dst < cols) || (dst_t_or_null && pitch_dst_t < rows)) return OFB_EINVAL;
    // the padding is written from the same tiles: it may extend at most to the next multiple of 32
    const long long cmax = ((long long)cols + 31) / 32 * 32, rmax = ((long long)rows + 31) / 32 * 32;
    if ((dst_or_null && pitch_dst > cmax) || (dst_t_or_null && pitch_dst_t > rmax) || batch > 65535) return OFB_EUNSUPPORTED;
    const dim3 grid((cols + 31) / 32, (rows + 31) / 32, batch);
    if (grid.y > 65535) return OFB_EUNSUPPORTED;
    cast_bf16_kernel<<<grid, 256, 0, (cudaStream_t)stream>>>(src, static_cast<__nv_bfloat16*>(dst_or_null),
                                                            static_cast<__nv_bfloat16*>(dst_t_or_null), rows, cols, pitch_dst,
                                                            pitch_dst_t);
    OFB_LAUNCH_CHECK();
    return OFB_OK;
}
