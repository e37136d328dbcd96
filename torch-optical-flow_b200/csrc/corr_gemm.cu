// corr_gemm.cu -- K2: all-pairs correlation pyramid on the 5th-generation tensor cores.
//
// Replaces CorrBlock.corr + CorrBlock.__init__ (reference methods/raft/model/corr.py:38-54,79-87):
//     corr[b,p,q] = sum_c f1[b,c,p] * f2[b,c,q] / sqrt(C)          (torch.matmul + full-volume divide)
//     level l     = 2x2 average pooling of level l-1 over the target image, complete blocks only
// The reference makes one cuBLAS launch plus 4 further full passes over a multi-GB volume.  Here
// ONE persistent kernel produces every level:
//
//   * operands are K-major bf16 (ofb_corr_prep_bf16), fetched by TMA with 128-byte swizzle;
//   * a CTA owns 128 queries (rows of the volume).  Their 128 x C slice of fmap1 stays RESIDENT in
//     shared memory while the CTA walks target tiles; only fmap2 streams (and hits L2: every CTA
//     walks the same tiles at the same time);
//   * a target tile is a SPATIAL PATCH of 8 rows x 32 columns of the h x w target image, fetched with
//     a 4-D tensor map (C, x, y, b): 256 accumulator columns, column n = 32*row + col.  TMA zero-fills
//     out-of-image elements, so edge tiles need no masking;
//   * tcgen05.mma (kind::f16, bf16 x bf16 -> fp32) accumulates 128 x 256 in TMEM; two accumulator
//     stages (2 x 256 of the 512 columns) let the epilogue of tile n overlap the MMAs of tile n+1;
//   * with cta_group = 2 a CTA PAIR shares each target tile (each CTA loads half of it, the MMA is
//     M = 256 across the pair), halving the L2 -> SM operand traffic per flop;
//   * epilogue: each thread owns one query row; because a tile is a patch, the 2x2, 4x4 and 8x8
//     means over the target image are sums of registers of ONE thread -- no shuffles, no re-read of
//     level 0.  1/sqrt(C) is applied to the fp32 accumulator, levels 0..2 go through a swizzled
//     shared-memory stage and leave as TMA tensor stores (which also clip to floor(h/2^l) x
//     floor(w/2^l): the reference's "complete blocks only" rule for free), level 3 (4 values per
//     thread and tile) is stored directly.
//
// Roofline: 2*B*N^2*C flops against the bf16 tensor peak, and
// 2 bytes * 1.33 * B*N^2 of pyramid writes against HBM -- at C = 256 the two are within 15 %.
#include <cuda.h>
#include <cudaTypedefs.h>

#include <cstdlib>

#include "common.cuh"

namespace {

constexpr int BLOCK_M = 128;
constexpr int PATCH_W = 32, PATCH_H = 8;
constexpr int TILE_N = PATCH_W * PATCH_H;   // 256 accumulator columns
constexpr int BLOCK_K = 64;                 // one 128-byte swizzle span of bf16
constexpr int UMMA_K = 16;
constexpr int MAX_KB = 4;                   // C <= 256
constexpr int A_KB_BYTES = BLOCK_M * BLOCK_K * 2;          // 16 KiB
constexpr int B_TILE_KB_BYTES = TILE_N * BLOCK_K * 2;      // 32 KiB per k-block for a whole tile
constexpr int B_RING_BYTES = 64 * 1024;                    // 2 stages (cta_group 1) / 4 stages (cta_group 2)
constexpr int ST_L0_BYTES = PATCH_H * BLOCK_M * 64;        // [8][128][64 B]   64 KiB
constexpr int ST_L1_BYTES = (PATCH_H / 2) * BLOCK_M * 32;  // [4][128][32 B]   16 KiB
constexpr int ST_L2_BYTES = (PATCH_H / 4) * BLOCK_M * 16;  // [2][128][16 B]    4 KiB
constexpr int NUM_THREADS = 192;            // warp 0: TMA, warp 1: MMA + TMEM, warps 2..5: epilogue
constexpr int NUM_EPI_WARPS = 4;
constexpr int TMEM_COLS = 512;

constexpr int OFF_A = 0;
constexpr int OFF_B = OFF_A + MAX_KB * A_KB_BYTES;
constexpr int OFF_ST0 = OFF_B + B_RING_BYTES;
constexpr int OFF_ST1 = OFF_ST0 + ST_L0_BYTES;
constexpr int OFF_ST2 = OFF_ST1 + ST_L1_BYTES;
constexpr int OFF_BAR = OFF_ST2 + ST_L2_BYTES;
constexpr int NUM_BARS = 2 + 4 + 4 + 2 + 2;   // a_full, a_empty, b_full[4], b_empty[4], t_full[2], t_empty[2]
constexpr int SMEM_BYTES = OFF_BAR + NUM_BARS * 8 + 16;
constexpr int SMEM_ALLOC = SMEM_BYTES + 1024;  // manual 1024-byte alignment of the dynamic segment

struct GemmParams {
    int B, C, h, w, N;           // N = h*w
    int levels;
    int kb;                      // C / 64
    int ntx, nty, ntiles;        // target tiles
    int tiles_per_item, n_chunks, mblk, n_items;
    float scale;
    __nv_bfloat16* l3_base;      // level 3 is stored directly
    long long l3_qs;
    int l3_pitch, l3_h, l3_w;
    int dbg;                     // PROF builds only: bit mask disabling parts of the epilogue (tools/k2_profile.py)
    unsigned long long* prof;    // optional per-CTA wait-cycle counters (ofb_corr_pyramid_bf16_profile)
};

// ------------------------------------------------------------------------------------ PTX helpers
__device__ __forceinline__ uint32_t smem_u32(const void* p) { return (uint32_t)__cvta_generic_to_shared(p); }

__device__ __forceinline__ void mbar_init(uint32_t bar, uint32_t count) {
    asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(bar), "r"(count));
}
__device__ __forceinline__ void mbar_expect_tx(uint32_t bar, uint32_t bytes) {
    asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(bar), "r"(bytes) : "memory");
}
__device__ __forceinline__ void mbar_arrive_local(uint32_t bar) {
    asm volatile("mbarrier.arrive.shared::cta.b64 _, [%0];" ::"r"(bar) : "memory");
}
// arrive on the barrier at the same offset in CTA `rank` of the cluster
__device__ __forceinline__ void mbar_arrive_cluster(uint32_t bar, uint32_t rank) {
    asm volatile(
        "{\n\t.reg .b32 ra;\n\t"
        "mapa.shared::cluster.u32 ra, %0, %1;\n\t"
        "mbarrier.arrive.release.cluster.shared::cluster.b64 _, [ra];\n\t}"
        ::"r"(bar), "r"(rank) : "memory");
}
__device__ __forceinline__ uint32_t map_to_rank(uint32_t addr, uint32_t rank) {
    uint32_t r;
    asm volatile("mapa.shared::cluster.u32 %0, %1, %2;" : "=r"(r) : "r"(addr), "r"(rank));
    return r;
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
// Bounded wait: a protocol bug becomes a trapped launch, never a hung GPU.
__device__ __forceinline__ void mbar_wait(uint32_t bar, uint32_t parity) {
    uint32_t spins = 0;
    while (!mbar_try_wait(bar, parity)) {
        if (++spins > (1u << 26)) __trap();
    }
}
template <bool PROF>
__device__ __forceinline__ void mbar_wait_p(uint32_t bar, uint32_t parity, unsigned long long& acc) {
    if (PROF) {
        const long long t0 = clock64();
        mbar_wait(bar, parity);
        acc += (unsigned long long)(clock64() - t0);
    } else {
        mbar_wait(bar, parity);
    }
}
__device__ __forceinline__ void fence_barrier_init() { asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory"); }
__device__ __forceinline__ void fence_proxy_async_smem() { asm volatile("fence.proxy.async.shared::cta;" ::: "memory"); }
__device__ __forceinline__ void tc_fence_before() { asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory"); }
__device__ __forceinline__ void tc_fence_after() { asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory"); }
__device__ __forceinline__ void cluster_sync_all() {
    asm volatile("barrier.cluster.arrive.release.aligned;\n\tbarrier.cluster.wait.acquire.aligned;" ::: "memory");
}
__device__ __forceinline__ uint32_t cluster_ctarank() {
    uint32_t r;
    asm volatile("mov.u32 %0, %%cluster_ctarank;" : "=r"(r));
    return r;
}
__device__ __forceinline__ void epi_bar_sync() { asm volatile("bar.sync 1, %0;" ::"n"(NUM_EPI_WARPS * 32) : "memory"); }

// TMA loads: the completion bytes are credited to `bar` (a shared::cluster address: for a CTA pair the
// leader's barrier collects both CTAs' loads).
template <int CG>
__device__ __forceinline__ void tma_load_3d(uint32_t dst, const CUtensorMap* map, uint32_t bar, int c0, int c1, int c2) {
    if (CG == 1)
        asm volatile(
            "cp.async.bulk.tensor.3d.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1, {%3, %4, %5}], [%2];"
            ::"r"(dst), "l"(map), "r"(bar), "r"(c0), "r"(c1), "r"(c2) : "memory");
    else
        asm volatile(
            "cp.async.bulk.tensor.3d.cta_group::2.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1, {%3, %4, %5}], [%2];"
            ::"r"(dst), "l"(map), "r"(bar), "r"(c0), "r"(c1), "r"(c2) : "memory");
}
template <int CG>
__device__ __forceinline__ void tma_load_4d(uint32_t dst, const CUtensorMap* map, uint32_t bar, int c0, int c1, int c2,
                                            int c3) {
    if (CG == 1)
        asm volatile(
            "cp.async.bulk.tensor.4d.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1, {%3, %4, %5, %6}], [%2];"
            ::"r"(dst), "l"(map), "r"(bar), "r"(c0), "r"(c1), "r"(c2), "r"(c3) : "memory");
    else
        asm volatile(
            "cp.async.bulk.tensor.4d.cta_group::2.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1, {%3, %4, %5, %6}], [%2];"
            ::"r"(dst), "l"(map), "r"(bar), "r"(c0), "r"(c1), "r"(c2), "r"(c3) : "memory");
}
__device__ __forceinline__ void tma_store_4d(const CUtensorMap* map, uint32_t src, int c0, int c1, int c2, int c3) {
    asm volatile("cp.async.bulk.tensor.4d.global.shared::cta.bulk_group [%0, {%2, %3, %4, %5}], [%1];"
                 ::"l"(map), "r"(src), "r"(c0), "r"(c1), "r"(c2), "r"(c3) : "memory");
}
__device__ __forceinline__ void tma_store_commit() { asm volatile("cp.async.bulk.commit_group;" ::: "memory"); }
__device__ __forceinline__ void tma_store_wait_read() { asm volatile("cp.async.bulk.wait_group.read 0;" ::: "memory"); }
__device__ __forceinline__ void tma_store_wait_all() { asm volatile("cp.async.bulk.wait_group 0;" ::: "memory"); }
__device__ __forceinline__ void prefetch_tmap(const CUtensorMap* m) {
    asm volatile("prefetch.tensormap [%0];" ::"l"(m) : "memory");
}

// tcgen05
template <int CG>
__device__ __forceinline__ void tmem_alloc(uint32_t dst_smem, uint32_t cols) {
    if (CG == 1) asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(dst_smem), "r"(cols) : "memory");
    else         asm volatile("tcgen05.alloc.cta_group::2.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(dst_smem), "r"(cols) : "memory");
}
template <int CG>
__device__ __forceinline__ void tmem_relinquish() {
    if (CG == 1) asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;" ::: "memory");
    else         asm volatile("tcgen05.relinquish_alloc_permit.cta_group::2.sync.aligned;" ::: "memory");
}
template <int CG>
__device__ __forceinline__ void tmem_dealloc(uint32_t addr, uint32_t cols) {
    if (CG == 1) asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" ::"r"(addr), "r"(cols) : "memory");
    else         asm volatile("tcgen05.dealloc.cta_group::2.sync.aligned.b32 %0, %1;" ::"r"(addr), "r"(cols) : "memory");
}
template <int CG>
__device__ __forceinline__ void umma_bf16(uint32_t d_tmem, uint64_t adesc, uint64_t bdesc, uint32_t idesc, uint32_t accumulate) {
    if (CG == 1)
        asm volatile(
            "{\n\t.reg .pred p;\n\tsetp.ne.b32 p, %4, 0;\n\t"
            "tcgen05.mma.cta_group::1.kind::f16 [%0], %1, %2, %3, p;\n\t}"
            ::"r"(d_tmem), "l"(adesc), "l"(bdesc), "r"(idesc), "r"(accumulate) : "memory");
    else
        asm volatile(
            "{\n\t.reg .pred p;\n\tsetp.ne.b32 p, %4, 0;\n\t"
            "tcgen05.mma.cta_group::2.kind::f16 [%0], %1, %2, %3, p;\n\t}"
            ::"r"(d_tmem), "l"(adesc), "l"(bdesc), "r"(idesc), "r"(accumulate) : "memory");
}
// all previously issued MMAs of this thread arrive on `bar` when they retire (both CTAs of a pair)
template <int CG>
__device__ __forceinline__ void umma_commit(uint32_t bar) {
    if (CG == 1)
        asm volatile("tcgen05.commit.cta_group::1.mbarrier::arrive::one.shared::cluster.b64 [%0];" ::"r"(bar) : "memory");
    else
        asm volatile(
            "tcgen05.commit.cta_group::2.mbarrier::arrive::one.shared::cluster.multicast::cluster.b64 [%0], %1;"
            ::"r"(bar), "h"((uint16_t)3) : "memory");
}
// 32 lanes x 64 consecutive fp32 columns
__device__ __forceinline__ void tmem_ld64(uint32_t taddr, uint32_t (&v)[64]) {
    asm volatile(
        "tcgen05.ld.sync.aligned.32x32b.x64.b32 "
        "{%0, %1, %2, %3, %4, %5, %6, %7, %8, %9, %10, %11, %12, %13, %14, %15, "
        "%16, %17, %18, %19, %20, %21, %22, %23, %24, %25, %26, %27, %28, %29, %30, %31, "
        "%32, %33, %34, %35, %36, %37, %38, %39, %40, %41, %42, %43, %44, %45, %46, %47, "
        "%48, %49, %50, %51, %52, %53, %54, %55, %56, %57, %58, %59, %60, %61, %62, %63}, [%64];"
        : "=r"(v[0]), "=r"(v[1]), "=r"(v[2]), "=r"(v[3]), "=r"(v[4]), "=r"(v[5]), "=r"(v[6]), "=r"(v[7]),
          "=r"(v[8]), "=r"(v[9]), "=r"(v[10]), "=r"(v[11]), "=r"(v[12]), "=r"(v[13]), "=r"(v[14]), "=r"(v[15]),
          "=r"(v[16]), "=r"(v[17]), "=r"(v[18]), "=r"(v[19]), "=r"(v[20]), "=r"(v[21]), "=r"(v[22]), "=r"(v[23]),
          "=r"(v[24]), "=r"(v[25]), "=r"(v[26]), "=r"(v[27]), "=r"(v[28]), "=r"(v[29]), "=r"(v[30]), "=r"(v[31]),
          "=r"(v[32]), "=r"(v[33]), "=r"(v[34]), "=r"(v[35]), "=r"(v[36]), "=r"(v[37]), "=r"(v[38]), "=r"(v[39]),
          "=r"(v[40]), "=r"(v[41]), "=r"(v[42]), "=r"(v[43]), "=r"(v[44]), "=r"(v[45]), "=r"(v[46]), "=r"(v[47]),
          "=r"(v[48]), "=r"(v[49]), "=r"(v[50]), "=r"(v[51]), "=r"(v[52]), "=r"(v[53]), "=r"(v[54]), "=r"(v[55]),
          "=r"(v[56]), "=r"(v[57]), "=r"(v[58]), "=r"(v[59]), "=r"(v[60]), "=r"(v[61]), "=r"(v[62]), "=r"(v[63])
        : "r"(taddr) : "memory");
}
__device__ __forceinline__ void tmem_ld_wait() { asm volatile("tcgen05.wait::ld.sync.aligned;" ::: "memory"); }

// K-major, 128-byte swizzled operand tile: rows of 128 bytes, 8-row groups 1024 bytes apart
// (cute::UMMA::SmemDescriptor: start >> 4 | LBO [16,30) | SBO [32,46) | version=1 [46,48) | layout [61,64))
__device__ __forceinline__ uint64_t make_smem_desc(uint32_t saddr) {
    uint64_t d = 0;
    d |= (uint64_t)((saddr & 0x3FFFFu) >> 4);
    d |= (uint64_t)1 << 16;                 // leading byte offset: unused for swizzled K-major, canonical value 1
    d |= (uint64_t)(1024 >> 4) << 32;       // stride byte offset: 8 rows x 128 B
    d |= (uint64_t)1 << 46;                 // descriptor version (Blackwell)
    d |= (uint64_t)2 << 61;                 // SWIZZLE_128B
    return d;
}
// kind::f16 instruction descriptor: D fp32, A = B = bf16, both K-major
__host__ __device__ constexpr uint32_t make_idesc(int M, int N) {
    return (1u << 4) | (1u << 7) | (1u << 10) | ((uint32_t)(N >> 3) << 17) | ((uint32_t)(M >> 4) << 24);
}

__device__ __forceinline__ uint32_t pack_bf16(float lo, float hi) {
    __nv_bfloat162 v = __floats2bfloat162_rn(lo, hi);
    return *reinterpret_cast<uint32_t*>(&v);
}
__device__ __forceinline__ void st_shared_v4(uint32_t addr, uint32_t a, uint32_t b, uint32_t c, uint32_t d) {
    asm volatile("st.shared.v4.b32 [%0], {%1, %2, %3, %4};" ::"r"(addr), "r"(a), "r"(b), "r"(c), "r"(d) : "memory");
}

struct ItemCoord {
    int b, chunk, m;
};
__device__ __forceinline__ ItemCoord decode_item(const GemmParams& P, int item) {
    ItemCoord ic;
    ic.m = item % P.mblk;
    const int r = item / P.mblk;
    ic.chunk = r % P.n_chunks;
    ic.b = r / P.n_chunks;
    return ic;
}

// ------------------------------------------------------------------------------------ the kernel
template <int CG, bool PROF>
__global__ void __launch_bounds__(NUM_THREADS, 1)
corr_pyramid_kernel(const __grid_constant__ CUtensorMap map_a, const __grid_constant__ CUtensorMap map_b,
                    const __grid_constant__ CUtensorMap map_l0, const __grid_constant__ CUtensorMap map_l1,
                    const __grid_constant__ CUtensorMap map_l2, const GemmParams P) {
    extern __shared__ uint8_t smem_raw[];
    // the 128-byte swizzle is a function of the absolute shared address: align the segment to 1024
    const uint32_t raw = smem_u32(smem_raw);
    const uint32_t sbase = (raw + 1023u) & ~1023u;
    uint8_t* sgen = smem_raw + (sbase - raw);
    constexpr int B_STAGE_BYTES = B_TILE_KB_BYTES / CG;
    constexpr int B_STAGES = B_RING_BYTES / B_STAGE_BYTES;

    const uint32_t bar0 = sbase + OFF_BAR;
    const uint32_t bar_a_full = bar0, bar_a_empty = bar0 + 8;
    const uint32_t bar_b_full = bar0 + 16, bar_b_empty = bar0 + 16 + 32;
    const uint32_t bar_t_full = bar0 + 80, bar_t_empty = bar0 + 96;
    const uint32_t tmem_slot = bar0 + NUM_BARS * 8;
    volatile uint32_t* tmem_slot_gen = reinterpret_cast<volatile uint32_t*>(sgen + OFF_BAR + NUM_BARS * 8);

    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    const uint32_t rank = (CG == 2) ? cluster_ctarank() : 0u;
    const bool leader = rank == 0;
    const int worker = blockIdx.x / CG, n_workers = gridDim.x / CG;

    if (warp == 0 && lane == 0) {
        prefetch_tmap(&map_a); prefetch_tmap(&map_b);
        prefetch_tmap(&map_l0); prefetch_tmap(&map_l1); prefetch_tmap(&map_l2);
        mbar_init(bar_a_full, 1);
        mbar_init(bar_a_empty, 1);
        for (int s = 0; s < 4; ++s) { mbar_init(bar_b_full + 8 * s, 1); mbar_init(bar_b_empty + 8 * s, 1); }
        for (int s = 0; s < 2; ++s) { mbar_init(bar_t_full + 8 * s, 1); mbar_init(bar_t_empty + 8 * s, NUM_EPI_WARPS * CG); }
        fence_barrier_init();
    }
    if (warp == 1) {
        tmem_alloc<CG>(tmem_slot, TMEM_COLS);
        tmem_relinquish<CG>();
    }
    tc_fence_before();
    if (CG == 2) cluster_sync_all(); else __syncthreads();
    tc_fence_after();
    const uint32_t tmem_base = *tmem_slot_gen;

    const int m_rows = BLOCK_M * CG;   // query rows per work item
    const long long t_start = PROF ? clock64() : 0;
    const int dbg = PROF ? P.dbg : 0;
    unsigned long long pw0 = 0, pw1 = 0, pw2 = 0, ptiles = 0;   // per-role wait cycles (PROF only)

    if (warp == 0) {
        // ============================== TMA producer (one lane) ==============================
        if (lane == 0) {
            const uint32_t full_a = (CG == 2) ? map_to_rank(bar_a_full, 0) : bar_a_full;
            uint32_t stage = 0, bphase = 0, aphase = 0;
            for (int item = worker; item < P.n_items; item += n_workers) {
                const ItemCoord ic = decode_item(P, item);
                mbar_wait_p<PROF>(bar_a_empty, aphase ^ 1, pw1);
                aphase ^= 1;
                if (leader) mbar_expect_tx(bar_a_full, (uint32_t)(P.kb * A_KB_BYTES * CG));
                const int row0 = ic.m * m_rows + (int)rank * BLOCK_M;
                for (int kb = 0; kb < P.kb; ++kb)
                    tma_load_3d<CG>(sbase + OFF_A + kb * A_KB_BYTES, &map_a, full_a, kb * BLOCK_K, row0, ic.b);
                const int t0 = ic.chunk * P.tiles_per_item;
                const int t1 = min(t0 + P.tiles_per_item, P.ntiles);
                for (int t = t0; t < t1; ++t) {
                    const int ty = t / P.ntx, tx = t - ty * P.ntx;
                    const int x0 = tx * PATCH_W, y0 = ty * PATCH_H + (int)rank * (PATCH_H / CG);
                    for (int kb = 0; kb < P.kb; ++kb) {
                        mbar_wait_p<PROF>(bar_b_empty + 8 * stage, bphase ^ 1, pw0);
                        const uint32_t full_b = (CG == 2) ? map_to_rank(bar_b_full + 8 * stage, 0) : bar_b_full + 8 * stage;
                        if (leader) mbar_expect_tx(bar_b_full + 8 * stage, (uint32_t)B_TILE_KB_BYTES);
                        tma_load_4d<CG>(sbase + OFF_B + stage * B_STAGE_BYTES, &map_b, full_b, kb * BLOCK_K, x0, y0, ic.b);
                        if (++stage == B_STAGES) { stage = 0; bphase ^= 1; }
                    }
                }
            }
            if (PROF && P.prof) {
                unsigned long long* o = P.prof + (size_t)blockIdx.x * 16;
                o[0] = pw0; o[1] = pw1;
            }
        }
        __syncwarp();
    } else if (warp == 1) {
        // ============================== MMA issuer (leader CTA, one lane) ====================
        if (leader && lane == 0) {
            constexpr uint32_t idesc = make_idesc(BLOCK_M * CG, TILE_N);
            uint32_t stage = 0, bphase = 0, aphase = 0, acc = 0, tphase = 0;
            for (int item = worker; item < P.n_items; item += n_workers) {
                const ItemCoord ic = decode_item(P, item);
                mbar_wait_p<PROF>(bar_a_full, aphase, pw0);
                aphase ^= 1;
                tc_fence_after();
                const int t0 = ic.chunk * P.tiles_per_item;
                const int t1 = min(t0 + P.tiles_per_item, P.ntiles);
                for (int t = t0; t < t1; ++t) {
                    mbar_wait_p<PROF>(bar_t_empty + 8 * acc, tphase ^ 1, pw1);
                    tc_fence_after();
                    const uint32_t d_tmem = tmem_base + acc * TILE_N;
                    for (int kb = 0; kb < P.kb; ++kb) {
                        mbar_wait_p<PROF>(bar_b_full + 8 * stage, bphase, pw2);
                        tc_fence_after();
                        const uint64_t adesc = make_smem_desc(sbase + OFF_A + kb * A_KB_BYTES);
                        const uint64_t bdesc = make_smem_desc(sbase + OFF_B + stage * B_STAGE_BYTES);
#pragma unroll
                        for (int k = 0; k < BLOCK_K / UMMA_K; ++k) {
                            // advance both start addresses by k * 16 elements * 2 B = 32 B (>> 4 = 2)
                            umma_bf16<CG>(d_tmem, adesc + (uint64_t)(2 * k), bdesc + (uint64_t)(2 * k), idesc,
                                          (uint32_t)((kb | k) != 0));
                        }
                        umma_commit<CG>(bar_b_empty + 8 * stage);      // smem stage free once these MMAs retire
                        if (++stage == B_STAGES) { stage = 0; bphase ^= 1; }
                    }
                    umma_commit<CG>(bar_t_full + 8 * acc);             // accumulator ready for the epilogue
                    if (++acc == 2) { acc = 0; tphase ^= 1; }
                }
                umma_commit<CG>(bar_a_empty);                          // resident A may be replaced
            }
            if (PROF && P.prof) {
                unsigned long long* o = P.prof + (size_t)blockIdx.x * 16;
                o[2] = pw0; o[3] = pw1; o[4] = pw2;
            }
        }
        __syncwarp();
    } else {
        // ============================== epilogue (4 warps, one query row per thread) =========
        const int q4 = warp & 3;                       // TMEM lane quarter this warp may read
        const int prow = q4 * 32 + lane;               // row inside the CTA's 128-row block
        const bool store_thread = (warp == 2 && lane == 0);
        const uint32_t st0 = sbase + OFF_ST0, st1 = sbase + OFF_ST1, st2 = sbase + OFF_ST2;
        const uint32_t sw0 = (uint32_t)((prow >> 1) & 3);   // SWIZZLE_64B : 16-B chunk index ^= addr bits [7,9)
        const uint32_t sw1 = (uint32_t)((prow >> 2) & 1);   // SWIZZLE_32B : 16-B chunk index ^= addr bit 7
        uint32_t acc = 0, tphase = 0;
        for (int item = worker; item < P.n_items; item += n_workers) {
            const ItemCoord ic = decode_item(P, item);
            const int row0 = ic.m * m_rows + (int)rank * BLOCK_M;
            const int t0 = ic.chunk * P.tiles_per_item;
            const int t1 = min(t0 + P.tiles_per_item, P.ntiles);
            for (int t = t0; t < t1; ++t) {
                const int ty = t / P.ntx, tx = t - ty * P.ntx;
                mbar_wait_p<PROF>(bar_t_full + 8 * acc, tphase, pw0);
                tc_fence_after();
                {
                    const long long t0 = PROF ? clock64() : 0;
                    if (store_thread) tma_store_wait_read();   // previous tile's stores have drained the stage
                    epi_bar_sync();
                    if (PROF) pw1 += (unsigned long long)(clock64() - t0);
                }
                ++ptiles;
                const uint32_t taddr = tmem_base + ((uint32_t)(q4 * 32) << 16) + acc * TILE_N;
                float l1[2][16];     // level-1 rows of the current 4-row band
                float l2[2][8];      // level-2 rows of the tile
#pragma unroll
                for (int rp = 0; rp < 4; ++rp) {           // patch rows 2rp, 2rp+1
                    uint32_t v[64];
                    if (PROF && (dbg & 32)) {
#pragma unroll
                        for (int c = 0; c < 64; ++c) v[c] = 0x3f800000u + c;
                    } else {
                        tmem_ld64(taddr + rp * 64, v);
                        tmem_ld_wait();
                    }
                    float f[64];
#pragma unroll
                    for (int c = 0; c < 64; ++c) f[c] = __uint_as_float(v[c]) * P.scale;
                    // level 0: two rows of 32 bf16 (64 B) -> [row][query][64 B], 16-B chunks XOR-swizzled
#pragma unroll
                    for (int rr = 0; rr < 2; ++rr) {
                        if (PROF && (dbg & 16)) break;
                        const uint32_t rowaddr = st0 + (uint32_t)((2 * rp + rr) * (BLOCK_M * 64) + prow * 64);
#pragma unroll
                        for (int ch = 0; ch < 4; ++ch) {
                            const float* s = f + rr * 32 + ch * 8;
                            st_shared_v4(rowaddr + ((ch ^ sw0) << 4), pack_bf16(s[0], s[1]), pack_bf16(s[2], s[3]),
                                         pack_bf16(s[4], s[5]), pack_bf16(s[6], s[7]));
                        }
                    }
                    // level 1: 2x2 means, summed in the reference's raster order then / 4 (corr.py:53)
                    float* l1r = l1[rp & 1];
#pragma unroll
                    for (int c = 0; c < 16; ++c)
                        l1r[c] = (((f[2 * c] + f[2 * c + 1]) + f[32 + 2 * c]) + f[32 + 2 * c + 1]) * 0.25f;
                    if (P.levels > 1 && !(PROF && (dbg & 16))) {
                        const uint32_t rowaddr = st1 + (uint32_t)(rp * (BLOCK_M * 32) + prow * 32);
#pragma unroll
                        for (int ch = 0; ch < 2; ++ch) {
                            const float* s = l1r + ch * 8;
                            st_shared_v4(rowaddr + ((ch ^ sw1) << 4), pack_bf16(s[0], s[1]), pack_bf16(s[2], s[3]),
                                         pack_bf16(s[4], s[5]), pack_bf16(s[6], s[7]));
                        }
                    }
                    if (rp & 1) {
                        float* l2r = l2[rp >> 1];
#pragma unroll
                        for (int c = 0; c < 8; ++c)
                            l2r[c] = (((l1[0][2 * c] + l1[0][2 * c + 1]) + l1[1][2 * c]) + l1[1][2 * c + 1]) * 0.25f;
                        if (P.levels > 2 && !(PROF && (dbg & 16))) {
                            const uint32_t a2 = st2 + (uint32_t)((rp >> 1) * (BLOCK_M * 16) + prow * 16);
                            st_shared_v4(a2, pack_bf16(l2r[0], l2r[1]), pack_bf16(l2r[2], l2r[3]),
                                         pack_bf16(l2r[4], l2r[5]), pack_bf16(l2r[6], l2r[7]));
                        }
                    }
                }
                // accumulator stage drained: hand it back to the MMA warp (leader CTA's barrier)
                tc_fence_before();
                __syncwarp();
                if (lane == 0) {
                    if (CG == 2) mbar_arrive_cluster(bar_t_empty + 8 * acc, 0);
                    else mbar_arrive_local(bar_t_empty + 8 * acc);
                }
                // level 3: one 8x8 block mean per 8 columns -> 4 values, stored directly
                if (P.levels > 3 && !(PROF && (dbg & 8))) {
                    float l3[4];
#pragma unroll
                    for (int c = 0; c < 4; ++c)
                        l3[c] = (((l2[0][2 * c] + l2[0][2 * c + 1]) + l2[1][2 * c]) + l2[1][2 * c + 1]) * 0.25f;
                    const int p = row0 + prow;
                    const int x3 = tx * 4, y3 = ty;
                    if (p < P.N && y3 < P.l3_h) {
                        __nv_bfloat16* dst = P.l3_base + ((long long)ic.b * P.N + p) * P.l3_qs + (long long)y3 * P.l3_pitch + x3;
                        if (x3 + 3 < P.l3_w) {
                            *reinterpret_cast<uint2*>(dst) = make_uint2(pack_bf16(l3[0], l3[1]), pack_bf16(l3[2], l3[3]));
                        } else {
#pragma unroll
                            for (int c = 0; c < 4; ++c)
                                if (x3 + c < P.l3_w) dst[c] = __float2bfloat16_rn(l3[c]);
                        }
                    }
                }
                fence_proxy_async_smem();                  // generic-proxy smem writes -> visible to TMA
                epi_bar_sync();
                if (store_thread) {
                    if (row0 < P.N) {
                        if (!(PROF && (dbg & 1))) tma_store_4d(&map_l0, st0, tx * PATCH_W, row0, ty * PATCH_H, ic.b);
                        if (P.levels > 1 && tx * 16 < (P.w >> 1) && ty * 4 < (P.h >> 1) && !(PROF && (dbg & 2)))
                            tma_store_4d(&map_l1, st1, tx * 16, row0, ty * 4, ic.b);
                        if (P.levels > 2 && tx * 8 < (P.w >> 2) && ty * 2 < (P.h >> 2) && !(PROF && (dbg & 4)))
                            tma_store_4d(&map_l2, st2, tx * 8, row0, ty * 2, ic.b);
                    }
                    tma_store_commit();
                }
                if (++acc == 2) { acc = 0; tphase ^= 1; }
            }
        }
        if (store_thread) tma_store_wait_all();
        if (PROF && P.prof && store_thread) {
            unsigned long long* o = P.prof + (size_t)blockIdx.x * 16;
            o[5] = pw0; o[6] = pw1; o[7] = ptiles; o[8] = (unsigned long long)(clock64() - t_start);
        }
        __syncwarp();
    }

    // teardown: nobody may leave (or free TMEM) while a peer can still signal / read this CTA
    tc_fence_before();
    if (CG == 2) cluster_sync_all(); else __syncthreads();
    if (warp == 1) tmem_dealloc<CG>(tmem_base, TMEM_COLS);
}

// ------------------------------------------------------------------------------------ host side
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

bool encode_map(CUtensorMap* m, void* base, int rank, const uint64_t* dims, const uint64_t* strides_bytes,
                const uint32_t* box, CUtensorMapSwizzle swz) {
    PFN_cuTensorMapEncodeTiled_v12000 enc = get_encode();
    if (!enc) return false;
    cuuint32_t estr[5] = {1, 1, 1, 1, 1};
    cuuint64_t gdim[5], gstr[5];
    cuuint32_t bx[5];
    for (int i = 0; i < rank; ++i) { gdim[i] = dims[i]; bx[i] = box[i]; }
    for (int i = 0; i < rank - 1; ++i) gstr[i] = strides_bytes[i];
    CUresult r = enc(m, CU_TENSOR_MAP_DATA_TYPE_BFLOAT16, (cuuint32_t)rank, base, gdim, gstr, bx, estr,
                     CU_TENSOR_MAP_INTERLEAVE_NONE, swz, CU_TENSOR_MAP_L2_PROMOTION_L2_256B,
                     CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
    return r == CUDA_SUCCESS;
}

template <int CG, bool PROF>
int launch_gemm(const CUtensorMap& ma, const CUtensorMap& mb, const CUtensorMap& m0, const CUtensorMap& m1,
                const CUtensorMap& m2, const GemmParams& P, int grid, cudaStream_t st) {
    static bool configured = false;
    if (!configured) {
        OFB_CUDA(cudaFuncSetAttribute(corr_pyramid_kernel<CG, PROF>, cudaFuncAttributeMaxDynamicSharedMemorySize, SMEM_ALLOC));
        configured = true;
    }
    cudaLaunchConfig_t cfg = {};
    cfg.gridDim = dim3(grid);
    cfg.blockDim = dim3(NUM_THREADS);
    cfg.dynamicSmemBytes = SMEM_ALLOC;
    cfg.stream = st;
    cudaLaunchAttribute attr[1];
    attr[0].id = cudaLaunchAttributeClusterDimension;
    attr[0].val.clusterDim.x = CG;
    attr[0].val.clusterDim.y = 1;
    attr[0].val.clusterDim.z = 1;
    cfg.attrs = attr;
    cfg.numAttrs = 1;
    OFB_CUDA(cudaLaunchKernelEx(&cfg, corr_pyramid_kernel<CG, PROF>, ma, mb, m0, m1, m2, P));
    OFB_LAUNCH_CHECK();
    return OFB_OK;
}

}  // namespace

static int corr_pyramid_impl(const void* f1_km, const void* f2_km, const ofb_pyramid* pyr, int B, int C, int h, int w,
                             float scale, int cta_group, unsigned long long* prof, void* stream) {
    if (!f1_km || !f2_km || !pyr || B < 0 || C <= 0 || h <= 0 || w <= 0) return OFB_EINVAL;
    if (cta_group < 0 || cta_group > 2) return OFB_EINVAL;
    if (pyr->levels < 1 || pyr->levels > OFB_MAX_LEVELS) return OFB_EINVAL;
    if (pyr->dtype != OFB_DTYPE_BF16) return OFB_EUNSUPPORTED;          // fp32 pyramids: ofb_corr_pyramid_simt_f32
    if (C % BLOCK_K != 0 || C > MAX_KB * BLOCK_K) return OFB_EUNSUPPORTED;
    if (B == 0) return OFB_OK;
    if ((reinterpret_cast<uintptr_t>(f1_km) | reinterpret_cast<uintptr_t>(f2_km)) & 15) return OFB_EALIGN;
    const int N = h * w;
    for (int l = 0; l < pyr->levels; ++l) {
        if (!pyr->base[l] || pyr->lvl_h[l] != (h >> l) || pyr->lvl_w[l] != (w >> l)) return OFB_EINVAL;
        if (pyr->lvl_h[l] <= 0 || pyr->lvl_w[l] <= 0 || pyr->row_pitch[l] < pyr->lvl_w[l]) return OFB_EINVAL;
        // TMA global strides are multiples of 16 bytes
        if ((pyr->row_pitch[l] & 7) || (pyr->q_stride[l] & 7) || (reinterpret_cast<uintptr_t>(pyr->base[l]) & 15))
            return OFB_EALIGN;
        if (pyr->q_stride[l] < (int64_t)pyr->row_pitch[l] * pyr->lvl_h[l]) return OFB_EINVAL;
    }
    const int cg = cta_group == 0 ? 2 : cta_group;

    GemmParams P = {};
    P.B = B; P.C = C; P.h = h; P.w = w; P.N = N; P.levels = pyr->levels; P.kb = C / BLOCK_K; P.scale = scale;
    P.ntx = (w + PATCH_W - 1) / PATCH_W; P.nty = (h + PATCH_H - 1) / PATCH_H; P.ntiles = P.ntx * P.nty;
    P.mblk = (N + BLOCK_M * cg - 1) / (BLOCK_M * cg);
    const int workers = ofb_num_sms() / cg;
    // split the target tiles of one (batch, query block) into chunks so the last wave is not mostly idle:
    // pick the chunk count (<= 8) that minimises ceil(items / workers) * tiles_per_item
    int best_chunks = 1;
    long long best_cost = -1;
    for (int nc = 1; nc <= 8 && nc <= P.ntiles; ++nc) {
        const int tpi = (P.ntiles + nc - 1) / nc;
        const int real_nc = (P.ntiles + tpi - 1) / tpi;
        const long long items = (long long)B * P.mblk * real_nc;
        const long long waves = (items + workers - 1) / workers;
        const long long cost = waves * (tpi + 1);   // +1: resident-A reload per item
        if (best_cost < 0 || cost < best_cost) { best_cost = cost; best_chunks = real_nc; }
    }
    P.tiles_per_item = (P.ntiles + best_chunks - 1) / best_chunks;
    P.n_chunks = (P.ntiles + P.tiles_per_item - 1) / P.tiles_per_item;
    const long long n_items = (long long)B * P.mblk * P.n_chunks;
    if (n_items > 0x7fffffffLL) return OFB_EUNSUPPORTED;
    P.n_items = (int)n_items;
    if (pyr->levels > 3) {
        P.l3_base = reinterpret_cast<__nv_bfloat16*>(pyr->base[3]);
        P.l3_qs = pyr->q_stride[3]; P.l3_pitch = pyr->row_pitch[3]; P.l3_h = pyr->lvl_h[3]; P.l3_w = pyr->lvl_w[3];
    }

    CUtensorMap ma, mb, ml[3];
    {
        const uint64_t dims[3] = {(uint64_t)C, (uint64_t)N, (uint64_t)B};
        const uint64_t str[2] = {(uint64_t)C * 2, (uint64_t)N * C * 2};
        const uint32_t box[3] = {BLOCK_K, BLOCK_M, 1};
        if (!encode_map(&ma, const_cast<void*>(f1_km), 3, dims, str, box, CU_TENSOR_MAP_SWIZZLE_128B)) return OFB_EDRIVER;
    }
    {
        const uint64_t dims[4] = {(uint64_t)C, (uint64_t)w, (uint64_t)h, (uint64_t)B};
        const uint64_t str[3] = {(uint64_t)C * 2, (uint64_t)w * C * 2, (uint64_t)N * C * 2};
        const uint32_t box[4] = {BLOCK_K, PATCH_W, (uint32_t)(PATCH_H / cg), 1};
        if (!encode_map(&mb, const_cast<void*>(f2_km), 4, dims, str, box, CU_TENSOR_MAP_SWIZZLE_128B)) return OFB_EDRIVER;
    }
    for (int l = 0; l < 3; ++l) {
        const int ll = l < pyr->levels ? l : 0;   // unused maps alias level 0 (never stored through)
        const uint64_t dims[4] = {(uint64_t)pyr->lvl_w[ll], (uint64_t)N, (uint64_t)pyr->lvl_h[ll], (uint64_t)B};
        const uint64_t str[3] = {(uint64_t)pyr->q_stride[ll] * 2, (uint64_t)pyr->row_pitch[ll] * 2,
                                 (uint64_t)N * pyr->q_stride[ll] * 2};
        const uint32_t box[4] = {(uint32_t)(PATCH_W >> ll), BLOCK_M, (uint32_t)(PATCH_H >> ll), 1};
        const CUtensorMapSwizzle swz = ll == 0 ? CU_TENSOR_MAP_SWIZZLE_64B
                                     : ll == 1 ? CU_TENSOR_MAP_SWIZZLE_32B : CU_TENSOR_MAP_SWIZZLE_NONE;
        if (!encode_map(&ml[l], pyr->base[ll], 4, dims, str, box, swz)) return OFB_EDRIVER;
    }
    int grid = workers * cg;
    if ((long long)grid > n_items * cg) grid = (int)(n_items * cg);
    cudaStream_t st = (cudaStream_t)stream;
    P.prof = prof;
    P.dbg = 0;
    if (prof) {
        const char* e = getenv("OFB_K2_DBG");
        if (e) P.dbg = atoi(e);
    }
    if (prof) {
        if (cg == 2) return launch_gemm<2, true>(ma, mb, ml[0], ml[1], ml[2], P, grid, st);
        return launch_gemm<1, true>(ma, mb, ml[0], ml[1], ml[2], P, grid, st);
    }
    if (cg == 2) return launch_gemm<2, false>(ma, mb, ml[0], ml[1], ml[2], P, grid, st);
    return launch_gemm<1, false>(ma, mb, ml[0], ml[1], ml[2], P, grid, st);
}

OFB_API int ofb_corr_pyramid_bf16(const void* f1_km, const void* f2_km, const ofb_pyramid* pyr, int B, int C, int h,
                                  int w, float scale, int cta_group, void* stream) {
    return corr_pyramid_impl(f1_km, f2_km, pyr, B, C, h, w, scale, cta_group, nullptr, stream);
}

OFB_API int ofb_corr_pyramid_bf16_profile(const void* f1_km, const void* f2_km, const ofb_pyramid* pyr, int B, int C,
                                          int h, int w, float scale, int cta_group, uint64_t* prof_dev, void* stream) {
    if (!prof_dev) return OFB_EINVAL;
    return corr_pyramid_impl(f1_km, f2_km, pyr, B, C, h, w, scale, cta_group,
                             reinterpret_cast<unsigned long long*>(prof_dev), stream);
}
