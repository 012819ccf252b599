// res_bf16.cu -- one ResidualLayer application on bf16 NHWC activations as ONE persistent tcgen05 kernel (sm_100a).
//
// Replaces residual.py:18-29 as it is evaluated (the in-place ReLU makes the layer
// relu(x) + W2 . relu(W1 (*) relu(x)); the caller passes r = relu(x) >= 0):
//     out = act( r + W2 . relu( W1 (*) r ) )         W1: 3x3 C -> Cmid (no bias), W2: 1x1 Cmid -> C
// One CTA per SM walks tiles of 8 x BH x BN = 128 pixels:
//   GEMM1  D1[128][Cmid] = sum_{9 taps, C/64 chunks} A_tap[128][64] * W1[Cmid][64]^T     (halo tile + shifted descriptors,
//          hconv.cu; W1 and W2 stay RESIDENT in shared memory for the whole kernel: 72 + 16 KB at C=128, Cmid=32)
//   epi1   tcgen05.ld D1 -> ReLU -> bf16 -> the swizzled K-major A operand of GEMM2, in shared memory (4 warps)
//   GEMM2  D2[128][C] = A2[128][Cmid] * W2[C][Cmid]^T
//   epi2   (8 warps) the tile's skip pixels were TMA-loaded into a staging buffer in the SAME box layout the TMA
//          store uses: tcgen05.ld D2 -> + skip (from shared memory) -> ReLU -> bf16 -> written back in place ->
//          ONE TMA store per 64 channels.  No thread touches global memory: the first version of this kernel
//          read the skip and wrote the result with one 256-byte row per thread (32 cache lines per warp instruction,
//          ~8000 LSU wavefronts per 256-pixel tile) and spent 7.7 us per tile against 3 us of MMAs.
// The MMA issuer software-pipelines across tiles: the first chunk of GEMM1(t+1) is issued before GEMM2(t), so the
// tensor pipe works while epilogue 1 turns D1(t) into A2(t); D1 and D2 are double buffered in TMEM.  A second issuer
// (warp 2) takes K steps 2-3 of every tap into a private accumulator; two epilogue-2 groups (warps 4-11, 16-23) take
// alternate tiles, each with its own staging set.
// The layer is HBM-bound at the cfg3 shape (268 MB per application: 41 us at the measured copy peak) and its
// N = Cmid = 32 MMAs cost ~40 cycles each whoever issues them (4 KB of A per MMA), which puts the tensor time of a tile
// (2.9k cycles) right at its HBM time: both pipes are busy.  66 us = 62 % of the HBM roofline.
#include <cstdlib>
#include <cstring>

#include "common.cuh"
#include "ptx.cuh"
#include "bf16_common.cuh"

int hconv_plan_rows(int kind, int Cin, int Cout);

#if VQB_DIAG
// in-kernel timeline of CTA 0 (SM cycle counter), tools/diag/res_timeline.py; diagnostic builds only
__device__ unsigned long long g_res_tl[32 * 16];
extern "C" int vqb_debug_read_res_timeline(unsigned long long *dst, int n) {
    if (!dst || n < 1 || n > 32 * 16) return VQB_ERR_BAD_ARG;
    return vqb_cuda_status(cudaMemcpyFromSymbol(dst, g_res_tl, sizeof(unsigned long long) * n));
}
#define RB_TL(it_, ev_)                                                                       \
    do {                                                                                      \
        if (blockIdx.x == 0 && (it_) >= 0 && (it_) < 32) {                                    \
            unsigned long long t_;                                                            \
            asm volatile("mov.u64 %0, %%clock64;" : "=l"(t_));                                \
            g_res_tl[(it_) * 16 + (ev_)] = t_;                                                \
        }                                                                                     \
    } while (0)
#else
#define RB_TL(it_, ev_) do { } while (0)
#endif

namespace {

constexpr int RB_THREADS = 768;       // warps: 0 halo producer, 1-2 MMA issuers (2: TMEM + weights first), 3 skip producer, 4-11 and 16-23 the
                                      // two epilogue-2 groups (alternate tiles), 12-15 epilogue 1
constexpr int RB_NHB = 2;             // halo buffers, rotating over the (tile, chunk) sequence (the third one's 23 KB went to the second
                                      // staging set; the loads are L2 hits, prefetched two to three tiles ahead)
constexpr int RB_NISS = 2;            // GEMM1 issuer warps (warp 1: K steps 0-1 of every tap, warp 2: K steps 2-3; private accumulators)

struct ResBfParams {
    int B, H, W, C, Cmid;
    int BH, BN, WP;
    int tiles_x, tiles_y, tiles_n;
    long long ntiles;
    int chunks, halo_bytes, halo_stride;
    int relu_out;
    uint32_t a_off16[9];
};

__device__ __forceinline__ void tma_store_5d(const CUtensorMap *m, uint32_t src, int c0, int c1, int c2, int c3, int c4) {
    asm volatile("cp.async.bulk.tensor.5d.global.shared::cta.bulk_group [%0, {%2, %3, %4, %5, %6}], [%1];" ::
                     "l"(reinterpret_cast<uint64_t>(m)), "r"(src), "r"(c0), "r"(c1), "r"(c2), "r"(c3), "r"(c4) : "memory");
}

__global__ void __launch_bounds__(RB_THREADS, 1)
res_bf16_kernel(const __grid_constant__ CUtensorMap tma_in, const __grid_constant__ CUtensorMap tma_skip,
                const __grid_constant__ CUtensorMap tma_w1, const __grid_constant__ CUtensorMap tma_w2,
                const __grid_constant__ CUtensorMap tma_out, const __grid_constant__ ResBfParams p) {
    extern __shared__ unsigned char smem_raw[];
    const uint32_t raw = ptx::smem_u32(smem_raw);
    const uint32_t sbase = (raw + 1023u) & ~1023u;
    unsigned char *sm = smem_raw + (sbase - raw);

    const int chunks = p.chunks, Cmid = p.Cmid, C = p.C;
    const uint32_t w1_off = (uint32_t)(RB_NHB * p.halo_stride);
    const uint32_t w1_step = (uint32_t)Cmid * 128u;                             // one (chunk, tap) weight tile
    const uint32_t w2_off = w1_off + ((9u * chunks * w1_step + 1023u) & ~1023u);
    const uint32_t a2_off = w2_off + (((uint32_t)C * 128u + 1023u) & ~1023u);
    const uint32_t st_off = a2_off + 16384u;                                    // staging: 2 groups x chunks x [128 px][128 B]
    const uint32_t st_grp = (uint32_t)chunks * 16384u;
    const uint32_t bar_off = st_off + 2u * st_grp;
    const uint32_t bars = sbase + bar_off;
    auto hfull = [&](int b) { return bars + 8u * b; };
    auto hempty = [&](int b) { return bars + 8u * (3 + b); };
    const uint32_t wfull = bars + 8u * 6;
    auto d1full = [&](int s) { return bars + 8u * (7 + s); };
    auto d1empty = [&](int s) { return bars + 8u * (9 + s); };
    const uint32_t a2ready = bars + 8u * 11;
    auto d2full = [&](int s) { return bars + 8u * (12 + s); };
    auto d2empty = [&](int s) { return bars + 8u * (14 + s); };
    auto sfull = [&](int g, int c) { return bars + 8u * (16 + 2 * g + c); };      // per epilogue-2 group and 64-channel chunk
    auto sfree = [&](int g, int c) { return bars + 8u * (20 + 2 * g + c); };
    volatile uint32_t *tmem_holder = reinterpret_cast<volatile uint32_t *>(sm + bar_off + 192);

    const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
    if (tid == 0) {
        // two GEMM1 issuer warps: each commits its own MMAs to the barriers the tensor pipe signals
        for (int b = 0; b < 3; ++b) { ptx::mbar_init(hfull(b), 1); ptx::mbar_init(hempty(b), RB_NISS); }
        ptx::mbar_init(wfull, 1);
        for (int s = 0; s < 2; ++s) {
            ptx::mbar_init(d1full(s), RB_NISS); ptx::mbar_init(d1empty(s), 4);
            ptx::mbar_init(d2full(s), 1); ptx::mbar_init(d2empty(s), 8);
            for (int c = 0; c < 2; ++c) { ptx::mbar_init(sfull(s, c), 1); ptx::mbar_init(sfree(s, c), 1); }
        }
        ptx::mbar_init(a2ready, 4);
        ptx::fence_mbar_init();
    }
    if (tid == 32) { ptx::prefetch_tmap(&tma_in); ptx::prefetch_tmap(&tma_skip); ptx::prefetch_tmap(&tma_w1); }
    if (tid == 64) { ptx::prefetch_tmap(&tma_w2); ptx::prefetch_tmap(&tma_out); }
    if (warp == 2) ptx::tmem_alloc(sbase + bar_off + 192, 512);
    ptx::tc_fence_before();
    __syncthreads();
    ptx::tc_fence_after();
    const uint32_t tmem_base = *tmem_holder;
    pdl_launch_dependents();

    const long long ntiles = p.ntiles;
    const int G = (int)gridDim.x;
    const uint32_t D2COL = 256;                 // D1: 2 stages x Cmid columns from 0; D2: 2 stages x C columns from 256
    const uint32_t D1BCOL = 128;                // the second GEMM1 issuer's partial sums: 2 stages x Cmid columns from 128

    if (warp == 0) {
        // ===================== producer: halo tiles (GEMM1's A operand) and skip tiles (epilogue 2) =====================
        const bool leader = ptx::elect_one();
        pdl_wait();
        uint32_t hb = 0, hpar = 0;
        int it = 0;
        for (long long tile = blockIdx.x; tile < ntiles; tile += G, ++it) {
            long long t = tile;
            const int tx = (int)(t % p.tiles_x); t /= p.tiles_x;
            const int ty = (int)(t % p.tiles_y); t /= p.tiles_y;
            const int gx0 = tx * 8, gy0 = ty * p.BH, n0 = (int)t * p.BN;
            // the halo tiles two and three tiles ahead go to L2 now (3 halo buffers = 1.5 tiles of shared-memory prefetch do
            // not cover an HBM round trip when every SM is streaming)
            for (int ahead = 2; ahead <= 3; ++ahead) {
                long long tp = tile + (long long)ahead * G;
                if (leader && tp < ntiles) {
                    const int px = (int)(tp % p.tiles_x); tp /= p.tiles_x;
                    const int py = (int)(tp % p.tiles_y); tp /= p.tiles_y;
                    for (int c = 0; c < chunks; ++c)
                        asm volatile("cp.async.bulk.prefetch.tensor.5d.L2.global.tile [%0, {%1, %2, %3, %4, %5}];" ::
                                         "l"(reinterpret_cast<uint64_t>(&tma_in)), "r"(c * 64), "r"(px * 8 - 1), "r"((int)tp * p.BN), "r"(0),
                                         "r"(py * p.BH - 1) : "memory");
                }
            }
            for (int c = 0; c < chunks; ++c) {
                ptx::mbar_wait_sleep(hempty((int)hb), hpar ^ 1, 100);
                if (leader) {
                    if (c == 0) RB_TL(it, 0);
                    ptx::mbar_expect_tx(hfull((int)hb), (uint32_t)p.halo_bytes);
                    tma_load_5d(sbase + hb * (uint32_t)p.halo_stride, &tma_in, hfull((int)hb), c * 64, gx0 - 1, n0, 0, gy0 - 1);
                }
                if (++hb == RB_NHB) { hb = 0; hpar ^= 1; }
            }
        }
    } else if (warp == 3) {
        // ===================== skip producer: the tile's own pixels into the staging buffers (epilogue 2 adds them) =====================
        // A separate warp: behind the halo loads in one loop, the wait for the previous tile's store (sfree) kept the NEXT
        // tile's halo request back until epilogue 2 had finished -- the halo then had less than one chunk time to arrive.
        const bool leader = ptx::elect_one();
        pdl_wait();
        int it = 0;
        for (long long tile = blockIdx.x; tile < ntiles; tile += G, ++it) {
            long long t = tile;
            const int tx = (int)(t % p.tiles_x); t /= p.tiles_x;
            const int ty = (int)(t % p.tiles_y); t /= p.tiles_y;
            const int gx0 = tx * 8, gy0 = ty * p.BH, n0 = (int)t * p.BN;
            for (int c = 0; c < chunks; ++c) {
                const int sg = it & 1;                                                   // the epilogue-2 group (and staging set) of this tile
                ptx::mbar_wait_sleep(sfree(sg, c), (uint32_t)(((it >> 1) & 1) ^ 1), 100);  // that group's previous store has read the buffer
                if (leader) {
                    if (c == 0) RB_TL(it, 1);
                    ptx::mbar_expect_tx(sfull(sg, c), 16384u);
                    tma_load_5d(sbase + st_off + (uint32_t)sg * st_grp + (uint32_t)c * 16384u, &tma_skip, sfull(sg, c), c * 64, gx0, n0, 0, gy0);
                }
            }
        }
    } else if (warp == 2) {
        // ===================== resident weights: W1 (9 x chunks tiles of Cmid rows) and W2 (C rows), once =====================
        if (ptx::elect_one()) {
            ptx::mbar_expect_tx(wfull, 9u * chunks * w1_step + (uint32_t)C * 128u);
            for (int i = 0; i < 9 * chunks; ++i)
                ptx::tma_load_2d(sbase + w1_off + (uint32_t)i * w1_step, &tma_w1, wfull, 0, i * Cmid);
            ptx::tma_load_2d(sbase + w2_off, &tma_w2, wfull, 0, 0);
        }
        __syncwarp();
        // ===================== second GEMM1 issuer: K steps 2-3 of every (chunk, tap) into its own accumulator =====================
        // One warp issues an N = 32 MMA per ~39 cycles (its own scalar instruction stream; the pipe needs 16), and GEMM1's
        // 36 x chunks MMAs made the issuer warp the busiest unit of the kernel -- all 4.05k cycles of a tile's period
        // (profiles/r02_res_timeline_before.txt).  Both issuers walk the SAME sequence of halo-buffer uses and both release
        // every one of them (count 2), so neither can get a parity ahead of the other (r02_hconv_notes.txt, section 5).
        {
            const bool leader = ptx::elect_one();
            const uint32_t a_hi = ptx::desc_hi_sw128((uint32_t)(p.WP * 128)), k_hi = ptx::desc_hi_sw128(1024);
            const uint32_t idesc1 = ptx::instr_desc(ptx::FMT_BF16, 128, (uint32_t)Cmid);
            const uint32_t halo16 = sbase >> 4, hstride16 = (uint32_t)p.halo_stride >> 4;
            const uint32_t w1_16 = (sbase + w1_off) >> 4, w1s16 = w1_step >> 4;
            uint32_t hb = 0, hpar = 0;
            ptx::mbar_wait(wfull, 0);
            int it = 0;
            for (long long tile = blockIdx.x; tile < ntiles; tile += G, ++it) {
                ptx::mbar_wait(d1empty(it & 1), (uint32_t)(((it >> 1) & 1) ^ 1));
                const uint32_t d1 = tmem_base + D1BCOL + (uint32_t)((it & 1) * Cmid);
                for (int c = 0; c < chunks; ++c) {
                    ptx::mbar_wait(hfull((int)hb), hpar);
                    ptx::tc_fence_after();
#pragma unroll
                    for (int t = 0; t < 9; ++t) {
                        const uint32_t a_lo = halo16 + hb * hstride16 + p.a_off16[t];
                        const uint32_t b_lo = w1_16 + (uint32_t)(c * 9 + t) * w1s16;
#pragma unroll
                        for (int kk = 2; kk < 4; ++kk)
                            if (leader) mma_bf16_w(d1, a_lo + 2u * kk, a_hi, b_lo + 2u * kk, k_hi, idesc1, (c == 0 && t == 0 && kk == 2) ? 0u : 1u);
                    }
                    if (leader) ptx::tc_commit(hempty((int)hb));
                    __syncwarp();
                    if (++hb == RB_NHB) { hb = 0; hpar ^= 1; }
                }
                if (leader) ptx::tc_commit(d1full(it & 1));
                __syncwarp();
            }
        }
    } else if (warp == 1) {
        // ===================== MMA issuer =====================
        const bool leader = ptx::elect_one();
        const uint32_t a_hi = ptx::desc_hi_sw128((uint32_t)(p.WP * 128)), k_hi = ptx::desc_hi_sw128(1024);
        const uint32_t idesc1 = ptx::instr_desc(ptx::FMT_BF16, 128, (uint32_t)Cmid);
        const uint32_t idesc2 = ptx::instr_desc(ptx::FMT_BF16, 128, (uint32_t)C);
        const uint32_t halo16 = sbase >> 4, hstride16 = (uint32_t)p.halo_stride >> 4;
        const uint32_t w1_16 = (sbase + w1_off) >> 4, w1s16 = w1_step >> 4, w2_16 = (sbase + w2_off) >> 4, a2_16 = (sbase + a2_off) >> 4;
        uint32_t hb = 0, hpar = 0;
        // GEMM1 of chunk c of tile number `it` (D1 stage it & 1); halo buffers are consumed in load order
        auto gemm1_chunk = [&](int it, int c) {
            ptx::mbar_wait(hfull((int)hb), hpar);
            ptx::tc_fence_after();
            if (leader) RB_TL(it, 2 + 2 * (c & 1));
            const uint32_t d1 = tmem_base + (uint32_t)((it & 1) * Cmid);
#pragma unroll
            for (int t = 0; t < 9; ++t) {
                const uint32_t a_lo = halo16 + hb * hstride16 + p.a_off16[t];
                const uint32_t b_lo = w1_16 + (uint32_t)(c * 9 + t) * w1s16;
#pragma unroll
                for (int kk = 0; kk < 2; ++kk)          // (K steps 2-3: the second issuer, warp 2)
                    if (leader) mma_bf16_w(d1, a_lo + 2u * kk, a_hi, b_lo + 2u * kk, k_hi, idesc1, (c == 0 && t == 0 && kk == 0) ? 0u : 1u);
            }
            if (leader) { RB_TL(it, 3 + 2 * (c & 1)); ptx::tc_commit(hempty((int)hb)); }     // this halo buffer may take a later chunk
            __syncwarp();
            if (++hb == RB_NHB) { hb = 0; hpar ^= 1; }
        };
        ptx::mbar_wait(wfull, 0);
        int it = 0;
        long long tile = blockIdx.x;
        if (tile < ntiles) {
            for (int c = 0; c < chunks; ++c) gemm1_chunk(0, c);     // prologue: all of GEMM1(0)
            if (leader) ptx::tc_commit(d1full(0));
            __syncwarp();
        }
        for (; tile < ntiles; tile += G, ++it) {
            const bool more = tile + G < ntiles;
            if (more) {
                // first chunk of GEMM1(t+1) runs while epilogue 1 turns D1(t) into A2(t)
                ptx::mbar_wait(d1empty((it + 1) & 1), (uint32_t)((((it + 1) >> 1) & 1) ^ 1));
                gemm1_chunk(it + 1, 0);
            }
            ptx::mbar_wait(a2ready, (uint32_t)(it & 1));
            if (leader) RB_TL(it, 6);
            ptx::mbar_wait(d2empty(it & 1), (uint32_t)(((it >> 1) & 1) ^ 1));
            ptx::tc_fence_after();
            if (leader) RB_TL(it, 7);
            for (int kk = 0; kk < Cmid / 16; ++kk)
                if (leader) mma_bf16_w(tmem_base + D2COL + (uint32_t)((it & 1) * C), a2_16 + 2u * kk, k_hi, w2_16 + 2u * kk, k_hi, idesc2, kk > 0 ? 1u : 0u);
            if (leader) ptx::tc_commit(d2full(it & 1));
            __syncwarp();
            if (more) {
                for (int c = 1; c < chunks; ++c) gemm1_chunk(it + 1, c);
                if (leader) ptx::tc_commit(d1full((it + 1) & 1));
                __syncwarp();
            }
        }
    } else if (warp >= 12 && warp < 16) {
        // ===================== epilogue 1 (4 warps, one per TMEM lane quadrant): relu(D1) -> A2 =====================
        const int q = warp & 3;
        const int row = q * 32 + lane;
        const uint32_t lane_t = tmem_base + ((uint32_t)(q * 32) << 16);
        unsigned char *arow = sm + a2_off + row * 128;
        int it = 0;
        for (long long tile = blockIdx.x; tile < ntiles; tile += G, ++it) {
            ptx::mbar_wait_sleep(d1full(it & 1), (uint32_t)((it >> 1) & 1), 100);
            ptx::tc_fence_after();
            if (tid == 384) RB_TL(it, 8);
            const uint32_t d1 = lane_t + (uint32_t)((it & 1) * Cmid);
            for (int c0 = 0; c0 < Cmid; c0 += 16) {
                float v[16], w[16];
                tmem_ld16(d1 + (uint32_t)c0, v);
                tmem_ld16(d1 + D1BCOL + (uint32_t)c0, w);                    // + the second issuer's partial sums, fixed order
#pragma unroll
                for (int i = 0; i < 16; ++i) v[i] += w[i];
#pragma unroll
                for (int h = 0; h < 2; ++h) {
                    const uint4 o = make_uint4(pack_bf16(fmaxf(v[h * 8 + 0], 0.f), fmaxf(v[h * 8 + 1], 0.f)),
                                               pack_bf16(fmaxf(v[h * 8 + 2], 0.f), fmaxf(v[h * 8 + 3], 0.f)),
                                               pack_bf16(fmaxf(v[h * 8 + 4], 0.f), fmaxf(v[h * 8 + 5], 0.f)),
                                               pack_bf16(fmaxf(v[h * 8 + 6], 0.f), fmaxf(v[h * 8 + 7], 0.f)));
                    const int c16 = (c0 >> 3) + h;
                    *reinterpret_cast<uint4 *>(arow + ((c16 ^ (row & 7)) << 4)) = o;
                }
            }
            ptx::fence_proxy_async();               // generic-proxy smem writes -> visible to the tensor core
            ptx::tc_fence_before();
            __syncwarp();
            if (tid == 384) RB_TL(it, 9);
            if (lane == 0) { ptx::mbar_arrive(a2ready); ptx::mbar_arrive(d1empty(it & 1)); }
        }
    } else if (warp >= 4) {
        // ===================== epilogue 2 (8 warps): D2 + skip -> ReLU -> bf16, in the staging buffer; TMA store =====================
        // Two groups of eight warps take ALTERNATE tiles (= one D2 stage and one staging set each).  With one group the
        // staging buffer carried store(t) -> skip-tile TMA(t+1) -> epilogue(t+1), 0.7k + 1.5k + 1.5k cycles: a whole tile
        // period (profiles/r02_res_timeline_before.txt).
        const int eg = warp < 12 ? 0 : 1;
        const int w8 = warp - (eg ? 16 : 4);
        const int q = warp & 3, g = w8 >> 2;
        const int row = q * 32 + lane;
        const uint32_t lane_t = tmem_base + ((uint32_t)(q * 32) << 16);
        const int cpg = C / 2;                            // columns per warp group (64 or 32)
        const int col0 = g * cpg;
        const int chunk = col0 >> 6;                      // staging buffer (64 channels) this group writes
        const bool storer = (chunks == 2) ? (q == 0 && lane == 0) : (g == 0 && q == 0 && lane == 0);
        int it = eg;
        for (long long tile = (long long)blockIdx.x + (long long)eg * G; tile < ntiles; tile += 2LL * G, it += 2) {
            long long t = tile;
            const int tx = (int)(t % p.tiles_x); t /= p.tiles_x;
            const int ty = (int)(t % p.tiles_y); t /= p.tiles_y;
            const int gx0 = tx * 8, gy0 = ty * p.BH, n0 = (int)t * p.BN;
            unsigned char *srow = sm + st_off + eg * st_grp + chunk * 16384 + row * 128;
            ptx::mbar_wait_sleep(sfull(eg, chunk), (uint32_t)((it >> 1) & 1), 200);
            if (tid == 128) RB_TL(it, 10);
            ptx::mbar_wait_sleep(d2full(it & 1), (uint32_t)((it >> 1) & 1), 200);
            ptx::tc_fence_after();
            if (tid == 128) RB_TL(it, 11);
            const uint32_t d2 = lane_t + D2COL + (uint32_t)((it & 1) * C);
            for (int c0 = col0; c0 < col0 + cpg; c0 += 32) {
                float v[32];
                ptx::tmem_ld32(d2 + (uint32_t)c0, v);
                ptx::tmem_ld_wait32(v);
#pragma unroll
                for (int i = 0; i < 4; ++i) {
                    const int piece = ((c0 & 63) >> 3) + i;                      // 16-byte piece of the 128-byte row
                    uint4 *ptr = reinterpret_cast<uint4 *>(srow + ((piece ^ (row & 7)) << 4));
                    const uint4 s4 = *ptr;
                    const uint32_t sw[4] = {s4.x, s4.y, s4.z, s4.w};
                    uint32_t ow[4];
#pragma unroll
                    for (int u = 0; u < 4; ++u) {
                        const __nv_bfloat162 sb = *reinterpret_cast<const __nv_bfloat162 *>(&sw[u]);
                        float o0 = v[8 * i + 2 * u] + __low2float(sb), o1 = v[8 * i + 2 * u + 1] + __high2float(sb);
                        if (p.relu_out) { o0 = fmaxf(o0, 0.f); o1 = fmaxf(o1, 0.f); }
                        ow[u] = pack_bf16(o0, o1);
                    }
                    *ptr = make_uint4(ow[0], ow[1], ow[2], ow[3]);
                }
            }
            ptx::tc_fence_before();
            ptx::fence_proxy_async();               // staged tile -> visible to the TMA store
            __syncwarp();
            if (lane == 0) ptx::mbar_arrive(d2empty(it & 1));
            // all warps that wrote this staging buffer are done -> one thread stores it and frees it
            if (chunks == 2) ptx::named_bar_sync(1 + 2 * eg + g, 128);
            else ptx::named_bar_sync(1 + eg, 256);
            if (storer) {
                if (tid == 128) RB_TL(it, 12);
                tma_store_5d(&tma_out, sbase + st_off + (uint32_t)eg * st_grp + (uint32_t)chunk * 16384u, chunk * 64, gx0, n0, 0, gy0);
                asm volatile("cp.async.bulk.commit_group;" ::: "memory");
                asm volatile("cp.async.bulk.wait_group.read 0;" ::: "memory");
                if (tid == 128) RB_TL(it, 13);
                ptx::mbar_arrive(sfree(eg, chunk));
            }
        }
        if (storer) asm volatile("cp.async.bulk.wait_group 0;" ::: "memory");
    }
    ptx::tc_fence_before();
    __syncthreads();
    if (warp == 2) ptx::tmem_dealloc(tmem_base, 512);
}

int rb_pow2_ceil(int x) { int p = 1; while (p < x) p <<= 1; return p; }

}  // namespace

bool res_bf16_supported(int C, int Cmid) {
    return (C == 64 || C == 128) && Cmid % 16 == 0 && Cmid >= 16 && Cmid <= 64;
}

// r, out: bf16 NHWC (B,H,W,C).  w1: packing of kind VQB_CONV_K3 with (Cout = Cmid, Cin = C); w2: VQB_RES_W2_KIND with
// (Cout = C, Cin = Cmid) -- both made by vqb_pack_conv_weight_bf16.
extern "C" int vqb_residual_layer_bf16(const void *r, const void *w1_packed, const void *w2_packed, void *out, int B, int H,
                                       int W, int C, int Cmid, int relu_out, void *stream) {
    if (!r || !w1_packed || !w2_packed || !out) return VQB_ERR_BAD_ARG;
    if (B <= 0 || H <= 0 || W <= 0 || C <= 0 || Cmid <= 0) return VQB_ERR_BAD_ARG;
    if (!res_bf16_supported(C, Cmid)) return VQB_ERR_UNSUPPORTED;
    if (r == out) return VQB_ERR_BAD_ARG;                       // neighbouring tiles read each other's halo: not in place
    if ((reinterpret_cast<uintptr_t>(r) | reinterpret_cast<uintptr_t>(out) | reinterpret_cast<uintptr_t>(w1_packed) |
         reinterpret_cast<uintptr_t>(w2_packed)) & 15) return VQB_ERR_ALIGNMENT;
    const int rows1 = hconv_plan_rows(VQB_CONV_K3, C, Cmid), rows2 = hconv_plan_rows(VQB_RES_W2_KIND, Cmid, C);
    if (rows1 < 0 || rows2 < 0) return VQB_ERR_UNSUPPORTED;
    cudaStream_t s = (cudaStream_t)stream;
    ResBfParams q;
    memset(&q, 0, sizeof(q));
    q.B = B; q.H = H; q.W = W; q.C = C; q.Cmid = Cmid; q.relu_out = relu_out;
    q.chunks = C / 64;
    q.BH = rb_pow2_ceil(H) < 16 ? rb_pow2_ceil(H) : 16;
    q.BN = 16 / q.BH;
    q.WP = 10;
    q.tiles_x = (W + 7) / 8; q.tiles_y = (H + q.BH - 1) / q.BH; q.tiles_n = (B + q.BN - 1) / q.BN;
    q.ntiles = (long long)q.tiles_x * q.tiles_y * q.tiles_n;
    q.halo_bytes = (q.BH + 2) * q.BN * q.WP * 128;
    q.halo_stride = (q.halo_bytes + 1023) & ~1023;
    for (int t = 0; t < 9; ++t) q.a_off16[t] = (uint32_t)(((t / 3) * q.BN * q.WP + (t % 3)) * 8);     // tap (r,s): dy+1 = r, dx+1 = s

    CUtensorMap tin, tskip, tw1, tw2, tout;
    {
        typedef unsigned long long u64;
        const u64 dims[5] = {(u64)C, (u64)W, (u64)B, 1, (u64)H};
        const u64 strides[4] = {(u64)C * 2, (u64)H * W * C * 2, (u64)H * W * C * 2, (u64)W * C * 2};
        const uint32_t box[5] = {64u, (uint32_t)q.WP, (uint32_t)q.BN, 1u, (uint32_t)(q.BH + 2)};
        const uint32_t tbox[5] = {64u, 8u, (uint32_t)q.BN, 1u, (uint32_t)q.BH};
        int rc = vqb_encode_tmap_nd(&tin, CU_TENSOR_MAP_DATA_TYPE_BFLOAT16, r, 5, dims, strides, box, CU_TENSOR_MAP_SWIZZLE_128B);
        if (rc) return rc;
        rc = vqb_encode_tmap_nd(&tskip, CU_TENSOR_MAP_DATA_TYPE_BFLOAT16, r, 5, dims, strides, tbox, CU_TENSOR_MAP_SWIZZLE_128B);
        if (rc) return rc;
        rc = vqb_encode_tmap_nd(&tout, CU_TENSOR_MAP_DATA_TYPE_BFLOAT16, out, 5, dims, strides, tbox, CU_TENSOR_MAP_SWIZZLE_128B);
        if (rc) return rc;
        rc = vqb_encode_tmap_2d(&tw1, CU_TENSOR_MAP_DATA_TYPE_BFLOAT16, w1_packed, 64, (uint64_t)rows1, 128, 64, (uint32_t)Cmid,
                                CU_TENSOR_MAP_SWIZZLE_128B);
        if (rc) return rc;
        rc = vqb_encode_tmap_2d(&tw2, CU_TENSOR_MAP_DATA_TYPE_BFLOAT16, w2_packed, 64, (uint64_t)rows2, 128, 64, (uint32_t)C,
                                CU_TENSOR_MAP_SWIZZLE_128B);
        if (rc) return rc;
    }
    const int w1_bytes = (9 * q.chunks * Cmid * 128 + 1023) & ~1023, w2_bytes = (C * 128 + 1023) & ~1023;
    const int smem = RB_NHB * q.halo_stride + w1_bytes + w2_bytes + 16384 + 2 * q.chunks * 16384 + 256 + 1024;
    if (smem > 227 * 1024) return VQB_ERR_UNSUPPORTED;
    static int attr_max = 0;
    if (smem > attr_max) {
        cudaError_t e = cudaFuncSetAttribute(res_bf16_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, smem);
        if (e != cudaSuccess) return (int)e;
        attr_max = smem;
    }
    int dev = 0, sms = 148;
    cudaGetDevice(&dev);
    cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, dev);
    const int grid = (int)(q.ntiles < sms ? q.ntiles : sms);
    if (cudaError_t le = vqb_launch(res_bf16_kernel, dim3((unsigned)grid), dim3(RB_THREADS), (size_t)smem, s, tin, tskip, tw1, tw2, tout, q)) return (int)le;
    VQB_COUNT_LAUNCH(1);
    return vqb_cuda_status(cudaGetLastError());
}
