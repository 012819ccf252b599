// res_tc.cu -- one ResidualLayer application as ONE tcgen05 kernel (sm_100a, TF32 operands).
//
// Replaces residual.py:18-29 as it is actually evaluated (the in-place ReLU makes the
// layer relu(x) + W2 . relu(W1 (*) relu(x)); the caller passes r = relu(x) >= 0):
//     out = act( r + W2 . relu( W1 (*) r ) )          W1: 3x3 C->Cmid (no bias), W2: 1x1 Cmid->C
// with act = ReLU when the next consumer applies one first (always inside a ResidualStack).
// Five PyTorch launches (relu, conv, relu, conv, add) and the (B,Cmid,H,W) intermediate's HBM
// round trip become two chained GEMMs inside one CTA per 128-pixel tile:
//   GEMM1  D1[128][Cmid] = sum_{9 taps, C/32 chunks} A[128][32] * W1[Cmid][32]^T   (A = one halo tile per
//          chunk, the nine taps are shifted UMMA descriptors into it, as conv_halo.cu; W1 streams)
//   epi1   tcgen05.ld D1 -> ReLU -> written back to shared memory as the K-major, 128B-swizzled
//          A operand of GEMM2 (one 128-byte row per pixel per 32 channels)
//   GEMM2  D2[128][C] = A2[128][Cmid] * W2[C][Cmid]^T
//   epi2   tcgen05.ld D2 -> + r (skip, from global/L2) -> ReLU -> NHWC store
#include <cstdlib>

#include "common.cuh"
#include "ptx.cuh"

// in-kernel timeline of CTA 0 (globaltimer ns), read back with vqb_debug_read_trace (diagnostic only)
__device__ unsigned long long g_vqb_trace[32];
__device__ __forceinline__ void trace_mark(int i) {
    if (blockIdx.x == 0) {
        unsigned long long t;
        asm volatile("mov.u64 %0, %%globaltimer;" : "=l"(t));
        g_vqb_trace[i] = t;
    }
}
extern "C" int vqb_debug_read_trace(unsigned long long *dst, int n) {
    if (!dst || n < 1 || n > 32) return VQB_ERR_BAD_ARG;
    return vqb_cuda_status(cudaMemcpyFromSymbol(dst, g_vqb_trace, sizeof(unsigned long long) * n));
}

namespace {

constexpr int RT_THREADS = 256;
constexpr int RT_MAX_STAGES = 32;     // W1 tile ring, as deep as shared memory allows
constexpr int RT_GROUP = 4;           // ring stages released per tcgen05.commit (a commit costs ~230 cycles)
constexpr int RT_HALO_BUFS = 2;       // double-buffered halo tiles
constexpr int RT_A_BYTES = 128 * 128;
constexpr int RT_MAX_CHUNKS = 8;
constexpr int RT_WP = 16;              // padded tile width (8 pixels + halo), multiple of 8

struct ResTcParams {
    const float *skip;      // r, NHWC (B,H,W,C)
    float *out;             // NHWC (B,H,W,C)
    int B, H, W, C, Cmid;
    int BH, BN, tiles_x, tiles_y;          // tile = 8 px wide x (BH rows x BN images = 16)
    int stages;
    int relu_out;
    int flags;              // perf experiments (env VQB_RES_FLAGS): 1 = skip GEMM1 MMAs, 2 = skip W1 loads/waits
};

__global__ void __launch_bounds__(RT_THREADS)
res_tc_kernel(const __grid_constant__ CUtensorMap tma_in, const __grid_constant__ CUtensorMap tma_w1,
              const __grid_constant__ CUtensorMap tma_w2, const ResTcParams p) {
    extern __shared__ unsigned char smem_raw[];
    const uint32_t raw = ptx::smem_u32(smem_raw);
    const uint32_t sbase = (raw + 1023u) & ~1023u;
    unsigned char *sm = smem_raw + (sbase - raw);

    const int S = p.stages;
    const int chunks = p.C / 32;
    const int halo_bytes = (p.BH + 2) * p.BN * RT_WP * 128; // per 32-channel chunk
    const int stage_bytes = p.Cmid * 128;                   // W1 tile of one (tap, chunk)
    const int matoms = p.Cmid / 32;                         // 128-byte atoms of the GEMM2 K dimension
    const int hbufs = chunks < RT_HALO_BUFS ? chunks : RT_HALO_BUFS;
    const uint32_t ring_off = (uint32_t)(hbufs * halo_bytes);
    const uint32_t a2_off = ring_off + (uint32_t)(S * stage_bytes);    // A2: matoms x [128 rows][128 B]
    const uint32_t w2_off = a2_off + (uint32_t)(matoms * RT_A_BYTES);   // W2: matoms x [C rows][128 B]
    const uint32_t bar_off = w2_off + (uint32_t)(matoms * p.C * 128);
    const uint32_t bars = sbase + bar_off;
    auto full = [&](int s) { return bars + 8u * s; };
    auto empty = [&](int s) { return bars + 8u * (RT_MAX_STAGES + s); };
    const uint32_t w2full = bars + 8u * (2 * RT_MAX_STAGES + 0);
    const uint32_t d1full = bars + 8u * (2 * RT_MAX_STAGES + 1);
    const uint32_t a2ready = bars + 8u * (2 * RT_MAX_STAGES + 2);
    const uint32_t d2full = bars + 8u * (2 * RT_MAX_STAGES + 3);
    auto hfull = [&](int b) { return bars + 8u * (2 * RT_MAX_STAGES + 4 + b); };
    auto hempty = [&](int b) { return bars + 8u * (2 * RT_MAX_STAGES + 4 + RT_HALO_BUFS + b); };
    constexpr int RT_MISC = 8 * (2 * RT_MAX_STAGES + 4 + 2 * RT_HALO_BUFS);
    volatile uint32_t *tmem_holder = reinterpret_cast<volatile uint32_t *>(sm + bar_off + RT_MISC);

    const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
    if (tid == 128) trace_mark(0);                     // kernel entry
    int tcols = 32;
    while (tcols < p.Cmid + p.C) tcols <<= 1;
    const uint32_t d2col = (uint32_t)p.Cmid;                // D1 at column 0, D2 right after it

    int tile = blockIdx.x;
    const int tx = tile % p.tiles_x; tile /= p.tiles_x;
    const int ty = tile % p.tiles_y; tile /= p.tiles_y;
    const int gx0 = tx * 8, gy0 = ty * p.BH, n0 = tile * p.BN;

    if (tid == 0) {
        ptx::prefetch_tmap(&tma_in);
        ptx::prefetch_tmap(&tma_w1);
        ptx::prefetch_tmap(&tma_w2);
        for (int s = 0; s < S; ++s) { ptx::mbar_init(full(s), 1); ptx::mbar_init(empty(s), 1); }
        ptx::mbar_init(w2full, 1);
        ptx::mbar_init(d1full, 1);
        ptx::mbar_init(a2ready, 4);                         // one arrival per epilogue warp
        ptx::mbar_init(d2full, 1);
        for (int b = 0; b < RT_HALO_BUFS; ++b) { ptx::mbar_init(hfull(b), 1); ptx::mbar_init(hempty(b), 1); }
        ptx::fence_mbar_init();
    }
    if (warp == 6) ptx::tmem_alloc(sbase + bar_off + RT_MISC, (uint32_t)tcols);
    ptx::tc_fence_before();
    __syncthreads();
    ptx::tc_fence_after();
    const uint32_t tmem_base = *tmem_holder;
    if (tid == 128) trace_mark(1);                     // barriers + TMEM ready

    const int ksteps = 9 * chunks;             // chunk-major: the first MMAs need only halo chunk 0

    // Warp roles: 0-3 epilogue (TMEM lane quadrant = warp id), 4 TMA producer, 5 MMA issuer, 6 TMEM
    // allocator.  The single-thread issuers get the HIGHER warp ids of their sub-partitions on purpose:
    // the scheduler favours high warp ids, and the epilogue warps poll barriers for most of the kernel.
    if (warp == 4) {
        if (lane == 0) {
            auto load_halo = [&](int c) {               // input tile + halo of one 32-channel chunk
                const int b = c % RT_HALO_BUFS;
                if (c >= RT_HALO_BUFS) ptx::mbar_wait(hempty(b), (uint32_t)(((c / RT_HALO_BUFS) - 1) & 1));
                ptx::mbar_expect_tx(hfull(b), (uint32_t)halo_bytes);
                ptx::tma_load_4d(sbase + b * halo_bytes, &tma_in, hfull(b), c * 32, gx0 - 1, n0, gy0 - 1);
            };
            for (int c = 0; c < hbufs; ++c) load_halo(c);
            ptx::mbar_expect_tx(w2full, (uint32_t)(matoms * p.C * 128));      // W2 (all of it) once
            for (int a = 0; a < matoms; ++a)
                ptx::tma_load_2d(sbase + w2_off + a * p.C * 128, &tma_w2, w2full, a * 32, 0);
            for (int i = 0; i < ksteps; ++i) {          // W1 tiles stream through the ring
                const int s = i % S;
                const uint32_t par = (uint32_t)((i / S) & 1);
                const int c = i / 9, t = i - c * 9;
                if (t == 0 && c >= 1 && c + 1 < chunks && c + 1 >= RT_HALO_BUFS) load_halo(c + 1);
                if (p.flags & 2) continue;
                if (s % RT_GROUP == 0) ptx::mbar_wait(empty(s / RT_GROUP), par ^ 1);
                ptx::mbar_expect_tx(full(s), (uint32_t)stage_bytes);
                ptx::tma_load_2d(sbase + ring_off + s * stage_bytes, &tma_w1, full(s), c * 32, t * p.Cmid);
            }
        }
    } else if (warp == 5) {
        if (lane == 0) {
            const uint32_t idesc1 = ptx::instr_desc(ptx::FMT_TF32, 128, (uint32_t)p.Cmid);
            const uint32_t idesc2 = ptx::instr_desc(ptx::FMT_TF32, 128, (uint32_t)p.C);
            for (int i = 0; i < ksteps; ++i) {
                const int s = i % S;
                const uint32_t par = (uint32_t)((i / S) & 1);
                const int c = i / 9, t = i - c * 9;
                const int dy = t / 3 - 1, dx = t % 3 - 1;   // 3x3, pad 1
                const int hb = c % RT_HALO_BUFS;
                if (t == 0) ptx::mbar_wait(hfull(hb), (uint32_t)((c / RT_HALO_BUFS) & 1));
                if (!(p.flags & 2)) ptx::mbar_wait(full(s), par);
                if (i == 0) trace_mark(2);             // first halo chunk + first W1 tile landed
                if (t == 0 && c > 0) trace_mark(2 + c);  // chunk c available
                ptx::tc_fence_after();
                // tap (dy,dx) = the halo tile read (dy+1) padded rows and (dx+1) pixels further in;
                // 8-pixel groups stay one padded row (RT_WP*128 B) apart.  base_offset stays 0: the
                // tensor core derives the swizzle phase from the absolute shared-memory address.
                const uint32_t a = sbase + hb * halo_bytes + (uint32_t)(((dy + 1) * p.BN * RT_WP + (dx + 1)) * 128);
                const uint32_t b = sbase + ring_off + s * stage_bytes;
#pragma unroll
                for (int kk = 0; kk < ((p.flags & 1) ? 0 : 4); ++kk)
                    ptx::mma_tf32(tmem_base, ptx::smem_desc_sw128_sbo(a + kk * 32, RT_WP * 128),
                                  ptx::smem_desc_sw128(b + kk * 32), idesc1, (i > 0 || kk > 0) ? 1u : 0u);
                if (!(p.flags & 2) && (s % RT_GROUP == RT_GROUP - 1 || i == ksteps - 1)) ptx::tc_commit(empty(s / RT_GROUP));
                if (t == 8) ptx::tc_commit(hempty(hb));        // chunk done: its halo buffer may be refilled
            }
            ptx::tc_commit(d1full);
            trace_mark(8);                             // all GEMM1 MMAs issued
            // GEMM2 once the epilogue has written relu(D1) as the A2 operand
            ptx::mbar_wait(w2full, 0);
            ptx::mbar_wait(a2ready, 0);
            ptx::tc_fence_after();
            for (int a = 0; a < matoms; ++a)
#pragma unroll
                for (int kk = 0; kk < 4; ++kk)
                    ptx::mma_tf32(tmem_base + d2col, ptx::smem_desc_sw128(sbase + a2_off + a * RT_A_BYTES + kk * 32),
                                  ptx::smem_desc_sw128(sbase + w2_off + a * p.C * 128 + kk * 32), idesc2,
                                  (a > 0 || kk > 0) ? 1u : 0u);
            ptx::tc_commit(d2full);
            trace_mark(11);                            // GEMM2 issued
        }
    } else if (warp < 4) {
        const int q = warp & 3;
        const int row = q * 32 + lane;
        const uint32_t lane_taddr = tmem_base + ((uint32_t)(q * 32) << 16);
        // ---- epilogue 1: relu(D1) -> A2 operand in shared memory ----
        ptx::mbar_wait_sleep(d1full, 0);
        ptx::tc_fence_after();
        if (tid == 0) trace_mark(9);                 // GEMM1 complete (epilogue sees D1)
        for (int a = 0; a < matoms; ++a) {
            float v[32];
            ptx::tmem_ld32(lane_taddr + (uint32_t)(a * 32), v);
            ptx::tmem_ld_wait32(v);
            unsigned char *arow = sm + a2_off + a * RT_A_BYTES + row * 128;
#pragma unroll
            for (int c16 = 0; c16 < 8; ++c16) {
                const float4 o = make_float4(fmaxf(v[c16 * 4 + 0], 0.f), fmaxf(v[c16 * 4 + 1], 0.f),
                                             fmaxf(v[c16 * 4 + 2], 0.f), fmaxf(v[c16 * 4 + 3], 0.f));
                *reinterpret_cast<float4 *>(arow + ((c16 ^ (row & 7)) << 4)) = o;
            }
        }
        ptx::fence_proxy_async();       // generic-proxy smem writes -> visible to the tensor core's async proxy
        ptx::tc_fence_before();
        __syncwarp();
        if (lane == 0) ptx::mbar_arrive(a2ready);
        if (tid == 0) trace_mark(10);                // A2 written

        // ---- epilogue 2: D2 + skip -> ReLU -> NHWC store ----
        const int bw = row & 7, grp = row >> 3;             // row = (y * BN + bn) * 8 + x
        const int bn = grp % p.BN, bh = grp / p.BN;
        const int gx = gx0 + bw, gy = gy0 + bh, n = n0 + bn;
        const bool valid = gx < p.W && gy < p.H && n < p.B;
        const long long ob = (((long long)n * p.H + gy) * p.W + gx) * p.C;
        ptx::mbar_wait_sleep(d2full, 0, 64);
        ptx::tc_fence_after();
        if (tid == 0) trace_mark(12);                // GEMM2 complete
        for (int c0 = 0; c0 < p.C; c0 += 32) {
            float v[32];
            ptx::tmem_ld32(lane_taddr + d2col + (uint32_t)c0, v);
            ptx::tmem_ld_wait32(v);
            if (valid) {
#pragma unroll
                for (int i = 0; i < 32; i += 4) {
                    const float4 sk = __ldg(reinterpret_cast<const float4 *>(p.skip + ob + c0 + i));
                    float4 o = make_float4(v[i] + sk.x, v[i + 1] + sk.y, v[i + 2] + sk.z, v[i + 3] + sk.w);
                    if (p.relu_out) {
                        o.x = fmaxf(o.x, 0.f); o.y = fmaxf(o.y, 0.f); o.z = fmaxf(o.z, 0.f); o.w = fmaxf(o.w, 0.f);
                    }
                    *reinterpret_cast<float4 *>(p.out + ob + c0 + i) = o;
                }
            }
        }
    }
    if (tid == 0) trace_mark(13);                    // epilogue 2 stores issued
    ptx::tc_fence_before();
    __syncthreads();
    if (warp == 6) ptx::tmem_dealloc(tmem_base, (uint32_t)tcols);
    if (tid == 128) trace_mark(14);                      // exit
}

int rt_pow2_ceil(int x) {
    int p = 1;
    while (p < x) p <<= 1;
    return p;
}

}  // namespace

bool res_tc_supported(int C, int Cmid, const void *r, const void *out) {
    return C % 32 == 0 && C >= 32 && C <= 256 && Cmid % 32 == 0 && Cmid >= 32 && Cmid <= 128 &&
           (reinterpret_cast<uintptr_t>(r) & 15) == 0 && (reinterpret_cast<uintptr_t>(out) & 15) == 0;
}

// w1_tc: [9][Cmid][C], w2_tc: [1][C][Cmid]  (the K-major halves of vqb_pack_conv_weight_f32)
int launch_res_tc(const float *r, const float *w1_tc, const float *w2_tc, float *out, int B, int H, int W, int C,
                  int Cmid, int relu_out, cudaStream_t s) {
    if (!res_tc_supported(C, Cmid, r, out)) return VQB_ERR_UNSUPPORTED;
    ResTcParams q;
    {
        const char *fl = getenv("VQB_RES_FLAGS");
        q.flags = fl ? atoi(fl) : 0;
    }
    q.skip = r; q.out = out; q.B = B; q.H = H; q.W = W; q.C = C; q.Cmid = Cmid; q.relu_out = relu_out;
    q.BH = rt_pow2_ceil(H) < 16 ? rt_pow2_ceil(H) : 16;
    q.BN = 16 / q.BH;
    q.tiles_x = (W + 7) / 8;
    q.tiles_y = (H + q.BH - 1) / q.BH;
    const int tiles_n = (B + q.BN - 1) / q.BN;

    CUtensorMap tin, tw1, tw2;
    // dims ordered (c, w, n, h): the BN images of a tile interleave row by row in shared memory
    const uint64_t dims[4] = {(uint64_t)C, (uint64_t)W, (uint64_t)B, (uint64_t)H};
    const uint64_t strides[3] = {(uint64_t)C * 4, (uint64_t)H * W * C * 4, (uint64_t)W * C * 4};
    const uint32_t box[4] = {32u, (uint32_t)RT_WP, (uint32_t)q.BN, (uint32_t)(q.BH + 2)};
    const uint32_t es[4] = {1u, 1u, 1u, 1u};
    int rc = vqb_encode_tmap_4d(&tin, CU_TENSOR_MAP_DATA_TYPE_FLOAT32, r, dims, strides, box, es,
                                CU_TENSOR_MAP_SWIZZLE_128B);
    if (rc) return rc;
    rc = vqb_encode_tmap_2d(&tw1, CU_TENSOR_MAP_DATA_TYPE_FLOAT32, w1_tc, (uint64_t)C, (uint64_t)9 * Cmid,
                            (uint64_t)C * 4, 32, (uint32_t)Cmid, CU_TENSOR_MAP_SWIZZLE_128B);
    if (rc) return rc;
    rc = vqb_encode_tmap_2d(&tw2, CU_TENSOR_MAP_DATA_TYPE_FLOAT32, w2_tc, (uint64_t)Cmid, (uint64_t)C,
                            (uint64_t)Cmid * 4, 32, (uint32_t)C, CU_TENSOR_MAP_SWIZZLE_128B);
    if (rc) return rc;

    const int stage_bytes = Cmid * 128;
    const int chunks = C / 32;
    const int hbufs = chunks < RT_HALO_BUFS ? chunks : RT_HALO_BUFS;
    const int fixed = hbufs * (q.BH + 2) * q.BN * RT_WP * 128 + (Cmid / 32) * RT_A_BYTES + (Cmid / 32) * C * 128 +
                      8 * (2 * RT_MAX_STAGES + 4 + 2 * RT_HALO_BUFS) + 16 + 1024;
    int stages = (226 * 1024 - fixed) / stage_bytes;
    if (stages > RT_MAX_STAGES) stages = RT_MAX_STAGES;
    if (stages >= 9 * chunks) stages = 9 * chunks;
    else stages -= stages % RT_GROUP;             // a reused ring must hold whole commit groups
    if (stages < RT_GROUP) return VQB_ERR_UNSUPPORTED;
    q.stages = stages;
    const int smem = stages * stage_bytes + fixed;
    static int attr_max = 0;
    if (smem > attr_max) {
        cudaError_t e = cudaFuncSetAttribute(res_tc_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, smem);
        if (e != cudaSuccess) return (int)e;
        attr_max = smem;
    }
    const long long grid = (long long)q.tiles_x * q.tiles_y * tiles_n;
    if (grid <= 0 || grid > 0x7fffffffLL) return VQB_ERR_UNSUPPORTED;
    res_tc_kernel<<<(unsigned)grid, RT_THREADS, smem, s>>>(tin, tw1, tw2, q);
    VQB_COUNT_LAUNCH(1);
    return vqb_cuda_status(cudaGetLastError());
}
