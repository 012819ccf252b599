// res_bf16.cu -- one ResidualLayer application on bf16 NHWC activations as ONE persistent tcgen05 kernel (sm_100a).
//
// Replaces residual.py:18-29 as it is evaluated (the in-place ReLU makes the layer
// relu(x) + W2 . relu(W1 (*) relu(x)); the caller passes r = relu(x) >= 0):
//     out = act( r + W2 . relu( W1 (*) r ) )         W1: 3x3 C -> Cmid (no bias), W2: 1x1 Cmid -> C
// One CTA per SM walks tiles of TW x BH x BN pixels (TW = 16: two M = 128 tiles):
//   GEMM1  D1[m][Cmid] = sum_{9 taps, C/64 chunks} A_tap[128][64] * W1[Cmid][64]^T      (halo tile + shifted descriptors,
//          hconv.cu; W1 and W2 stay RESIDENT in shared memory for the whole kernel: 72 + 16 KB at C=128, Cmid=32)
//   epi1   tcgen05.ld D1 -> ReLU -> bf16 -> the swizzled K-major A operand of GEMM2, in shared memory
//   GEMM2  D2[m][C] = A2[128][Cmid] * W2[C][Cmid]^T
//   epi2   tcgen05.ld D2 -> + r (bf16, re-read through L2) -> ReLU -> bf16 NHWC store
// The MMA issuer software-pipelines across tiles: the first chunk of GEMM1(t+1) is issued before GEMM2(t), so the
// tensor pipe works while the epilogue warps turn D1(t) into A2(t); D1 is double buffered in TMEM.
// The layer is HBM-bound at the cfg3 shape (268 MB per application: 41 us at the measured copy peak) and its
// N = Cmid = 32 MMAs run at 40 cycles instead of the 16-cycle floor (profiles/r02_ubench_mma_rate.txt), which puts the
// tensor time of a tile (3.1 us at 1.9 GHz) right at its HBM time (3.3 us): both pipes are busy.
#include <cstdlib>
#include <cstring>

#include "common.cuh"
#include "ptx.cuh"
#include "bf16_common.cuh"

int hconv_plan_rows(int kind, int Cin, int Cout);

namespace {

constexpr int RB_THREADS = 384;
constexpr int RB_MAX_CHUNKS = 2;

struct ResBfParams {
    const __nv_bfloat16 *r;     // layer input = skip, NHWC (B,H,W,C)
    __nv_bfloat16 *out;         // NHWC (B,H,W,C)
    int B, H, W, C, Cmid;
    int TW, BH, BN, MT, WP;
    int tiles_x, tiles_y, tiles_n;
    long long ntiles;
    int chunks, halo_bytes, halo_stride;
    int relu_out;
    uint32_t a_off16[9];
};

__global__ void __launch_bounds__(RB_THREADS, 1)
res_bf16_kernel(const __grid_constant__ CUtensorMap tma_in, const __grid_constant__ CUtensorMap tma_w1,
                const __grid_constant__ CUtensorMap tma_w2, const __grid_constant__ ResBfParams p) {
    extern __shared__ unsigned char smem_raw[];
    const uint32_t raw = ptx::smem_u32(smem_raw);
    const uint32_t sbase = (raw + 1023u) & ~1023u;
    unsigned char *sm = smem_raw + (sbase - raw);

    const int chunks = p.chunks, Cmid = p.Cmid, C = p.C, MT = p.MT;
    const uint32_t w1_off = (uint32_t)(chunks * p.halo_stride);
    const uint32_t w1_step = (uint32_t)Cmid * 128u;                             // one (chunk, tap) weight tile
    const uint32_t w2_off = w1_off + ((9u * chunks * w1_step + 1023u) & ~1023u);
    const uint32_t a2_off = w2_off + (((uint32_t)C * 128u + 1023u) & ~1023u);
    const uint32_t bar_off = a2_off + (uint32_t)MT * 16384u;
    const uint32_t bars = sbase + bar_off;
    auto hfull = [&](int c) { return bars + 8u * c; };
    auto hempty = [&](int c) { return bars + 8u * (2 + c); };
    const uint32_t wfull = bars + 8u * 4;
    auto d1full = [&](int s) { return bars + 8u * (5 + s); };
    auto d1empty = [&](int s) { return bars + 8u * (7 + s); };
    const uint32_t a2ready = bars + 8u * 9, d2full = bars + 8u * 10, d2empty = bars + 8u * 11;
    volatile uint32_t *tmem_holder = reinterpret_cast<volatile uint32_t *>(sm + bar_off + 128);

    const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
    if (tid == 0) {
        for (int c = 0; c < 2; ++c) { ptx::mbar_init(hfull(c), 1); ptx::mbar_init(hempty(c), 1); }
        ptx::mbar_init(wfull, 1);
        for (int s = 0; s < 2; ++s) { ptx::mbar_init(d1full(s), 1); ptx::mbar_init(d1empty(s), 8); }
        ptx::mbar_init(a2ready, 8); ptx::mbar_init(d2full, 1); ptx::mbar_init(d2empty, 8);
        ptx::fence_mbar_init();
    }
    if (tid == 32) { ptx::prefetch_tmap(&tma_in); ptx::prefetch_tmap(&tma_w1); ptx::prefetch_tmap(&tma_w2); }
    if (warp == 2) ptx::tmem_alloc(sbase + bar_off + 128, 512);
    ptx::tc_fence_before();
    __syncthreads();
    ptx::tc_fence_after();
    const uint32_t tmem_base = *tmem_holder;
    pdl_launch_dependents();

    const long long ntiles = p.ntiles;
    const int G = (int)gridDim.x;
    const uint32_t D2COL = 256;                 // D1: 2 stages x MT x Cmid columns from 0; D2: MT x C columns from 256

    if (warp == 0) {
        // ===================== halo producer (buffer c = chunk c of the current tile) =====================
        const bool leader = ptx::elect_one();
        pdl_wait();
        uint32_t par = 0;
        for (long long tile = blockIdx.x; tile < ntiles; tile += G, par ^= 1) {
            long long t = tile;
            const int tx = (int)(t % p.tiles_x); t /= p.tiles_x;
            const int ty = (int)(t % p.tiles_y); t /= p.tiles_y;
            const int gx0 = tx * p.TW, gy0 = ty * p.BH, n0 = (int)t * p.BN;
            for (int c = 0; c < chunks; ++c) {
                ptx::mbar_wait(hempty(c), par ^ 1);
                if (leader) {
                    ptx::mbar_expect_tx(hfull(c), (uint32_t)p.halo_bytes);
                    tma_load_5d(sbase + (uint32_t)(c * p.halo_stride), &tma_in, hfull(c), c * 64, gx0 - 1, n0, 0, gy0 - 1);
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
    } else if (warp == 1) {
        // ===================== MMA issuer =====================
        const bool leader = ptx::elect_one();
        const uint32_t a_hi = ptx::desc_hi_sw128((uint32_t)(p.WP * 128)), k_hi = ptx::desc_hi_sw128(1024);
        const uint32_t idesc1 = ptx::instr_desc(ptx::FMT_BF16, 128, (uint32_t)Cmid);
        const uint32_t idesc2 = ptx::instr_desc(ptx::FMT_BF16, 128, (uint32_t)C);
        const uint32_t halo16 = sbase >> 4, hstride16 = (uint32_t)p.halo_stride >> 4;
        const uint32_t w1_16 = (sbase + w1_off) >> 4, w1s16 = w1_step >> 4, w2_16 = (sbase + w2_off) >> 4, a2_16 = (sbase + a2_off) >> 4;
        // GEMM1 of one chunk of tile number `it` (D1 stage it & 1)
        auto gemm1_chunk = [&](int it, int c) {
            ptx::mbar_wait(hfull(c), (uint32_t)(it & 1));
            ptx::tc_fence_after();
            const uint32_t d1 = tmem_base + (uint32_t)((it & 1) * MT * Cmid);
#pragma unroll
            for (int t = 0; t < 9; ++t) {
                const uint32_t a_lo = halo16 + (uint32_t)c * hstride16 + p.a_off16[t];
                const uint32_t b_lo = w1_16 + (uint32_t)(c * 9 + t) * w1s16;
                for (int m = 0; m < MT; ++m) {
#pragma unroll
                    for (int kk = 0; kk < 4; ++kk)
                        if (leader) mma_bf16_w(d1 + (uint32_t)(m * Cmid), a_lo + (uint32_t)(m * 64) + 2u * kk, a_hi, b_lo + 2u * kk, k_hi, idesc1,
                                               (c == 0 && t == 0 && kk == 0) ? 0u : 1u);
                }
            }
            if (leader) ptx::tc_commit(hempty(c));           // this chunk's halo buffer may take the next tile
            __syncwarp();
        };
        ptx::mbar_wait(wfull, 0);
        int it = 0;
        long long tile = blockIdx.x;
        if (tile < ntiles) {
            // prologue: all of GEMM1(0)
            for (int c = 0; c < chunks; ++c) gemm1_chunk(0, c);
            if (leader) ptx::tc_commit(d1full(0));
            __syncwarp();
        }
        for (; tile < ntiles; tile += G, ++it) {
            const bool more = tile + G < ntiles;
            if (more) {
                // first chunk of GEMM1(t+1) runs while the epilogue turns D1(t) into A2(t)
                ptx::mbar_wait(d1empty((it + 1) & 1), (uint32_t)((((it + 1) >> 1) & 1) ^ 1));
                gemm1_chunk(it + 1, 0);
            }
            ptx::mbar_wait(a2ready, (uint32_t)(it & 1));
            ptx::mbar_wait(d2empty, (uint32_t)((it & 1) ^ 1));
            ptx::tc_fence_after();
            for (int m = 0; m < MT; ++m)
                for (int kk = 0; kk < Cmid / 16; ++kk)
                    if (leader) mma_bf16_w(tmem_base + D2COL + (uint32_t)(m * C), a2_16 + (uint32_t)(m * 1024) + 2u * kk, k_hi, w2_16 + 2u * kk, k_hi,
                                           idesc2, kk > 0 ? 1u : 0u);
            if (leader) ptx::tc_commit(d2full);
            __syncwarp();
            if (more) {
                for (int c = 1; c < chunks; ++c) gemm1_chunk(it + 1, c);
                if (leader) ptx::tc_commit(d1full((it + 1) & 1));
                __syncwarp();
            }
        }
    } else if (warp >= 4) {
        // ===================== epilogue warps: epi1 then epi2 of every tile =====================
        const int q = warp & 3, g = (warp - 4) >> 2;
        const int row = q * 32 + lane;
        const int em = MT == 2 ? g : 0;
        const int xx = row & 7, grp = row >> 3;
        const int bn = grp % p.BN, yy = grp / p.BN;
        const uint32_t lane_t = tmem_base + ((uint32_t)(q * 32) << 16);
        // epi2 columns of this thread: MT == 2 -> the whole row of M-tile g; MT == 1 -> column half g
        const int c2_lo = MT == 2 ? 0 : g * (C / 2), c2_hi = MT == 2 ? C : (g + 1) * (C / 2);
        pdl_wait();                                 // the skip tensor is the previous layer's output
        int it = 0;
        for (long long tile = blockIdx.x; tile < ntiles; tile += G, ++it) {
            long long t = tile;
            const int tx = (int)(t % p.tiles_x); t /= p.tiles_x;
            const int ty = (int)(t % p.tiles_y); t /= p.tiles_y;
            const int gx = tx * p.TW + em * 8 + xx, gy = ty * p.BH + yy, n = (int)t * p.BN + bn;
            const bool valid = gx < p.W && gy < p.H && n < p.B;
            const long long ob = (((long long)n * p.H + gy) * p.W + gx) * C;
            // ---- epilogue 1: relu(D1) -> A2 (bf16, K-major, 128-byte swizzle) ----
            ptx::mbar_wait(d1full(it & 1), (uint32_t)((it >> 1) & 1));
            ptx::tc_fence_after();
            if (MT == 2 || g == 0) {
                unsigned char *arow = sm + a2_off + em * 16384 + row * 128;
                const uint32_t d1 = lane_t + (uint32_t)((it & 1) * MT * Cmid + em * Cmid);
                for (int c0 = 0; c0 < Cmid; c0 += 16) {
                    float v[16];
                    tmem_ld16(d1 + (uint32_t)c0, v);
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
            }
            ptx::fence_proxy_async();               // generic-proxy smem writes -> visible to the tensor core
            ptx::tc_fence_before();
            __syncwarp();
            if (lane == 0) { ptx::mbar_arrive(a2ready); ptx::mbar_arrive(d1empty(it & 1)); }
            // ---- epilogue 2: D2 + skip -> ReLU -> bf16 NHWC ----
            // (the skip pixels were fetched by this tile's halo load a moment ago: these reads hit L2)
            ptx::mbar_wait(d2full, (uint32_t)(it & 1));
            ptx::tc_fence_after();
            {
                const uint32_t d2 = lane_t + D2COL + (uint32_t)(em * C);
                float va[32], vb[32];
                auto emit = [&](const float (&v)[32], int c0) {
                    if (!valid) return;
                    const uint4 *sk = reinterpret_cast<const uint4 *>(p.r + ob + c0);
                    uint4 *dst = reinterpret_cast<uint4 *>(p.out + ob + c0);
#pragma unroll
                    for (int i = 0; i < 32; i += 8) {
                        const uint4 s4 = __ldg(sk + (i >> 3));
                        const uint32_t sw[4] = {s4.x, s4.y, s4.z, s4.w};
                        uint32_t ow[4];
#pragma unroll
                        for (int u = 0; u < 4; ++u) {
                            const __nv_bfloat162 sb = *reinterpret_cast<const __nv_bfloat162 *>(&sw[u]);
                            float o0 = v[i + 2 * u] + __low2float(sb), o1 = v[i + 2 * u + 1] + __high2float(sb);
                            if (p.relu_out) { o0 = fmaxf(o0, 0.f); o1 = fmaxf(o1, 0.f); }
                            ow[u] = pack_bf16(o0, o1);
                        }
                        dst[i >> 3] = make_uint4(ow[0], ow[1], ow[2], ow[3]);
                    }
                };
                ptx::tmem_ld32(d2 + (uint32_t)c2_lo, va);
                for (int c0 = c2_lo; c0 < c2_hi; c0 += 64) {
                    ptx::tmem_ld_wait32(va);
                    if (c0 + 32 < c2_hi) ptx::tmem_ld32(d2 + (uint32_t)(c0 + 32), vb);
                    emit(va, c0);
                    if (c0 + 32 < c2_hi) {
                        ptx::tmem_ld_wait32(vb);
                        if (c0 + 64 < c2_hi) ptx::tmem_ld32(d2 + (uint32_t)(c0 + 64), va);
                        emit(vb, c0 + 32);
                    }
                }
            }
            ptx::tc_fence_before();
            __syncwarp();
            if (lane == 0) ptx::mbar_arrive(d2empty);
        }
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
    if ((reinterpret_cast<uintptr_t>(r) | reinterpret_cast<uintptr_t>(out) | reinterpret_cast<uintptr_t>(w1_packed) |
         reinterpret_cast<uintptr_t>(w2_packed)) & 15) return VQB_ERR_ALIGNMENT;
    const int rows1 = hconv_plan_rows(VQB_CONV_K3, C, Cmid), rows2 = hconv_plan_rows(VQB_RES_W2_KIND, Cmid, C);
    if (rows1 < 0 || rows2 < 0) return VQB_ERR_UNSUPPORTED;
    cudaStream_t s = (cudaStream_t)stream;
    ResBfParams q;
    memset(&q, 0, sizeof(q));
    q.r = reinterpret_cast<const __nv_bfloat16 *>(r); q.out = reinterpret_cast<__nv_bfloat16 *>(out);
    q.B = B; q.H = H; q.W = W; q.C = C; q.Cmid = Cmid; q.relu_out = relu_out;
    q.chunks = C / 64;
    q.MT = W > 8 ? 2 : 1;
    q.TW = 8 * q.MT;
    q.BH = rb_pow2_ceil(H) < 16 ? rb_pow2_ceil(H) : 16;
    q.BN = 16 / q.BH;
    q.WP = q.TW + 2;
    q.tiles_x = (W + q.TW - 1) / q.TW; q.tiles_y = (H + q.BH - 1) / q.BH; q.tiles_n = (B + q.BN - 1) / q.BN;
    q.ntiles = (long long)q.tiles_x * q.tiles_y * q.tiles_n;
    q.halo_bytes = (q.BH + 2) * q.BN * q.WP * 128;
    q.halo_stride = (q.halo_bytes + 1023) & ~1023;
    for (int t = 0; t < 9; ++t) q.a_off16[t] = (uint32_t)(((t / 3) * q.BN * q.WP + (t % 3)) * 8);     // tap (r,s): dy+1 = r, dx+1 = s

    CUtensorMap tin, tw1, tw2;
    {
        typedef unsigned long long u64;
        const u64 dims[5] = {(u64)C, (u64)W, (u64)B, 1, (u64)H};
        const u64 strides[4] = {(u64)C * 2, (u64)H * W * C * 2, (u64)H * W * C * 2, (u64)W * C * 2};
        const uint32_t box[5] = {64u, (uint32_t)q.WP, (uint32_t)q.BN, 1u, (uint32_t)(q.BH + 2)};
        int rc = vqb_encode_tmap_nd(&tin, CU_TENSOR_MAP_DATA_TYPE_BFLOAT16, r, 5, dims, strides, box, CU_TENSOR_MAP_SWIZZLE_128B);
        if (rc) return rc;
        rc = vqb_encode_tmap_2d(&tw1, CU_TENSOR_MAP_DATA_TYPE_BFLOAT16, w1_packed, 64, (uint64_t)rows1, 128, 64, (uint32_t)Cmid,
                                CU_TENSOR_MAP_SWIZZLE_128B);
        if (rc) return rc;
        rc = vqb_encode_tmap_2d(&tw2, CU_TENSOR_MAP_DATA_TYPE_BFLOAT16, w2_packed, 64, (uint64_t)rows2, 128, 64, (uint32_t)C,
                                CU_TENSOR_MAP_SWIZZLE_128B);
        if (rc) return rc;
    }
    const int w1_bytes = (9 * q.chunks * Cmid * 128 + 1023) & ~1023, w2_bytes = (C * 128 + 1023) & ~1023;
    const int smem = q.chunks * q.halo_stride + w1_bytes + w2_bytes + q.MT * 16384 + 256 + 1024;
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
    if (cudaError_t le = vqb_launch(res_bf16_kernel, dim3((unsigned)grid), dim3(RB_THREADS), (size_t)smem, s, tin, tw1, tw2, q)) return (int)le;
    VQB_COUNT_LAUNCH(1);
    return vqb_cuda_status(cudaGetLastError());
}
