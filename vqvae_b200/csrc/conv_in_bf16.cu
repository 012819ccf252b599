// conv_in_bf16.cu -- encoder.py:29-31 for the VQB_BF16 pipeline: Conv2d(3 -> 64, k4 s2 p1) + ReLU, fp32 NCHW image in,
// bf16 NHWC activation out, as ONE persistent tcgen05 kernel (sm_100a).
//
// The layer is HBM-bound (cfg3: 100 MB of fp32 pixels in, 268 MB of bf16 activations out: 56 us at the measured copy
// peak; 12.9 GFLOP of tensor work is nothing).  conv_in_tc.cu runs it as one CTA per 128-pixel tile: at 256x256 that
// is 16 384 CTAs, each re-packing the 48 x 64 weight into its UMMA layout, allocating TMEM and initialising barriers
// for 16 KB of output -- 219 us measured.  Here ONE CTA per SM keeps the B operand, TMEM and barriers for its whole
// life and pipelines tiles through five stages on separate warps:
//   warp 0      TMA: the 3 x (2R+2) input rows a tile touches (box {W, 2R+2, 3, 1}; rows outside the image arrive as
//               zeros = the conv's top / bottom padding), four buffers deep
//   warps 4-11  im2col, two groups of four warps on ALTERNATE tiles: thread = pixel reads its 48 taps from the staged rows
//               (left / right padding by predicate) and writes them as twelve 16-byte pieces of the 128-byte-swizzled
//               K-major A operand (K = 48 fp32); three A tiles in flight
//   warp 1      6 x tcgen05.mma kind::tf32 (M128 N64 K8) per tile, three accumulators in TMEM
//   warps 12-19 epilogue, two groups on alternate tiles: tcgen05.ld -> +bias -> ReLU -> bf16 -> staged as the 128 x 128-byte
//               swizzled tile (two staging tiles per group) -> one thread: TMA store (NHWC pixel rows are contiguous: one box)
// Every stage is one long dependent chain per thread; what set the pace was always the stage whose chain was as long as a
// tile period with a single group of warps on it (149 -> 109 -> 100 -> 75 us).
#include <cuda_bf16.h>

#include <cstring>

#include "common.cuh"
#include "ptx.cuh"

namespace {

constexpr int CIB_THREADS = 640;        // warps: 0 raw-row producer, 1 MMA, 2 TMEM, 4-11 builders (2 groups), 12-19 epilogue (2 groups)
constexpr int COUT = 64;
constexpr uint32_t A_BYTES = 2 * 16384;                 // two K atoms x [128 rows][128 B]
constexpr uint32_t B_ATOM = COUT * 128;
constexpr int NA = 3;                                   // im2col tiles / TMEM accumulators in flight
constexpr int NS = 4;                                   // staging tiles: two per epilogue group
constexpr int NRAW = 4;                                 // staged-input buffers: the TMA loads run three tiles ahead (HBM round trip)

struct CibParams {
    const float *wp, *bias;
    int B, H, W, relu;
    int R, log2_ow;             // output rows per tile (128 / OW), log2(OW)
    int raw_bytes;              // 3 * (2R+2) * W * 4
    long long ntiles;
};

__device__ __forceinline__ void tma_load_4d_nosw(uint32_t dst, const CUtensorMap *m, uint32_t bar, int c0, int c1, int c2, int c3) {
    asm volatile(
        "cp.async.bulk.tensor.4d.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1, {%3, %4, %5, %6}], [%2];" ::
            "r"(dst), "l"(reinterpret_cast<uint64_t>(m)), "r"(bar), "r"(c0), "r"(c1), "r"(c2), "r"(c3) : "memory");
}

__global__ void __launch_bounds__(CIB_THREADS, 1)
conv_in_bf16_kernel(const __grid_constant__ CUtensorMap tma_x, const __grid_constant__ CUtensorMap tma_out, const CibParams p) {
    extern __shared__ unsigned char smem_raw[];
    const uint32_t raw = ptx::smem_u32(smem_raw);
    const uint32_t sbase = (raw + 1023u) & ~1023u;
    unsigned char *sm = smem_raw + (sbase - raw);
    // [A x2][B: 2 atoms][staging 16 KB][raw rows x2][barriers, bias]
    const uint32_t b_off = NA * A_BYTES, st_off = b_off + 2 * B_ATOM, raw_off = st_off + NS * 16384u;
    const uint32_t raw_stride = ((uint32_t)p.raw_bytes + 127u) & ~127u;
    const uint32_t bar_off = raw_off + NRAW * raw_stride;
    const uint32_t bars = sbase + bar_off;
    auto rfull = [&](int s) { return bars + 8u * s; };
    auto rempty = [&](int s) { return bars + 8u * (NRAW + s); };
    auto afull = [&](int s) { return bars + 8u * (2 * NRAW + s); };
    auto aempty = [&](int s) { return bars + 8u * (2 * NRAW + NA + s); };
    auto tfull = [&](int s) { return bars + 8u * (2 * NRAW + 2 * NA + s); };
    auto tempty = [&](int s) { return bars + 8u * (2 * NRAW + 3 * NA + s); };
    auto sfree = [&](int s) { return bars + 8u * (2 * NRAW + 4 * NA + s); };
    volatile uint32_t *tmem_holder = reinterpret_cast<volatile uint32_t *>(sm + bar_off + 8 * (2 * NRAW + 4 * NA + NS));
    float *bias_s = reinterpret_cast<float *>(sm + bar_off + 8 * (2 * NRAW + 4 * NA + NS + 2));

    const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
    const int H = p.H, W = p.W, OW = W / 2, OH = H / 2, R = p.R, NR = 2 * R + 2;
    if (tid == 0) {
        for (int s = 0; s < NRAW; ++s) { ptx::mbar_init(rfull(s), 1); ptx::mbar_init(rempty(s), 4); }
        for (int s = 0; s < NA; ++s) {
            ptx::mbar_init(afull(s), 4); ptx::mbar_init(aempty(s), 1);
            ptx::mbar_init(tfull(s), 1); ptx::mbar_init(tempty(s), 4);
        }
        for (int s = 0; s < NS; ++s) ptx::mbar_init(sfree(s), 1);
        ptx::fence_mbar_init();
        ptx::prefetch_tmap(&tma_x); ptx::prefetch_tmap(&tma_out);
    }
    if (warp == 2) ptx::tmem_alloc(sbase + bar_off + 8 * (2 * NRAW + 4 * NA + NS), 256);
    for (int c = tid; c < COUT; c += CIB_THREADS) bias_s[c] = p.bias ? __ldg(p.bias + c) : 0.f;
    // ---- B operand, once per CTA: wp[k][co] (k = (r*4+s)*3 + c, 48 rows) -> K-major swizzled rows ----
    for (int i = tid; i < COUT * 12; i += CIB_THREADS) {
        const int co = i % COUT, k = (i / COUT) * 4, atom = k >> 5, c16 = (k & 31) >> 2;
        *reinterpret_cast<float4 *>(sm + b_off + atom * B_ATOM + co * 128 + ((c16 ^ (co & 7)) << 4)) =
            make_float4(__ldg(p.wp + (size_t)k * COUT + co), __ldg(p.wp + (size_t)(k + 1) * COUT + co),
                        __ldg(p.wp + (size_t)(k + 2) * COUT + co), __ldg(p.wp + (size_t)(k + 3) * COUT + co));
    }
    // the unused tail of K atom 1 (k = 48..63) is never read: the MMAs cover K = 48 only
    ptx::fence_proxy_async();
    ptx::tc_fence_before();
    __syncthreads();
    ptx::tc_fence_after();
    const uint32_t tmem_base = *tmem_holder;
    pdl_launch_dependents();

    const long long ntiles = p.ntiles;
    const int G = (int)gridDim.x;
    const int tiles_per_img = OH / R;

    if (warp == 0) {
        // ===================== TMA producer: input rows =====================
        if (ptx::elect_one()) {
            pdl_wait();
            uint32_t rs = 0, rpar = 0;
            for (long long tile = blockIdx.x; tile < ntiles; tile += G) {
                const int n = (int)(tile / tiles_per_img), oy0 = (int)(tile % tiles_per_img) * R;
                ptx::mbar_wait_sleep(rempty((int)rs), rpar ^ 1, 100);
                ptx::mbar_expect_tx(rfull((int)rs), (uint32_t)p.raw_bytes);
                tma_load_4d_nosw(sbase + raw_off + rs * raw_stride, &tma_x, rfull((int)rs), 0, 2 * oy0 - 1, 0, n);
                if (++rs == NRAW) { rs = 0; rpar ^= 1; }
            }
        }
    } else if (warp == 1) {
        // ===================== MMA issuer =====================
        const bool leader = ptx::elect_one();
        constexpr uint32_t idesc = ptx::instr_desc(ptx::FMT_TF32, 128, COUT);
        int it = 0;
        for (long long tile = blockIdx.x; tile < ntiles; tile += G, ++it) {
            const int s = it % NA;
            ptx::mbar_wait(afull(s), (uint32_t)((it / NA) & 1));
            ptx::mbar_wait(tempty(s), (uint32_t)(((it / NA) & 1) ^ 1));
            ptx::tc_fence_after();
#pragma unroll
            for (int ks = 0; ks < 6; ++ks)       // K = 48: four k-steps of atom 0, two of atom 1
                if (leader)
                    ptx::mma_tf32(tmem_base + (uint32_t)(s * COUT), ptx::smem_desc_sw128(sbase + s * A_BYTES + (ks >> 2) * 16384 + (ks & 3) * 32),
                                  ptx::smem_desc_sw128(sbase + b_off + (ks >> 2) * B_ATOM + (ks & 3) * 32), idesc, ks > 0 ? 1u : 0u);
            if (leader) { ptx::tc_commit(aempty(s)); ptx::tc_commit(tfull(s)); }
            __syncwarp();
        }
    } else if (warp >= 4 && warp < 12) {
        // ===================== im2col builders: thread = (pixel, kernel-row half) =====================
        const int bt = tid - 128;
        // Two groups of four warps build ALTERNATE tiles (thread = pixel, both kernel-row halves): a tile's build is one long
        // dependent chain per thread (loads -> stores -> proxy fence -> arrive), so with all eight warps on the same tile the
        // stage time was that chain's latency; two tiles in flight halve it.
        const int row = bt & 127, grp = bt >> 7;
        const int r = row >> p.log2_ow, ox = row & (OW - 1);
        int it = grp;
        for (long long tile = (long long)blockIdx.x + (long long)grp * G; tile < ntiles; tile += 2LL * G, it += 2) {
            const int s = it % NA;
            const uint32_t rs = (uint32_t)(it % NRAW), rpar = (uint32_t)((it / NRAW) & 1);
            ptx::mbar_wait_sleep(rfull((int)rs), rpar, 32);
            ptx::mbar_wait_sleep(aempty(s), (uint32_t)(((it / NA) & 1) ^ 1), 32);
            const float *rawp = reinterpret_cast<const float *>(sm + raw_off + rs * raw_stride);
            unsigned char *arow = sm + s * A_BYTES + row * 128;
#pragma unroll
            for (int hf = 0; hf < 2; ++hf) {
            float v[24];                                          // k_local = (trl*4 + s)*3 + c
#pragma unroll
            for (int trl = 0; trl < 2; ++trl)
#pragma unroll
                for (int c = 0; c < 3; ++c) {
                    const float *b = rawp + (c * NR + 2 * r + 2 * hf + trl) * W + 2 * ox;     // taps at ix = 2ox-1 .. 2ox+2
                    const float2 mid = *reinterpret_cast<const float2 *>(b);                 // ix = 2ox, 2ox+1 (8-byte aligned)
                    v[(trl * 4 + 0) * 3 + c] = ox > 0 ? b[-1] : 0.f;
                    v[(trl * 4 + 1) * 3 + c] = mid.x;
                    v[(trl * 4 + 2) * 3 + c] = mid.y;
                    v[(trl * 4 + 3) * 3 + c] = ox < OW - 1 ? b[2] : 0.f;
                }
#pragma unroll
            for (int q = 0; q < 6; ++q) {
                const int kq = hf * 24 + 4 * q, atom = kq >> 5, c16 = (kq & 31) >> 2;
                *reinterpret_cast<float4 *>(arow + atom * 16384 + ((c16 ^ (row & 7)) << 4)) = make_float4(v[4 * q], v[4 * q + 1], v[4 * q + 2], v[4 * q + 3]);
            }
            }
            __syncwarp();
            if (lane == 0) ptx::mbar_arrive(rempty((int)rs));     // the staged rows may be overwritten
            ptx::fence_proxy_async();            // generic-proxy smem writes -> visible to the tensor core
            __syncwarp();
            if (lane == 0) ptx::mbar_arrive(afull(s));
        }
    } else if (warp >= 12) {
        // ===================== epilogue: thread = pixel row; bias, ReLU, bf16, staged tile, one TMA store =====================
        const int q = warp & 3;
        const int row = q * 32 + lane;
        // two groups of four warps take ALTERNATE tiles: an epilogue is one dependent chain per thread (TMEM load -> bias /
        // ReLU / pack -> staging stores -> proxy fence -> barrier -> TMA store) of about a tile period; each group owns two of
        // the four staging tiles, so the groups never wait for each other's stores
        const int eg = (warp - 12) >> 2;
        const bool storer = tid == 384 + eg * 128;
        int it = eg;
        for (long long tile = (long long)blockIdx.x + (long long)eg * G; tile < ntiles; tile += 2LL * G, it += 2) {
            const int s = it % NA, ss = it % NS;
            ptx::mbar_wait_sleep(tfull(s), (uint32_t)((it / NA) & 1), 32);
            ptx::tc_fence_after();
            float va[32], vb[32];
            const uint32_t t0 = tmem_base + ((uint32_t)(q * 32) << 16) + (uint32_t)(s * COUT);
            ptx::tmem_ld32(t0, va);
            ptx::tmem_ld32(t0 + 32, vb);
            ptx::tmem_ld_wait32(va);
            ptx::tmem_ld_wait32(vb);
            ptx::tc_fence_before();
            __syncwarp();
            if (lane == 0) ptx::mbar_arrive(tempty(s));          // the accumulator is in registers
            unsigned char *orow = sm + st_off + ss * 16384 + row * 128;
            ptx::mbar_wait(sfree(ss), (uint32_t)(((it / NS) & 1) ^ 1));      // this group's store of tile it-4 has read the staging tile
#pragma unroll
            for (int hh = 0; hh < 2; ++hh)
#pragma unroll
                for (int i = 0; i < 32; i += 8) {
                    float o[8];
#pragma unroll
                    for (int u = 0; u < 8; ++u) {
                        o[u] = (hh ? vb[i + u] : va[i + u]) + bias_s[hh * 32 + i + u];
                        if (p.relu) o[u] = fmaxf(o[u], 0.f);
                    }
                    const __nv_bfloat162 h0 = __floats2bfloat162_rn(o[0], o[1]), h1 = __floats2bfloat162_rn(o[2], o[3]);
                    const __nv_bfloat162 h2 = __floats2bfloat162_rn(o[4], o[5]), h3 = __floats2bfloat162_rn(o[6], o[7]);
                    *reinterpret_cast<uint4 *>(orow + ((((hh * 32 + i) >> 3) ^ (row & 7)) << 4)) =
                        make_uint4(*reinterpret_cast<const uint32_t *>(&h0), *reinterpret_cast<const uint32_t *>(&h1),
                                   *reinterpret_cast<const uint32_t *>(&h2), *reinterpret_cast<const uint32_t *>(&h3));
                }
            ptx::fence_proxy_async();
            ptx::named_bar_sync(1 + eg, 128);
            if (storer) {
                asm volatile("cp.async.bulk.tensor.2d.global.shared::cta.bulk_group [%0, {%2, %3}], [%1];" ::
                                 "l"(reinterpret_cast<uint64_t>(&tma_out)), "r"(sbase + st_off + (uint32_t)ss * 16384u), "r"(0), "r"((int)(tile * 128)) : "memory");
                asm volatile("cp.async.bulk.commit_group;" ::: "memory");
                // release the staging tile of this group's PREVIOUS store (tile it-2): never wait for the store just issued
                asm volatile("cp.async.bulk.wait_group.read 1;" ::: "memory");
                if (it >= 2) ptx::mbar_arrive(sfree((it - 2) % NS));
            }
        }
        if (storer) asm volatile("cp.async.bulk.wait_group 0;" ::: "memory");
    }
    ptx::tc_fence_before();
    __syncthreads();
    if (warp == 2) ptx::tmem_dealloc(tmem_base, 256);
}

}  // namespace

// Persistent path: whole output rows per 128-pixel tile (OW a power of two in [2, 128], OH % (128/OW) == 0), W <= 256.
bool conv_in_bf16_persistent_ok(int H, int W, const void *x) {
    const int OW = W / 2, OH = H / 2;
    if (H % 2 || W % 4 || OW < 2 || OW > 128 || (OW & (OW - 1)) || W > 256) return false;
    const int R = 128 / OW;
    return R <= OH && OH % R == 0 && 2 * R + 2 <= 256 && (reinterpret_cast<uintptr_t>(x) & 15) == 0;
}

int launch_conv_in_bf16_persistent(const float *x, const float *wp, const float *bias, void *y, int B, int H, int W, int relu,
                                   cudaStream_t s) {
    const int OW = W / 2, OH = H / 2, R = 128 / OW;
    CibParams q;
    memset(&q, 0, sizeof(q));
    q.wp = wp; q.bias = bias; q.B = B; q.H = H; q.W = W; q.relu = relu; q.R = R;
    while ((1 << q.log2_ow) < OW) ++q.log2_ow;
    q.raw_bytes = 3 * (2 * R + 2) * W * 4;
    q.ntiles = (long long)B * (OH / R);
    const long long npix = (long long)B * OH * OW;
    if (npix > 0x7fffffffLL) return VQB_ERR_UNSUPPORTED;
    CUtensorMap tx, tout;
    {
        const uint64_t dims[4] = {(uint64_t)W, (uint64_t)H, 3, (uint64_t)B};
        const uint64_t strides[3] = {(uint64_t)W * 4, (uint64_t)H * W * 4, (uint64_t)3 * H * W * 4};
        const uint32_t box[4] = {(uint32_t)W, (uint32_t)(2 * R + 2), 3u, 1u};
        const uint32_t es[4] = {1u, 1u, 1u, 1u};
        int rc = vqb_encode_tmap_4d(&tx, CU_TENSOR_MAP_DATA_TYPE_FLOAT32, x, dims, strides, box, es, CU_TENSOR_MAP_SWIZZLE_NONE);
        if (rc) return rc;
        rc = vqb_encode_tmap_2d(&tout, CU_TENSOR_MAP_DATA_TYPE_BFLOAT16, y, COUT, (uint64_t)npix, COUT * 2, 64, 128, CU_TENSOR_MAP_SWIZZLE_128B);
        if (rc) return rc;
    }
    const int raw_stride = (q.raw_bytes + 127) & ~127;
    const int smem = NA * (int)A_BYTES + 2 * (int)B_ATOM + NS * 16384 + NRAW * raw_stride + 8 * (2 * NRAW + 4 * NA + NS + 2) + COUT * 4 + 1024;
    if (smem > 227 * 1024) return VQB_ERR_UNSUPPORTED;
    static int attr_max = 0;
    if (smem > attr_max) {
        cudaError_t e = cudaFuncSetAttribute(conv_in_bf16_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, smem);
        if (e != cudaSuccess) return (int)e;
        attr_max = smem;
    }
    int dev = 0, sms = 148;
    cudaGetDevice(&dev);
    cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, dev);
    const int grid = (int)(q.ntiles < sms ? q.ntiles : sms);
    if (cudaError_t le = vqb_launch(conv_in_bf16_kernel, dim3((unsigned)grid), dim3(CIB_THREADS), (size_t)smem, s, tx, tout, q)) return (int)le;
    VQB_COUNT_LAUNCH(1);
    return vqb_cuda_status(cudaGetLastError());
}
