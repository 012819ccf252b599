// conv_in_tc.cu -- encoder.py:29-31, Conv2d(3 -> Cout, k4 s2 p1) + ReLU on tcgen05 (TF32 mode).
//
// The NCHW module input cannot be TMA'd into K-major rows (a pixel's 48 taps are scattered over
// three planes), so this one layer builds its im2col tile by hand.  One 128-pixel tile per CTA:
//   1. the input rows the tile touches (3 planes x (2R+2) rows, R = 128/OW output rows) are copied
//      into shared memory with coalesced 16-byte loads, zero padded left/right/top/bottom;
//   2. thread (pixel, kernel-row half) reads its 24 taps from there (12 aligned 8-byte loads, constant
//      offsets) and writes them as six 16-byte pieces of the 128-byte-swizzled K-major A operand;
//   3. the packed weight (48 x Cout) goes next to it the same way (B operand);
//   4. one thread issues 6 tcgen05.mma (M128, N = Cout, K8: K = 48, no padding k-steps);
//   5. the epilogue (tcgen05.ld -> +bias -> ReLU) stages the 128 x Cout tile in shared memory over the dead
//      A operand and ONE thread TMA-stores it (NHWC rows are contiguous): no scattered 16-byte stores.
// The first version gathered straight from global memory with per-tap index arithmetic: 1480
// instructions per thread, issue-bound at 20 us for cfg2 (profiles/r01_step_kernels_ncu.txt).
// Shapes outside the fast path (OW not dividing 128, tiles straddling images) keep that gather.
#include <cuda_bf16.h>

#include "common.cuh"
#include "ptx.cuh"

namespace {

constexpr int CI_THREADS = 256;

struct ConvInParams {
    const float *x, *wp, *bias;
    float *y;
    int B, H, W, relu;
    int R;          // fast path: output rows per tile (128 / OW), 0 = generic gather
    int log2_ow;    // fast path
    int raw_floats; // fast path: 3 * (2R+2) * (W+2)
    int tma_store;  // epilogue through shared memory + TMA (Cout == 64)
    int out_bf16;   // VQB_BF16 mode: the NHWC output is bf16 (the 128 x 64 tile is one 16 KB swizzled atom); needs tma_store
};

template <int COUT>
__global__ void __launch_bounds__(CI_THREADS)
conv_in_tc_kernel(const __grid_constant__ CUtensorMap tma_out, const ConvInParams p) {
    constexpr int Cout = COUT;
    extern __shared__ unsigned char smem_raw[];
    const uint32_t raw = ptx::smem_u32(smem_raw);
    const uint32_t sbase = (raw + 1023u) & ~1023u;
    unsigned char *sm = smem_raw + (sbase - raw);
    // A: 2 atoms x [128 rows][128 B]; B: 2 atoms x [Cout rows][128 B]; misc; raw input rows (fast path)
    constexpr int b_atom = Cout * 128;
    constexpr uint32_t b_off = 2 * 16384;
    constexpr uint32_t misc_off = b_off + 2 * (uint32_t)b_atom;
    constexpr uint32_t raw_off = misc_off + 16 + Cout * 4;
    const uint32_t bar = sbase + misc_off;
    volatile uint32_t *tmem_holder = reinterpret_cast<volatile uint32_t *>(sm + misc_off + 8);
    float *bias_s = reinterpret_cast<float *>(sm + misc_off + 16);
    float *rawp = reinterpret_cast<float *>(sm + raw_off);

    const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
    const int H = p.H, W = p.W;
    const int OH = H / 2, OW = W / 2;
    const long long npix = (long long)p.B * OH * OW;
    const long long pix0 = (long long)blockIdx.x * 128;
    constexpr int tcols = COUT <= 64 ? 64 : 128;

    if (tid == 0) {
        ptx::prefetch_tmap(&tma_out);
        ptx::mbar_init(bar, 1);
        ptx::fence_mbar_init();
    }
    if (warp == 1) ptx::tmem_alloc(sbase + misc_off + 8, (uint32_t)tcols);
    for (int c = tid; c < Cout; c += CI_THREADS) bias_s[c] = p.bias ? __ldg(p.bias + c) : 0.f;

    // ---- B operand: wp[k][co] (k = (r*4+s)*3 + c, 48 rows) -> K-major swizzled rows; thread = (co, k group) ----
    {
        constexpr int KPT = 48 * COUT / CI_THREADS;          // 12 (Cout 64) or 24 (Cout 128): multiples of 4
        const int co = tid % COUT, k0 = (tid / COUT) * KPT;
        float wv[KPT];
#pragma unroll
        for (int u = 0; u < KPT; ++u) wv[u] = __ldg(p.wp + (size_t)(k0 + u) * COUT + co);   // coalesced over co
#pragma unroll
        for (int j = 0; j < KPT / 4; ++j) {
            const int k = k0 + 4 * j, atom = k >> 5, c16 = (k & 31) >> 2;
            *reinterpret_cast<float4 *>(sm + b_off + atom * b_atom + co * 128 + ((c16 ^ (co & 7)) << 4)) =
                make_float4(wv[4 * j], wv[4 * j + 1], wv[4 * j + 2], wv[4 * j + 3]);
        }
    }
    pdl_launch_dependents();
    pdl_wait();                    // x may be written by the previous kernel / copy
    if (p.R > 0) {
        // ---- fast path: stage the input rows, then im2col from shared memory ----
        const int R = p.R, NR = 2 * R + 2, RS = W + 2;       // row j <-> iy = 2*oy0 - 1 + j; col <-> ix + 1
        const int tiles_per_img = OH / R;
        const int n = (int)(blockIdx.x / tiles_per_img), oy0 = (int)(blockIdx.x % tiles_per_img) * R;
        const int w4 = W >> 2;
        const int total4 = 3 * NR * w4;
        const float *xn = p.x + (size_t)n * 3 * H * W;
        for (int i = tid; i < total4; i += CI_THREADS) {
            const int j4 = i % w4, rowid = i / w4;
            const int c = rowid / NR, j = rowid - c * NR;
            const int iy = 2 * oy0 - 1 + j;
            float4 v = make_float4(0.f, 0.f, 0.f, 0.f);
            if (iy >= 0 && iy < H) v = __ldg(reinterpret_cast<const float4 *>(xn + ((size_t)c * H + iy) * W) + j4);
            float *d = rawp + rowid * RS + 1 + 4 * j4;
            d[0] = v.x; d[1] = v.y; d[2] = v.z; d[3] = v.w;
        }
        for (int i = tid; i < 3 * NR * 2; i += CI_THREADS)    // left / right zero padding
            rawp[(i >> 1) * RS + ((i & 1) ? W + 1 : 0)] = 0.f;
        __syncthreads();
        const int row = tid & 127, hf = tid >> 7;             // pixel of the tile, kernel rows {2hf, 2hf+1}
        const int r = row >> p.log2_ow, ox = row & (OW - 1);
        float v[24];                                          // k_local = (trl*4 + s)*3 + c
#pragma unroll
        for (int trl = 0; trl < 2; ++trl)
#pragma unroll
            for (int c = 0; c < 3; ++c) {
                const float *b = rawp + (c * NR + 2 * r + 2 * hf + trl) * RS + 2 * ox;    // 8-byte aligned (RS even)
                const float2 p0 = *reinterpret_cast<const float2 *>(b), p1 = *reinterpret_cast<const float2 *>(b + 2);
                v[(trl * 4 + 0) * 3 + c] = p0.x; v[(trl * 4 + 1) * 3 + c] = p0.y;
                v[(trl * 4 + 2) * 3 + c] = p1.x; v[(trl * 4 + 3) * 3 + c] = p1.y;
            }
        unsigned char *arow = sm + row * 128;
#pragma unroll
        for (int q = 0; q < 6; ++q) {
            const int kq = hf * 24 + 4 * q, atom = kq >> 5, c16 = (kq & 31) >> 2;
            *reinterpret_cast<float4 *>(arow + atom * 16384 + ((c16 ^ (row & 7)) << 4)) =
                make_float4(v[4 * q], v[4 * q + 1], v[4 * q + 2], v[4 * q + 3]);
        }
    } else {
        // ---- generic gather: thread = (pixel row, K atom), straight from global memory ----
        const int row = tid & 127, atom = tid >> 7;
        const long long pix = pix0 + row;
        const bool live = pix < npix;
        const long long pp = live ? pix : 0;
        const int ox = (int)(pp % OW);
        const long long t = pp / OW;
        const int oy = (int)(t % OH);
        const int n = (int)(t / OH);
        unsigned char *arow = sm + atom * 16384 + row * 128;
        float vals[32];                      // all gathers in flight before the first store
#pragma unroll
        for (int kk = 0; kk < 32; ++kk) {
            const int k = atom * 32 + kk;
            float v = 0.f;
            if (live && k < 48) {
                const int tap = k / 3, c = k - tap * 3;
                const int iy = 2 * oy - 1 + (tap >> 2), ix = 2 * ox - 1 + (tap & 3);
                if (iy >= 0 && iy < H && ix >= 0 && ix < W) v = __ldg(p.x + (((long long)n * 3 + c) * H + iy) * W + ix);
            }
            vals[kk] = v;
        }
#pragma unroll
        for (int c16 = 0; c16 < 8; ++c16)
            *reinterpret_cast<float4 *>(arow + ((c16 ^ (row & 7)) << 4)) =
                make_float4(vals[c16 * 4], vals[c16 * 4 + 1], vals[c16 * 4 + 2], vals[c16 * 4 + 3]);
    }
    ptx::fence_proxy_async();            // generic-proxy smem writes -> visible to the tensor core
    ptx::tc_fence_before();
    __syncthreads();
    ptx::tc_fence_after();
    const uint32_t tmem_base = *tmem_holder;

    if (tid == 0) {
        const uint32_t idesc = ptx::instr_desc(ptx::FMT_TF32, 128, (uint32_t)Cout);
#pragma unroll
        for (int ks = 0; ks < 6; ++ks)       // K = 48: four k-steps of atom 0, two of atom 1
            ptx::mma_tf32(tmem_base, ptx::smem_desc_sw128(sbase + (ks >> 2) * 16384 + (ks & 3) * 32),
                          ptx::smem_desc_sw128(sbase + b_off + (ks >> 2) * b_atom + (ks & 3) * 32), idesc, ks > 0 ? 1u : 0u);
        ptx::tc_commit(bar);
    }
    ptx::mbar_wait(bar, 0);
    ptx::tc_fence_after();

    // ---- epilogue: warp w reads TMEM lanes 32*(w%4).., column half w/4 ----
    {
        const int q = warp & 3, half = warp >> 2;
        const int row = q * 32 + lane;
        const long long pix = pix0 + row;
        const int cbeg = half * (Cout / 2), cend = cbeg + Cout / 2;      // Cout % 64 == 0: halves are 32-multiples
        for (int c0 = cbeg; c0 < cend; c0 += 32) {
            float v[32];
            ptx::tmem_ld32(tmem_base + ((uint32_t)(q * 32) << 16) + (uint32_t)c0, v);
            ptx::tmem_ld_wait32(v);
            if (p.tma_store && p.out_bf16) {
                // bf16 output: 64 channels = one 128-byte row; this thread's 32 columns are four 16-byte pieces
                unsigned char *orow = sm + row * 128;
#pragma unroll
                for (int i = 0; i < 32; i += 8) {
                    float o[8];
#pragma unroll
                    for (int u = 0; u < 8; ++u) {
                        o[u] = v[i + u] + bias_s[c0 + i + u];
                        if (p.relu) o[u] = fmaxf(o[u], 0.f);
                    }
                    const __nv_bfloat162 h0 = __floats2bfloat162_rn(o[0], o[1]), h1 = __floats2bfloat162_rn(o[2], o[3]);
                    const __nv_bfloat162 h2 = __floats2bfloat162_rn(o[4], o[5]), h3 = __floats2bfloat162_rn(o[6], o[7]);
                    *reinterpret_cast<uint4 *>(orow + ((((c0 + i) >> 3) ^ (row & 7)) << 4)) =
                        make_uint4(*reinterpret_cast<const uint32_t *>(&h0), *reinterpret_cast<const uint32_t *>(&h1),
                                   *reinterpret_cast<const uint32_t *>(&h2), *reinterpret_cast<const uint32_t *>(&h3));
                }
            } else if (p.tma_store) {
                // the A operand is dead (all MMAs completed): 128 x Cout tile, one 16 KB swizzled atom per 32 channels
                unsigned char *orow = sm + (c0 >> 5) * 16384 + row * 128;
#pragma unroll
                for (int i = 0; i < 32; i += 4) {
                    const float4 bb = *reinterpret_cast<const float4 *>(bias_s + c0 + i);
                    float4 o = make_float4(v[i] + bb.x, v[i + 1] + bb.y, v[i + 2] + bb.z, v[i + 3] + bb.w);
                    if (p.relu) { o.x = fmaxf(o.x, 0.f); o.y = fmaxf(o.y, 0.f); o.z = fmaxf(o.z, 0.f); o.w = fmaxf(o.w, 0.f); }
                    *reinterpret_cast<float4 *>(orow + (((i >> 2) ^ (row & 7)) << 4)) = o;
                }
            } else if (pix < npix) {
                float4 *dst = reinterpret_cast<float4 *>(p.y + pix * Cout + c0);
#pragma unroll
                for (int i = 0; i < 32; i += 4) {
                    const float4 bb = *reinterpret_cast<const float4 *>(bias_s + c0 + i);
                    float4 o = make_float4(v[i] + bb.x, v[i + 1] + bb.y, v[i + 2] + bb.z, v[i + 3] + bb.w);
                    if (p.relu) { o.x = fmaxf(o.x, 0.f); o.y = fmaxf(o.y, 0.f); o.z = fmaxf(o.z, 0.f); o.w = fmaxf(o.w, 0.f); }
                    dst[i >> 2] = o;
                }
            }
        }
    }
    if (p.tma_store) ptx::fence_proxy_async();       // staged tile -> visible to the TMA store
    ptx::tc_fence_before();
    __syncthreads();
    if (p.tma_store && tid == 0) {
        // box {32 channels, 128 pixels} ({64, 128} for bf16); rows past the last pixel are clipped by the tensor map
        const int natoms = p.out_bf16 ? 1 : Cout / 32;
        for (int a = 0; a < natoms; ++a)
            asm volatile("cp.async.bulk.tensor.2d.global.shared::cta.bulk_group [%0, {%2, %3}], [%1];" ::
                             "l"(reinterpret_cast<uint64_t>(&tma_out)), "r"(sbase + a * 16384), "r"(a * 32),
                             "r"((int)pix0) : "memory");
        asm volatile("cp.async.bulk.commit_group;" ::: "memory");
        asm volatile("cp.async.bulk.wait_group.read 0;" ::: "memory");
    }
    if (warp == 1) ptx::tmem_dealloc(tmem_base, (uint32_t)tcols);
}

}  // namespace

bool conv_in_tc_supported(int Cin, int Cout, int H, int W, const void *y) {
    return Cin == 3 && (Cout == 64 || Cout == 128) && H % 2 == 0 && W % 2 == 0 &&
           (reinterpret_cast<uintptr_t>(y) & 15) == 0;
}

// wp = FFMA packing [(r*4+s)*3 + c][co] (first region of vqb_pack_conv_weight_f32)
int launch_conv_in_tc_ex(const float *x, const float *wp, const float *bias, void *y, int B, int H, int W, int Cout,
                         int relu, int out_bf16, cudaStream_t s);
int launch_conv_in_tc(const float *x, const float *wp, const float *bias, float *y, int B, int H, int W, int Cout,
                      int relu, cudaStream_t s) {
    return launch_conv_in_tc_ex(x, wp, bias, y, B, H, W, Cout, relu, 0, s);
}

int launch_conv_in_tc_ex(const float *x, const float *wp, const float *bias, void *y, int B, int H, int W, int Cout,
                         int relu, int out_bf16, cudaStream_t s) {
    const int OH = H / 2, OW = W / 2;
    const long long npix = (long long)B * OH * OW;
    const long long blocks = (npix + 127) / 128;
    if (blocks <= 0 || blocks > 0x7fffffffLL || npix > 0x7fffffffLL) return VQB_ERR_UNSUPPORTED;
    ConvInParams q;
    if (out_bf16 && Cout != 64) return VQB_ERR_UNSUPPORTED;
    q.x = x; q.wp = wp; q.bias = bias; q.y = reinterpret_cast<float *>(y); q.B = B; q.H = H; q.W = W; q.relu = relu;
    q.out_bf16 = out_bf16;
    q.R = 0; q.log2_ow = 0; q.raw_floats = 0;
    // fast path: whole output rows per tile, tiles inside one image, 16-byte aligned input rows
    if (OW >= 2 && OW <= 128 && (OW & (OW - 1)) == 0 && (reinterpret_cast<uintptr_t>(x) & 15) == 0) {
        const int R = 128 / OW;
        if (R <= OH && OH % R == 0) {
            q.R = R;
            while ((1 << q.log2_ow) < OW) ++q.log2_ow;
            q.raw_floats = 3 * (2 * R + 2) * (W + 2);
        }
    }
    q.tma_store = Cout == 64 ? 1 : 0;        // (Cout 128 would need 64 KB of staging: direct stores)
    CUtensorMap tout;
    int rc = out_bf16
        ? vqb_encode_tmap_2d(&tout, CU_TENSOR_MAP_DATA_TYPE_BFLOAT16, y, (uint64_t)Cout, (uint64_t)npix,
                             (uint64_t)Cout * 2, 64, 128, CU_TENSOR_MAP_SWIZZLE_128B)
        : vqb_encode_tmap_2d(&tout, CU_TENSOR_MAP_DATA_TYPE_FLOAT32, y, (uint64_t)Cout, (uint64_t)npix,
                             (uint64_t)Cout * 4, 32, 128, CU_TENSOR_MAP_SWIZZLE_128B);
    if (rc) return rc;
    const int smem = 2 * 16384 + 2 * Cout * 128 + 16 + Cout * 4 + q.raw_floats * 4 + 16 + 1024;
    if (smem > 200 * 1024) return VQB_ERR_UNSUPPORTED;
    static int attr_max[2] = {0, 0};
    const int ti = Cout == 64 ? 0 : 1;
    if (smem > attr_max[ti]) {
        cudaError_t e = Cout == 64
            ? cudaFuncSetAttribute(conv_in_tc_kernel<64>, cudaFuncAttributeMaxDynamicSharedMemorySize, smem)
            : cudaFuncSetAttribute(conv_in_tc_kernel<128>, cudaFuncAttributeMaxDynamicSharedMemorySize, smem);
        if (e != cudaSuccess) return (int)e;
        attr_max[ti] = smem;
    }
    cudaError_t le;
    if (Cout == 64) le = vqb_launch(conv_in_tc_kernel<64>, dim3((unsigned)blocks), dim3(CI_THREADS), (size_t)smem, s, tout, q);
    else le = vqb_launch(conv_in_tc_kernel<128>, dim3((unsigned)blocks), dim3(CI_THREADS), (size_t)smem, s, tout, q);
    if (le != cudaSuccess) return (int)le;
    VQB_COUNT_LAUNCH(1);
    return vqb_cuda_status(cudaGetLastError());
}
