// conv_in_tc.cu -- encoder.py:29-31, Conv2d(3 -> Cout, k4 s2 p1) + ReLU on tcgen05 (TF32 mode).
//
// The NCHW module input cannot be TMA'd into K-major rows (a pixel's 48 taps are scattered over
// three planes), so this one layer builds its im2col tile by hand: 256 threads gather the
// 128 pixels x 48 taps (zero-padded to K = 64) of the tile straight into shared memory in the
// 128-byte-swizzled K-major layout UMMA reads, the packed weight goes in next to it, one thread
// issues 8 tcgen05.mma (M128, N = Cout, K8), and all 8 warps run the epilogue
// (tcgen05.ld -> +bias -> ReLU -> 16-byte NHWC stores).  One 128-pixel tile per CTA, ~50 KB of
// shared memory, several CTAs per SM hide each other's gather latency.
#include "common.cuh"
#include "ptx.cuh"

namespace {

constexpr int CI_THREADS = 256;

template <int COUT>
__global__ void __launch_bounds__(CI_THREADS)
conv_in_tc_kernel(const float *__restrict__ x, const float *__restrict__ wp, const float *__restrict__ bias,
                  float *__restrict__ y, int B, int H, int W, int relu) {
    constexpr int Cout = COUT;
    extern __shared__ unsigned char smem_raw[];
    const uint32_t raw = ptx::smem_u32(smem_raw);
    const uint32_t sbase = (raw + 1023u) & ~1023u;
    unsigned char *sm = smem_raw + (sbase - raw);
    // A: 2 atoms x [128 rows][128 B]; B: 2 atoms x [Cout rows][128 B]
    const int b_atom = Cout * 128;
    const uint32_t b_off = 2 * 16384;
    const uint32_t misc_off = b_off + 2 * (uint32_t)b_atom;
    const uint32_t bar = sbase + misc_off;
    volatile uint32_t *tmem_holder = reinterpret_cast<volatile uint32_t *>(sm + misc_off + 8);
    float *bias_s = reinterpret_cast<float *>(sm + misc_off + 16);

    const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
    const int OH = H / 2, OW = W / 2;
    const long long npix = (long long)B * OH * OW;
    const long long pix0 = (long long)blockIdx.x * 128;
    int tcols = 32;
    while (tcols < Cout) tcols <<= 1;

    if (tid == 0) {
        ptx::mbar_init(bar, 1);
        ptx::fence_mbar_init();
    }
    if (warp == 1) ptx::tmem_alloc(sbase + misc_off + 8, (uint32_t)tcols);
    for (int c = tid; c < Cout; c += CI_THREADS) bias_s[c] = bias ? __ldg(bias + c) : 0.f;

    // ---- B operand: wp[k][co] (k = (r*4+s)*3 + c, 48 rows) -> K-major swizzled rows of 64 (zero padded) ----
    {
        // all weight loads of the thread in flight before the first store (a rolled loop serialises
        // Cout*64/256 L2 round trips: it was half of this kernel's 21 us)
        constexpr int NB = COUT * 64 / CI_THREADS;
        float wv[NB];
#pragma unroll
        for (int u = 0; u < NB; ++u) {
            const int i = tid + u * CI_THREADS;
            const int co = i % COUT, k = i / COUT;             // co fastest: coalesced reads of wp
            wv[u] = k < 48 ? __ldg(wp + (size_t)k * COUT + co) : 0.f;
        }
#pragma unroll
        for (int u = 0; u < NB; ++u) {
            const int i = tid + u * CI_THREADS;
            const int co = i % COUT, k = i / COUT;
            const int atom = k >> 5, kk = k & 31;
            *reinterpret_cast<float *>(sm + b_off + atom * b_atom + co * 128 + (((kk >> 2) ^ (co & 7)) << 4) + (kk & 3) * 4) = wv[u];
        }
    }
    pdl_launch_dependents();
    pdl_wait();                    // x may be written by the previous kernel / copy
    // ---- A operand: im2col of the tile, thread = (pixel row, K atom) ----
    {
        const int row = tid & 127, atom = tid >> 7;
        const long long pix = pix0 + row;
        const bool live = pix < npix;
        const long long pp = live ? pix : 0;
        const int ox = (int)(pp % OW);
        const long long t = pp / OW;
        const int oy = (int)(t % OH);
        const int n = (int)(t / OH);
        unsigned char *arow = sm + atom * 16384 + row * 128;
        float vals[32];                      // all 32 gathers in flight before the first store
#pragma unroll
        for (int kk = 0; kk < 32; ++kk) {
            const int k = atom * 32 + kk;
            float v = 0.f;
            if (live && k < 48) {
                const int tap = k / 3, c = k - tap * 3;
                const int iy = 2 * oy - 1 + (tap >> 2), ix = 2 * ox - 1 + (tap & 3);
                if (iy >= 0 && iy < H && ix >= 0 && ix < W) v = __ldg(x + (((long long)n * 3 + c) * H + iy) * W + ix);
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
        for (int ks = 0; ks < 8; ++ks)
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
            if (pix < npix) {
                float4 *dst = reinterpret_cast<float4 *>(y + pix * Cout + c0);
#pragma unroll
                for (int i = 0; i < 32; i += 4) {
                    const float4 bb = *reinterpret_cast<const float4 *>(bias_s + c0 + i);
                    float4 o = make_float4(v[i] + bb.x, v[i + 1] + bb.y, v[i + 2] + bb.z, v[i + 3] + bb.w);
                    if (relu) { o.x = fmaxf(o.x, 0.f); o.y = fmaxf(o.y, 0.f); o.z = fmaxf(o.z, 0.f); o.w = fmaxf(o.w, 0.f); }
                    dst[i >> 2] = o;
                }
            }
        }
    }
    ptx::tc_fence_before();
    __syncthreads();
    if (warp == 1) ptx::tmem_dealloc(tmem_base, (uint32_t)tcols);
}

}  // namespace

bool conv_in_tc_supported(int Cin, int Cout, int H, int W, const void *y) {
    return Cin == 3 && (Cout == 64 || Cout == 128) && H % 2 == 0 && W % 2 == 0 &&
           (reinterpret_cast<uintptr_t>(y) & 15) == 0;
}

// wp = FFMA packing [(r*4+s)*3 + c][co] (first region of vqb_pack_conv_weight_f32)
int launch_conv_in_tc(const float *x, const float *wp, const float *bias, float *y, int B, int H, int W, int Cout,
                      int relu, cudaStream_t s) {
    const long long npix = (long long)B * (H / 2) * (W / 2);
    const long long blocks = (npix + 127) / 128;
    if (blocks <= 0 || blocks > 0x7fffffffLL) return VQB_ERR_UNSUPPORTED;
    const int smem = 2 * 16384 + 2 * Cout * 128 + 16 + Cout * 4 + 1024;
    static bool attr_set = false;
    if (!attr_set) {
        cudaError_t e = cudaFuncSetAttribute(conv_in_tc_kernel<64>, cudaFuncAttributeMaxDynamicSharedMemorySize,
                                             2 * 16384 + 2 * 64 * 128 + 16 + 64 * 4 + 1024);
        if (e != cudaSuccess) return (int)e;
        e = cudaFuncSetAttribute(conv_in_tc_kernel<128>, cudaFuncAttributeMaxDynamicSharedMemorySize,
                                 2 * 16384 + 2 * 128 * 128 + 16 + 128 * 4 + 1024);
        if (e != cudaSuccess) return (int)e;
        attr_set = true;
    }
    cudaError_t le;
    if (Cout == 64) le = vqb_launch(conv_in_tc_kernel<64>, dim3((unsigned)blocks), dim3(CI_THREADS), (size_t)smem, s, x, wp, bias, y, B, H, W, relu);
    else le = vqb_launch(conv_in_tc_kernel<128>, dim3((unsigned)blocks), dim3(CI_THREADS), (size_t)smem, s, x, wp, bias, y, B, H, W, relu);
    if (le != cudaSuccess) return (int)le;
    VQB_COUNT_LAUNCH(1);
    return vqb_cuda_status(cudaGetLastError());
}
