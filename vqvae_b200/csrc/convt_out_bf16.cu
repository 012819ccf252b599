// convt_out_bf16.cu -- the decoder's output layer (decoder.py:34-35): ConvTranspose2d(64 -> 3, k4 s2 p1) on bf16 NHWC
// activations, fp32 NCHW image out, in SCATTER form on tcgen05 (sm_100a).
//
// hconv.cu runs this layer as a gather: every output pixel's 3x3 input neighbourhood, nine shifted K = 64 GEMM steps
// with N = 16 columns -- 36 MMAs per 128 pixels, each re-reading the 4 KB A slice from shared memory for 16 useful
// columns (r02_bf16_layer_times_v1.txt: 172 us at cfg3, shared-memory bound in the tensor pipe's operand fetch).
// A transposed convolution is cheaper the other way round: input pixel (y, x) contributes W[:, co, ky, kx] . in[y, x, :]
// to output (2y - 1 + ky, 2x - 1 + kx), so ONE GEMM  P[pixel][(ky, kx, co)] = in[pixel][:] . W  (K = 64, N = 48 -> 64)
// holds every product of the layer -- 4 MMAs per 128 pixels, the A tile read once -- and an output pixel is the sum
// of four P entries of a 2x2 input neighbourhood.  The epilogue does that sum through shared memory:
//   * tile = 16x16 input pixels of one image (TMA box with a 1-pixel halo, out-of-image pixels zero-filled = the layer's
//     padding), two M = 128 halves, accumulators 2 x 64 TMEM columns, double buffered;
//   * the 64 GEMM columns are ordered by DESTINATION pixel: [0,12) the four taps that land in the pixel's own 2x2
//     output block, then the taps for the pixel above / below / left / right (6 + 2 pad each) and the four diagonal
//     ones (3 + 1 pad each), so that every thread (= input pixel) keeps 12 values in registers, writes 48 to the
//     exchange buffer with 16-byte stores and collects the 36 it needs from its 8 neighbours with 12 16-byte loads
//     (row stride 52 words: conflict free);
//   * interior pixels (14x14) add the bias and store their 2x2x3 outputs as 8-byte pieces of the NCHW image.
// Two epilogue groups of 8 warps take alternate tiles (one TMEM buffer and one exchange buffer each), a TMA producer
// warp and a single-thread MMA issuer run ahead through a 3-stage input ring.  Persistent, one CTA per SM.
#include <cuda_bf16.h>

#include <cstring>

#include "bf16_common.cuh"
#include "common.cuh"
#include "ptx.cuh"

namespace {

constexpr int CO_T = 16;                         // tile edge in input pixels (halo included)
constexpr int CO_IN = CO_T - 2;                  // interior edge
constexpr int CO_STAGE = CO_T * CO_T * 128;      // 32 KB: 256 pixels x 64 bf16
constexpr int CO_NST = 3;
constexpr int CO_WBYTES = 64 * 128;              // 64 GEMM columns x 64 bf16
constexpr int CO_PSTRIDE = 52;                   // words per pixel in the exchange buffer (columns 12..59, 4 pad)
constexpr int CO_PBYTES = 256 * CO_PSTRIDE * 4;
constexpr int CO_NG = 2;                         // epilogue groups
constexpr int CO_THREADS = 128 + CO_NG * 256;

constexpr int CO_OFF_W = CO_NST * CO_STAGE;
constexpr int CO_OFF_P = CO_OFF_W + CO_WBYTES;
constexpr int CO_OFF_BAR = CO_OFF_P + CO_NG * CO_PBYTES;
constexpr int CO_OFF_TMEM = CO_OFF_BAR + 16 * 8;
constexpr int CO_SMEM = CO_OFF_TMEM + 16 + 1024;
static_assert(CO_SMEM <= 227 * 1024, "shared memory budget");

struct CoParams {
    const float *bias;
    float *out;
    int B, H, W, tiles_x, tiles_y;
    long long ntiles;
};

__global__ void __launch_bounds__(CO_THREADS, 1)
convt_out_scatter_kernel(const __grid_constant__ CUtensorMap tin, const __grid_constant__ CUtensorMap tw, const int w_row0,
                         const CoParams p) {
    extern __shared__ unsigned char smem_raw[];
    const uint32_t raw = ptx::smem_u32(smem_raw);
    const uint32_t sbase = (raw + 1023u) & ~1023u;
    unsigned char *sm = smem_raw + (sbase - raw);
    const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
    const uint32_t bars = sbase + CO_OFF_BAR;
    auto full = [&](int s) { return bars + 8u * (uint32_t)s; };
    auto empty = [&](int s) { return bars + 8u * (uint32_t)(CO_NST + s); };
    auto tfull = [&](int b) { return bars + 8u * (uint32_t)(2 * CO_NST + b); };
    auto tempty = [&](int b) { return bars + 8u * (uint32_t)(2 * CO_NST + 2 + b); };
    const uint32_t wfull = bars + 8u * (uint32_t)(2 * CO_NST + 4);
    volatile uint32_t *tmem_holder = reinterpret_cast<volatile uint32_t *>(sm + CO_OFF_TMEM);

    if (tid == 0) {
        ptx::prefetch_tmap(&tin); ptx::prefetch_tmap(&tw);
        for (int s = 0; s < CO_NST; ++s) { ptx::mbar_init(full(s), 1); ptx::mbar_init(empty(s), 1); }
        for (int b = 0; b < 2; ++b) { ptx::mbar_init(tfull(b), 1); ptx::mbar_init(tempty(b), 8); }
        ptx::mbar_init(wfull, 1);
        ptx::fence_mbar_init();
    }
    if (warp == 2) ptx::tmem_alloc(sbase + CO_OFF_TMEM, 256);
    ptx::tc_fence_before();
    __syncthreads();
    ptx::tc_fence_after();
    const uint32_t tmem_base = *tmem_holder;
    pdl_launch_dependents();
    if (warp == 2 && lane == 0) {                      // weights do not depend on the previous layer
        ptx::mbar_expect_tx(wfull, CO_WBYTES);
        ptx::tma_load_2d(sbase + CO_OFF_W, &tw, wfull, 0, w_row0);
    }
    pdl_wait();                                        // the input is written by the previous layer

    const int tiles_per_img = p.tiles_x * p.tiles_y;

    if (warp == 0) {
        // ===================== TMA producer =====================
        if (lane == 0) {
            int it = 0;
            for (long long t = blockIdx.x; t < p.ntiles; t += gridDim.x, ++it) {
                const int s = it % CO_NST;
                const int n = (int)(t / tiles_per_img), rem = (int)(t % tiles_per_img);
                const int ty = rem / p.tiles_x, tx = rem % p.tiles_x;
                ptx::mbar_wait(empty(s), (uint32_t)(((it / CO_NST) & 1) ^ 1));
                ptx::mbar_expect_tx(full(s), CO_STAGE);
                ptx::tma_load_4d(sbase + s * CO_STAGE, &tin, full(s), 0, tx * CO_IN - 1, ty * CO_IN - 1, n);
            }
        }
    } else if (warp == 1) {
        // ===================== MMA issuer (converged warp, elected leader lane issues) =====================
        const bool leader = ptx::elect_one();
        constexpr uint32_t idesc = ptx::instr_desc(ptx::FMT_BF16, 128, 64);
        const uint32_t d_hi = ptx::desc_hi_sw128(1024);
        const uint32_t b_lo = (sbase + CO_OFF_W) >> 4;
        ptx::mbar_wait(wfull, 0);
        int it = 0;
        for (long long t = blockIdx.x; t < p.ntiles; t += gridDim.x, ++it) {
            const int s = it % CO_NST, buf = it & 1;
            ptx::mbar_wait(full(s), (uint32_t)((it / CO_NST) & 1));
            ptx::mbar_wait(tempty(buf), (uint32_t)(((it >> 1) & 1) ^ 1));
            ptx::tc_fence_after();
            const uint32_t a_lo = (sbase + s * CO_STAGE) >> 4;
#pragma unroll
            for (int hf = 0; hf < 2; ++hf)
#pragma unroll
                for (int ks = 0; ks < 4; ++ks)
                    if (leader)
                        mma_bf16_w(tmem_base + (uint32_t)(buf * 128 + hf * 64), a_lo + hf * (CO_STAGE >> 5) + ks * 2, d_hi, b_lo + ks * 2, d_hi,
                                   idesc, ks > 0 ? 1u : 0u);
            if (leader) {
                ptx::tc_commit(tfull(buf));
                ptx::tc_commit(empty(s));
            }
            __syncwarp();
        }
    } else if (warp >= 4) {
        // ===================== epilogue groups: thread = input pixel of the tile =====================
        const int g = (warp - 4) >> 3;                   // group = TMEM buffer = exchange buffer
        const int e = (warp - 4) & 7;
        const int q = e & 3, hf = e >> 2;                // TMEM lane quadrant (= warp % 4), M half
        const int m = hf * 128 + q * 32 + lane;          // pixel of the tile, row-major 16x16
        const int ly = m >> 4, lx = m & 15;
        float *P = reinterpret_cast<float *>(sm + CO_OFF_P + g * CO_PBYTES);
        float *mine = P + m * CO_PSTRIDE;
        const uint32_t taddr = tmem_base + ((uint32_t)(q * 32) << 16) + (uint32_t)(g * 128 + hf * 64);
        const bool interior = ly >= 1 && ly <= CO_IN && lx >= 1 && lx <= CO_IN;
        const float b0 = p.bias ? __ldg(p.bias) : 0.f, b1 = p.bias ? __ldg(p.bias + 1) : 0.f, b2 = p.bias ? __ldg(p.bias + 2) : 0.f;
        const int OH = 2 * p.H, OW = 2 * p.W;
        int it = g;
        for (long long t = (long long)blockIdx.x + (long long)g * gridDim.x; t < p.ntiles; t += 2ll * gridDim.x, it += 2) {
            const int n = (int)(t / tiles_per_img), rem = (int)(t % tiles_per_img);
            const int ty = rem / p.tiles_x, tx = rem % p.tiles_x;
            ptx::mbar_wait_sleep(tfull(g), (uint32_t)((it >> 1) & 1), 32);
            ptx::tc_fence_after();
            float v[64];
            ptx::tmem_ld32(taddr, v);
            ptx::tmem_ld32(taddr + 32, v + 32);
            ptx::tmem_ld_wait32(v);
            ptx::tmem_ld_wait32(v + 32);
            ptx::tc_fence_before();
            __syncwarp();
            if (lane == 0) ptx::mbar_arrive(tempty(g));
            ptx::named_bar_sync(2 + g, 256);             // every thread of the group is done with the previous tile's exchange buffer
#pragma unroll
            for (int i = 0; i < 12; ++i)
                *reinterpret_cast<float4 *>(mine + 4 * i) = make_float4(v[12 + 4 * i], v[13 + 4 * i], v[14 + 4 * i], v[15 + 4 * i]);
            ptx::named_bar_sync(2 + g, 256);
            const int gy = ty * CO_IN + ly - 1, gx = tx * CO_IN + lx - 1;
            if (interior && gy < p.H && gx < p.W) {
                float o[12];                             // (r, s, co): output (2 gy + r, 2 gx + s), channel co
#pragma unroll
                for (int i = 0; i < 12; ++i) o[i] = v[i];
                auto ld4 = [&](int dm, int word) { return *reinterpret_cast<const float4 *>(mine + dm * CO_PSTRIDE + word); };
                {   // pixel above: its taps ky = 3 ("down" group, columns 20..25) land in our row r = 0
                    const float4 a = ld4(-16, 8), b = ld4(-16, 12);
                    o[0] += a.x; o[1] += a.y; o[2] += a.z; o[3] += a.w; o[4] += b.x; o[5] += b.y;
                }
                {   // pixel below: ky = 0 ("up" group, columns 12..17) -> r = 1
                    const float4 a = ld4(16, 0), b = ld4(16, 4);
                    o[6] += a.x; o[7] += a.y; o[8] += a.z; o[9] += a.w; o[10] += b.x; o[11] += b.y;
                }
                {   // pixel to the left: kx = 3 ("right" group, columns 36..41, index r*3+co) -> s = 0
                    const float4 a = ld4(-1, 24), b = ld4(-1, 28);
                    o[0] += a.x; o[1] += a.y; o[2] += a.z; o[6] += a.w; o[7] += b.x; o[8] += b.y;
                }
                {   // pixel to the right: kx = 0 ("left" group, columns 28..33) -> s = 1
                    const float4 a = ld4(1, 16), b = ld4(1, 20);
                    o[3] += a.x; o[4] += a.y; o[5] += a.z; o[9] += a.w; o[10] += b.x; o[11] += b.y;
                }
                {   // diagonals: one tap each
                    const float4 c00 = ld4(-17, 44), c01 = ld4(-15, 40), c10 = ld4(15, 36), c11 = ld4(17, 32);
                    o[0] += c00.x; o[1] += c00.y; o[2] += c00.z;
                    o[3] += c01.x; o[4] += c01.y; o[5] += c01.z;
                    o[6] += c10.x; o[7] += c10.y; o[8] += c10.z;
                    o[9] += c11.x; o[10] += c11.y; o[11] += c11.z;
                }
                float *ob = p.out + ((size_t)n * 3 * OH + 2 * gy) * OW + 2 * gx;
                const size_t cs = (size_t)OH * OW;
#pragma unroll
                for (int r = 0; r < 2; ++r) {
                    *reinterpret_cast<float2 *>(ob + r * OW) = make_float2(o[r * 6 + 0] + b0, o[r * 6 + 3] + b0);
                    *reinterpret_cast<float2 *>(ob + cs + r * OW) = make_float2(o[r * 6 + 1] + b1, o[r * 6 + 4] + b1);
                    *reinterpret_cast<float2 *>(ob + 2 * cs + r * OW) = make_float2(o[r * 6 + 2] + b2, o[r * 6 + 5] + b2);
                }
            }
        }
    }

    ptx::tc_fence_before();
    __syncthreads();
    if (warp == 2) ptx::tmem_dealloc(tmem_base, 256);
}

}  // namespace

// GEMM column n of the scatter form -> (co, ky, kx) of the ConvTranspose2d weight, or co = -1 for a padding column.
void convt_out_scatter_column(int n, int *co, int *ky, int *kx) {
    *co = -1; *ky = 0; *kx = 0;
    if (n < 12) { const int r = n / 6, s = (n / 3) & 1; *co = n % 3; *ky = r + 1; *kx = s + 1; return; }
    const int grp = (n - 12) / 8, j = (n - 12) % 8;
    if (n < 44) {
        if (j >= 6) return;
        const int a = j / 3;                     // s for the up / down groups, r for the left / right groups
        *co = j % 3;
        if (grp == 0) { *ky = 0; *kx = a + 1; }
        else if (grp == 1) { *ky = 3; *kx = a + 1; }
        else if (grp == 2) { *ky = a + 1; *kx = 0; }
        else { *ky = a + 1; *kx = 3; }
        return;
    }
    const int c = (n - 44) / 4, jj = (n - 44) % 4;
    if (c > 3 || jj >= 3) return;
    *co = jj; *ky = (c >> 1) ? 3 : 0; *kx = (c & 1) ? 3 : 0;
}

// (test hook, tests/test_abi_cpu.py: the column layout is host logic and is pinned without a GPU)
extern "C" int vqb_debug_convt_out_scatter_column(int n, int *co, int *ky, int *kx) {
    if (n < 0 || n >= 64 || !co || !ky || !kx) return VQB_ERR_BAD_ARG;
    convt_out_scatter_column(n, co, ky, kx);
    return 0;
}

// in: bf16 NHWC (B, H, W, 64); packed: the layer's packed weights (hconv.cu), GEMM columns of the scatter form at rows
// [w_row0, w_row0 + 64); out: fp32 NCHW (B, 3, 2H, 2W).
int launch_convt_out_scatter(const void *in, const void *packed, int packed_rows, int w_row0, const float *bias, float *out, int B, int H,
                             int W, cudaStream_t s) {
    CUtensorMap tin, tw;
    typedef unsigned long long u64;
    const u64 dims[4] = {64ull, (u64)W, (u64)H, (u64)B};
    const u64 strides[3] = {64ull * 2, (u64)W * 64 * 2, (u64)H * W * 64 * 2};
    const uint32_t box[4] = {64u, (uint32_t)CO_T, (uint32_t)CO_T, 1u};
    int rc = vqb_encode_tmap_nd(&tin, CU_TENSOR_MAP_DATA_TYPE_BFLOAT16, in, 4, dims, strides, box, CU_TENSOR_MAP_SWIZZLE_128B);
    if (rc) return rc;
    rc = vqb_encode_tmap_2d(&tw, CU_TENSOR_MAP_DATA_TYPE_BFLOAT16, packed, 64, (uint64_t)packed_rows, 128, 64, 64, CU_TENSOR_MAP_SWIZZLE_128B);
    if (rc) return rc;
    static bool attr_set = false;
    if (!attr_set) {
        cudaError_t e = cudaFuncSetAttribute(convt_out_scatter_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, CO_SMEM);
        if (e != cudaSuccess) return (int)e;
        attr_set = true;
    }
    CoParams p;
    memset(&p, 0, sizeof(p));
    p.bias = bias; p.out = out; p.B = B; p.H = H; p.W = W;
    p.tiles_x = (W + CO_IN - 1) / CO_IN; p.tiles_y = (H + CO_IN - 1) / CO_IN;
    p.ntiles = (long long)p.tiles_x * p.tiles_y * B;
    int dev = 0, sms = 148;
    cudaGetDevice(&dev);
    cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, dev);
    const int grid = (int)(p.ntiles < sms ? p.ntiles : sms);
    if (cudaError_t le = vqb_launch(convt_out_scatter_kernel, dim3((unsigned)grid), dim3(CO_THREADS), (size_t)CO_SMEM, s, tin, tw, w_row0, p))
        return (int)le;
    VQB_COUNT_LAUNCH(1);
    return vqb_cuda_status(cudaGetLastError());
}
