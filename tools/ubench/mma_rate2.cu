// micro-benchmark (round 2): what slows an M128 N128 K16 bf16 tcgen05.mma stream down from its 64-cycle rate?
//   variant 0: canonical operands (SBO = 1024)
//   variant 1: A operand as hconv.cu addresses its halo tiles: SBO = WP * 128 (2304 or 1280), start row shifted per "tap"
//   variant 2: canonical + two other warps reading the OTHER accumulator with tcgen05.ld all the time (an epilogue)
//   variant 3: canonical + eight warps streaming 16-byte stores into other shared memory (stands in for TMA fills)
//   variant 4: canonical + both
// Build: nvcc -gencode arch=compute_100a,code=sm_100a -O3 -o mma_rate2 mma_rate2.cu
#include <cstdio>
#include <cuda_runtime.h>
#include "../../vqvae_b200/csrc/ptx.cuh"

__global__ void __launch_bounds__(384) k(int variant, int sbo, int reps, long long *out) {
    extern __shared__ unsigned char smem_raw[];
    const uint32_t raw = ptx::smem_u32(smem_raw);
    const uint32_t sbase = (raw + 1023u) & ~1023u;
    unsigned char *sm = smem_raw + (sbase - raw);
    __shared__ uint64_t bars[2];
    __shared__ uint32_t holder;
    __shared__ volatile int stop;
    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    for (int i = threadIdx.x; i < 40 * 1024; i += 384) reinterpret_cast<uint32_t *>(sm)[i] = 0x3c003c00u;
    if (threadIdx.x == 0) { ptx::mbar_init(ptx::smem_u32(&bars[0]), 1); ptx::fence_mbar_init(); stop = 0; }
    if (warp == 1) ptx::tmem_alloc(ptx::smem_u32(&holder), 512);
    ptx::fence_proxy_async();
    ptx::tc_fence_before(); __syncthreads(); ptx::tc_fence_after();
    const uint32_t tm = holder;
    if (warp == 0) {
        const bool leader = ptx::elect_one();
        const uint32_t idesc = ptx::instr_desc(ptx::FMT_BF16, 128, 128);
        const uint32_t bar = ptx::smem_u32(&bars[0]);
        const uint32_t a_hi = ptx::desc_hi_sw128((uint32_t)sbo), b_hi = ptx::desc_hi_sw128(1024);
        const uint32_t a0 = sbase >> 4, b0 = (sbase + 96 * 1024) >> 4;     // A region 96 KB, B tiles 4 x 16 KB
        long long t0 = clock64();
        for (int r = 0; r < reps; ++r) {
            // nine "taps": start row offsets (dy * WP + dx) * 128 B like hconv's shifted descriptors (variant 1), else 0
            const uint32_t tap = (uint32_t)(r % 9);
            const uint32_t shift = variant == 1 ? (((tap / 3) * (uint32_t)(sbo / 128) + tap % 3) * 8u) : 0u;
            const uint32_t a_lo = a0 + shift, b_lo = b0 + (uint32_t)(r & 3) * (16384 >> 4);
            const uint32_t acc = tm;
#pragma unroll
            for (int kk = 0; kk < 4; ++kk)
                if (leader)
                    asm volatile(
                        "{\n\t.reg .pred p;\n\t.reg .b64 da, db;\n\tmov.b64 da, {%1, %2};\n\tmov.b64 db, {%3, %4};\n\t"
                        "setp.ne.b32 p, %6, 0;\n\ttcgen05.mma.cta_group::1.kind::f16 [%0], da, db, %5, p;\n\t}" ::
                            "r"(acc), "r"(a_lo + 2u * kk), "r"(a_hi), "r"(b_lo + 2u * kk), "r"(b_hi), "r"(idesc), "r"(1u) : "memory");
        }
        long long t1 = clock64();
        if (leader) ptx::tc_commit(bar);
        __syncwarp();
        ptx::mbar_wait(bar, 0);
        long long t2 = clock64();
        if (leader) { stop = 1; if (blockIdx.x == 0) { out[0] = t1 - t0; out[1] = t2 - t0; } }
    } else if (warp >= 2 && warp < 4 && (variant == 2 || variant == 4)) {
        float v[32], s = 0.f;
        const uint32_t taddr = tm + ((uint32_t)((warp & 3) * 32) << 16) + 256u;      // the other accumulator half
        while (!stop) {
#pragma unroll
            for (int c = 0; c < 128; c += 32) {
                ptx::tmem_ld32(taddr + c, v);
                ptx::tmem_ld_wait32(v);
                s += v[0] + v[31];
            }
        }
        if (s == 12345.f) out[1] = 0;
    } else if (warp >= 4 && (variant == 3 || variant == 4)) {
        uint4 *dst = reinterpret_cast<uint4 *>(sm + 160 * 1024) + (warp - 4) * 256 + lane;
        uint4 val = make_uint4(1, 2, 3, 4);
        while (!stop) {
#pragma unroll
            for (int i = 0; i < 8; ++i) dst[i * 32] = val;
            val.x++;
        }
    }
    ptx::tc_fence_before(); __syncthreads();
    if (warp == 1) ptx::tmem_dealloc(tm, 512);
}

int main() {
    long long *d; cudaMalloc(&d, 16);
    const int smem = 200 * 1024;
    cudaFuncSetAttribute(k, cudaFuncAttributeMaxDynamicSharedMemorySize, smem);
    const int reps = 1152;       // x4 MMAs
    const int cases[][2] = {{0, 1024}, {1, 2304}, {1, 1280}, {1, 2048}, {2, 1024}, {3, 1024}, {4, 1024}};
    for (auto &c : cases)
        for (int grid : {1, 148}) {
            long long h[2];
            for (int it = 0; it < 2; ++it) k<<<grid, 384, smem>>>(c[0], c[1], reps, d);
            cudaError_t e = cudaDeviceSynchronize();
            if (e != cudaSuccess) { printf("err %s\n", cudaGetErrorString(e)); return 1; }
            cudaMemcpy(h, d, 16, cudaMemcpyDeviceToHost);
            printf("variant %d sbo=%4d grid=%3d : issue %.1f  total %.1f cycles per M128 N128 K16 MMA\n", c[0], c[1], grid,
                   (double)h[0] / (reps * 4), (double)h[1] / (reps * 4));
        }
    return 0;
}
