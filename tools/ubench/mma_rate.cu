// micro-benchmark (round 2): sustained tcgen05.mma rate per SM, issued from a CONVERGED warp with an elected
// leader lane (the way the product kernels issue), operands in shared memory (SS), M = 128.
//   kind f16 (bf16 operands, K = 16) and tf32 (K = 8), N in {32, 64, 128, 256}; 1 CTA and 148 CTAs.
// Build: nvcc -gencode arch=compute_100a,code=sm_100a -O3 -o mma_rate mma_rate.cu
#include <cstdio>
#include <cuda_runtime.h>
#include "../../vqvae_b200/csrc/ptx.cuh"

template <int KIND>   // 0 = bf16 (kind::f16), 1 = tf32
__global__ void __launch_bounds__(128) k(int N, int reps, int nbuf, long long *out) {
    extern __shared__ unsigned char smem_raw[];
    const uint32_t raw = ptx::smem_u32(smem_raw);
    const uint32_t sbase = (raw + 1023u) & ~1023u;
    __shared__ uint64_t bars[2];
    __shared__ uint32_t holder;
    const int warp = threadIdx.x >> 5;
    for (int i = threadIdx.x; i < 48 * 1024; i += 128) reinterpret_cast<uint32_t *>(smem_raw + (sbase - raw))[i] = 0x3c003c00u;
    if (threadIdx.x == 0) { ptx::mbar_init(ptx::smem_u32(&bars[0]), 1); ptx::fence_mbar_init(); }
    if (warp == 1) ptx::tmem_alloc(ptx::smem_u32(&holder), 512);
    ptx::fence_proxy_async();
    ptx::tc_fence_before(); __syncthreads(); ptx::tc_fence_after();
    const uint32_t tm = holder;
    if (warp == 0) {
        const bool leader = ptx::elect_one();
        const uint32_t idesc = ptx::instr_desc(KIND == 0 ? ptx::FMT_BF16 : ptx::FMT_TF32, 128, (uint32_t)N);
        const uint32_t bar = ptx::smem_u32(&bars[0]);
        const uint32_t hi = ptx::desc_hi_sw128(1024);
        const uint32_t a0 = sbase >> 4, b0 = (sbase + 64 * 1024) >> 4;     // A tiles: 4 x 16 KB; B tiles: 4 x 32 KB
        long long t0 = clock64();
        uint32_t buf = 0;
        for (int r = 0; r < reps; ++r) {
            const uint32_t a_lo = a0 + buf * (16384 >> 4), b_lo = b0 + buf * (32768 >> 4);
            const uint32_t acc = tm + (uint32_t)((r & 1) * 256);
#pragma unroll
            for (int kk = 0; kk < 4; ++kk) {
                if (leader) {
                    if (KIND == 0) {
                        asm volatile(
                            "{\n\t.reg .pred p;\n\t.reg .b64 da, db;\n\tmov.b64 da, {%1, %2};\n\tmov.b64 db, {%3, %4};\n\t"
                            "setp.ne.b32 p, %6, 0;\n\ttcgen05.mma.cta_group::1.kind::f16 [%0], da, db, %5, p;\n\t}" ::
                                "r"(acc), "r"(a_lo + 2u * kk), "r"(hi), "r"(b_lo + 2u * kk), "r"(hi), "r"(idesc), "r"(1u) : "memory");
                    } else {
                        ptx::mma_tf32_w(acc, a_lo + 2u * kk, hi, b_lo + 2u * kk, hi, idesc, 1u);
                    }
                }
            }
            if (++buf == (uint32_t)nbuf) buf = 0;
        }
        long long t1 = clock64();
        if (leader) ptx::tc_commit(bar);
        __syncwarp();
        ptx::mbar_wait(bar, 0);
        long long t2 = clock64();
        if (leader && blockIdx.x == 0) { out[0] = t1 - t0; out[1] = t2 - t0; }
    }
    ptx::tc_fence_before(); __syncthreads();
    if (warp == 1) ptx::tmem_dealloc(tm, 512);
}

int main() {
    long long *d; cudaMalloc(&d, 16);
    const int smem = 64 * 1024 + 128 * 1024 + 2048;
    cudaFuncSetAttribute(k<0>, cudaFuncAttributeMaxDynamicSharedMemorySize, smem);
    cudaFuncSetAttribute(k<1>, cudaFuncAttributeMaxDynamicSharedMemorySize, smem);
    const int reps = 256;       // x4 MMAs
    for (int kind : {0, 1})
        for (int grid : {1, 148})
            for (int N : {32, 64, 128, 256})
                for (int nbuf : {1, 4}) {
                    long long h[2];
                    for (int it = 0; it < 2; ++it) {
                        if (kind == 0) k<0><<<grid, 128, smem>>>(N, reps, nbuf, d);
                        else k<1><<<grid, 128, smem>>>(N, reps, nbuf, d);
                    }
                    cudaError_t e = cudaDeviceSynchronize();
                    if (e != cudaSuccess) { printf("err %s\n", cudaGetErrorString(e)); return 1; }
                    cudaMemcpy(h, d, 16, cudaMemcpyDeviceToHost);
                    const double cyc = (double)h[1] / (reps * 4);
                    const double flop = 2.0 * 128 * N * (kind == 0 ? 16 : 8);
                    printf("%s grid=%3d N=%3d nbuf=%d : issue %.1f  total %.1f cyc/mma  -> %.0f flop/cyc/SM (floor %.0f cyc)\n",
                           kind == 0 ? "bf16" : "tf32", grid, N, nbuf, (double)h[0] / (reps * 4), cyc, flop / cyc, 128.0 * N / 256);
                }
    return 0;
}
