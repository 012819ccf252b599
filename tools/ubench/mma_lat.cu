// micro-benchmark: tcgen05.mma kind::tf32 issue/latency vs N, accumulator dependence and commit cadence
#include <cstdio>
#include <cuda_runtime.h>
#include "../../vqvae_b200/csrc/ptx.cuh"

__global__ void __launch_bounds__(128) k(int N, int reps, int naccum, int commit_every, long long *out) {
    extern __shared__ unsigned char smem_raw[];
    const uint32_t raw = ptx::smem_u32(smem_raw);
    const uint32_t sbase = (raw + 1023u) & ~1023u;
    __shared__ uint64_t bars[2];
    __shared__ uint32_t holder;
    const int warp = threadIdx.x >> 5;
    if (threadIdx.x == 0) { ptx::mbar_init(ptx::smem_u32(&bars[0]), 1); ptx::fence_mbar_init(); }
    if (warp == 1) ptx::tmem_alloc(ptx::smem_u32(&holder), 512);
    ptx::tc_fence_before(); __syncthreads(); ptx::tc_fence_after();
    const uint32_t tm = holder;
    if (threadIdx.x == 0) {
        const uint32_t idesc = ptx::instr_desc(ptx::FMT_TF32, 128, (uint32_t)N);
        const uint32_t bar = ptx::smem_u32(&bars[0]);
        uint32_t phase = 0;
        long long t0 = clock64();
        for (int r = 0; r < reps; ++r) {
            const uint32_t acc = tm + (uint32_t)((r % naccum) * N);
            for (int kk = 0; kk < 4; ++kk)
                ptx::mma_tf32(acc, ptx::smem_desc_sw128(sbase + kk * 32), ptx::smem_desc_sw128(sbase + 16384 + kk * 32), idesc, 1u);
            if (commit_every > 0 && (r % commit_every) == commit_every - 1) {
                ptx::tc_commit(bar);
                if (commit_every >= 1000) { ptx::mbar_wait(bar, phase); phase ^= 1; }
            }
        }
        long long t1 = clock64();
        ptx::tc_commit(bar);
        ptx::mbar_wait(bar, phase);
        long long t2 = clock64();
        out[0] = t1 - t0; out[1] = t2 - t0;
    }
    ptx::tc_fence_before(); __syncthreads();
    if (warp == 1) ptx::tmem_dealloc(tm, 512);
}

int main() {
    long long *d; cudaMalloc(&d, 16);
    cudaFuncSetAttribute(k, cudaFuncAttributeMaxDynamicSharedMemorySize, 100 * 1024);
    const int reps = 64;
    for (int N : {16, 32, 64, 128, 256})
        for (int naccum : {1, 2, 4}) {
            if (naccum * N > 512) continue;
            for (int ce : {0, 1, 4}) {
                long long h[2];
                k<<<1, 128, 100 * 1024>>>(N, reps, naccum, ce, d);
                cudaError_t e = cudaDeviceSynchronize();
                if (e != cudaSuccess) { printf("err %s\n", cudaGetErrorString(e)); return 1; }
                cudaMemcpy(h, d, 16, cudaMemcpyDeviceToHost);
                printf("N=%3d accum=%d commit_every=%d : issue %6lld cyc  total %6lld cyc  -> %.1f cyc per mma (x%d)\n", N, naccum, ce,
                       h[0], h[1], (double)h[1] / (reps * 4), reps * 4);
            }
        }
    return 0;
}
