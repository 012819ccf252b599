// micro-benchmark 2: what makes an MMA in the conv k-loop cost ~210 cycles instead of ~101?
#include <cstdio>
#include <cuda_runtime.h>
#include "../../vqvae_b200/csrc/ptx.cuh"

__global__ void __launch_bounds__(128) k(int N, int reps, int sbo, int do_wait, int do_fence, int rotate, long long *out) {
    extern __shared__ unsigned char smem_raw[];
    const uint32_t raw = ptx::smem_u32(smem_raw);
    const uint32_t sbase = (raw + 1023u) & ~1023u;
    __shared__ uint64_t bars[4];
    __shared__ uint32_t holder;
    const int warp = threadIdx.x >> 5;
    if (threadIdx.x == 0) {
        ptx::mbar_init(ptx::smem_u32(&bars[0]), 1); ptx::mbar_init(ptx::smem_u32(&bars[1]), 1);
        ptx::fence_mbar_init();
        ptx::mbar_arrive(ptx::smem_u32(&bars[1]));      // bars[1]: phase 0 already complete
    }
    if (warp == 1) ptx::tmem_alloc(ptx::smem_u32(&holder), 512);
    ptx::tc_fence_before(); __syncthreads(); ptx::tc_fence_after();
    const uint32_t tm = holder;
    if (threadIdx.x == 0) {
        const uint32_t idesc = ptx::instr_desc(ptx::FMT_TF32, 128, (uint32_t)N);
        const uint32_t bar = ptx::smem_u32(&bars[0]), done = ptx::smem_u32(&bars[1]);
        long long t0 = clock64();
        for (int r = 0; r < reps; ++r) {
            if (do_wait) ptx::mbar_wait(done, 0);
            if (do_fence) ptx::tc_fence_after();
            const uint32_t a = sbase + (rotate ? (uint32_t)((r % 9) * 128 + ((r % 3) * 2048)) : 0u);
            const uint32_t b = sbase + 65536 + (rotate ? (uint32_t)((r % 8) * 4096) : 0u);
            for (int kk = 0; kk < 4; ++kk)
                ptx::mma_tf32(tm, ptx::smem_desc_sw128_sbo(a + kk * 32, (uint32_t)sbo), ptx::smem_desc_sw128(b + kk * 32), idesc, 1u);
        }
        long long t1 = clock64();
        ptx::tc_commit(bar);
        ptx::mbar_wait(bar, 0);
        long long t2 = clock64();
        out[0] = t1 - t0; out[1] = t2 - t0;
    }
    ptx::tc_fence_before(); __syncthreads();
    if (warp == 1) ptx::tmem_dealloc(tm, 512);
}

int main() {
    long long *d; cudaMalloc(&d, 16);
    cudaFuncSetAttribute(k, cudaFuncAttributeMaxDynamicSharedMemorySize, 160 * 1024);
    const int reps = 64;
    for (int N : {32, 128})
        for (int sbo : {1024, 2048})
            for (int w : {0, 1})
                for (int f : {0, 1})
                    for (int rot : {0, 1}) {
                        long long h[2];
                        k<<<1, 128, 160 * 1024>>>(N, reps, sbo, w, f, rot, d);
                        cudaError_t e = cudaDeviceSynchronize();
                        if (e != cudaSuccess) { printf("err %s\n", cudaGetErrorString(e)); return 1; }
                        cudaMemcpy(h, d, 16, cudaMemcpyDeviceToHost);
                        printf("N=%3d sbo=%4d wait=%d fence=%d rotate=%d : %.1f cyc per mma\n", N, sbo, w, f, rot, (double)h[1] / (reps * 4));
                    }
    return 0;
}
