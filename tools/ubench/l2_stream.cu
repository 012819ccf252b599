// micro-benchmark (round 2): how fast can every SM stream an L2-resident weight set into shared memory?
// One CTA per SM; a producer lane issues cp.async.bulk (1-D, `chunk` bytes) from a `wbytes` buffer (cyclic) into an
// S-stage ring; a consumer warp waits on the full barrier and releases the stage.  Reports GB/s per SM and chip-wide.
// Build: nvcc -gencode arch=compute_100a,code=sm_100a -O3 -o l2_stream l2_stream.cu
#include <cstdio>
#include <cuda_runtime.h>
#include "../../vqvae_b200/csrc/ptx.cuh"

__global__ void __launch_bounds__(64) k(const unsigned char *w, int wbytes, int chunk, int S, int iters, int per_cta_offset,
                                         unsigned long long *out) {
    extern __shared__ unsigned char smem_raw[];
    const uint32_t raw = ptx::smem_u32(smem_raw);
    const uint32_t sbase = (raw + 1023u) & ~1023u;
    __shared__ uint64_t bars[64];
    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    if (threadIdx.x < S) { ptx::mbar_init(ptx::smem_u32(&bars[threadIdx.x]), 1); ptx::mbar_init(ptx::smem_u32(&bars[32 + threadIdx.x]), 1); }
    ptx::fence_mbar_init();
    __syncthreads();
    unsigned long long t0, t1;
    asm volatile("mov.u64 %0, %%globaltimer;" : "=l"(t0));
    if (warp == 0) {
        if (lane == 0) {
            int off = per_cta_offset ? (int)((blockIdx.x * 37u * (unsigned)chunk) % (unsigned)wbytes) : 0;
            uint32_t st = 0, par = 0;
            for (int i = 0; i < iters; ++i) {
                ptx::mbar_wait(ptx::smem_u32(&bars[32 + st]), par ^ 1);
                ptx::mbar_expect_tx(ptx::smem_u32(&bars[st]), (uint32_t)chunk);
                ptx::bulk_load_1d(sbase + st * (uint32_t)chunk, w + off, (uint32_t)chunk, ptx::smem_u32(&bars[st]));
                off += chunk; if (off >= wbytes) off = 0;
                if (++st == (uint32_t)S) { st = 0; par ^= 1; }
            }
        }
    } else {
        uint32_t st = 0, par = 0;
        for (int i = 0; i < iters; ++i) {
            ptx::mbar_wait(ptx::smem_u32(&bars[st]), par);
            if (lane == 0) ptx::mbar_arrive(ptx::smem_u32(&bars[32 + st]));
            __syncwarp();
            if (++st == (uint32_t)S) { st = 0; par ^= 1; }
        }
    }
    __syncthreads();
    asm volatile("mov.u64 %0, %%globaltimer;" : "=l"(t1));
    if (threadIdx.x == 0) { out[2 * blockIdx.x] = t0; out[2 * blockIdx.x + 1] = t1; }
}

int main() {
    const int wmax = 1 << 20;
    unsigned char *w; cudaMalloc(&w, wmax); cudaMemset(w, 1, wmax);
    unsigned long long *d; cudaMalloc(&d, 16 * 256);
    cudaFuncSetAttribute(k, cudaFuncAttributeMaxDynamicSharedMemorySize, 200 * 1024);
    for (int grid : {1, 148})
        for (int wbytes : {288 * 1024})
            for (int chunk : {8192, 16384, 32768})
                for (int S : {2, 4, 6, 12}) {
                    if (S * chunk > 196 * 1024) continue;
                    for (int pco : {0, 1}) {
                        const int iters = 2048;
                        unsigned long long h[512];
                        for (int it = 0; it < 2; ++it) k<<<grid, 64, 200 * 1024>>>(w, wbytes, chunk, S, iters, pco, d);
                        cudaError_t e = cudaDeviceSynchronize();
                        if (e != cudaSuccess) { printf("err %s\n", cudaGetErrorString(e)); return 1; }
                        cudaMemcpy(h, d, 16 * grid, cudaMemcpyDeviceToHost);
                        unsigned long long lo = ~0ull, hi = 0; double sum = 0;
                        for (int b = 0; b < grid; ++b) { if (h[2 * b] < lo) lo = h[2 * b]; if (h[2 * b + 1] > hi) hi = h[2 * b + 1]; sum += (double)(h[2 * b + 1] - h[2 * b]); }
                        const double bytes = (double)iters * chunk;
                        printf("grid=%3d chunk=%5d S=%2d offset=%d : per-SM %.1f GB/s (mean CTA time)  chip %.2f TB/s (span)\n", grid, chunk, S, pco,
                               bytes / (sum / grid), bytes * grid / (double)(hi - lo) / 1e3);
                    }
                }
    return 0;
}
