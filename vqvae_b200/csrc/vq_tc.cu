// vq_tc.cu -- fused VectorQuantizer.forward on tcgen05 tensor cores (sm_100a), D = 64.
//
// Replaces quantizer.py:45-71.  One persistent, warp-specialised CTA per SM:
//   warp 0      TMA producer: z row tiles (128 x 64 fp32) and codebook chunks
//               (256 x 64 fp32) land in shared memory in the 128-byte-swizzled K-major
//               layout UMMA consumes; ||e_k||^2 of the chunk comes with a 1-D bulk copy.
//               A codebook of <= 512 codes stays resident in shared memory.  The same
//               thread TMA-stores finished z_q tiles (written in place over the z tile).
//   warp 1      issues tcgen05.mma kind::tf32 (M=128, N=256, K=8 x 8): the dense
//               contraction z . e^T, fp32 accumulators in TMEM, double buffered
//               (2 x 256 columns) so the epilogue of chunk c overlaps the MMA of c+1.
//   warp 2      allocates / frees TMEM.
//   warps 4-11  epilogue: thread = (row, column half).  tcgen05.ld the scores (double
//               buffered in registers), form s = ||e||^2 - 2 z.e, keep per-8-code group
//               minima, and push every group within tau of the running minimum into a
//               small per-thread list.
//
// Bit-exactness (DESIGN.md "VQ arithmetic contract"): the TF32 scores only SELECT
// candidates.  tau bounds twice the worst-case error of a score (tf32 truncation of
// both operands, fp32 accumulation, and the rounding of the canonical formula), so the
// canonical fp32 winner -- and every code tied with it -- is always inside a listed
// group.  Listed groups are re-scored with the canonical arithmetic of
// oracle/csrc/oracle.c (sequential fmaf chain, fl(fl(A+B) - fl(2M)), first minimum
// wins, NaN wins), so idx and z_q are bit-identical to vq_exact.cu and to the oracle.
// The two threads of a row split each candidate group (4 codes each, 4 independent
// chains).  Rows with non-finite data, a non-finite codebook, or overflowing candidate
// lists fall back to scanning every code exactly.
//
// The same launch gathers e_idx, forms z_q = z + (e - z), accumulates the SSE (double)
// and a shared-memory code histogram.
#include <cstdlib>

#include <cuda_bf16.h>

#include "common.cuh"
#include "ptx.cuh"

__device__ unsigned long long g_vqb_trace_vq[64];            // diagnostic timeline (vqb_debug_read_trace_vq)
extern "C" int vqb_debug_read_trace_vq(unsigned long long *dst, int n) {
    if (!dst || n < 1 || n > 64) return VQB_ERR_BAD_ARG;
    return vqb_cuda_status(cudaMemcpyFromSymbol(dst, g_vqb_trace_vq, sizeof(unsigned long long) * n));
}

__device__ unsigned long long g_vqb_cta_t[512];              // per-CTA (start, end) globaltimer, flags & 8
extern "C" int vqb_debug_read_cta_times(unsigned long long *dst, int n) {
    if (!dst || n < 1 || n > 512) return VQB_ERR_BAD_ARG;
    return vqb_cuda_status(cudaMemcpyFromSymbol(dst, g_vqb_cta_t, sizeof(unsigned long long) * n));
}

namespace {

// timeline of epilogue warp 4 / CTA 0 for local tiles 1 and 2 (16 marks each), globaltimer ns
__device__ int g_vqb_trace_tile = 1;      // first of the two local tiles that are traced
__device__ __forceinline__ void vq_mark(bool on, int it, int i) {
    const int t0 = g_vqb_trace_tile;
    if (on && it >= t0 && it <= t0 + 1) {
        unsigned long long t;
        asm volatile("mov.u64 %0, %%globaltimer;" : "=l"(t));
        g_vqb_trace_vq[(it - t0) * 16 + i] = t;
    }
}

constexpr int TM = 128;          // latent rows per tile (UMMA M)
constexpr int CN = 256;          // codes per chunk (UMMA N)
constexpr int DD = 64;           // embedding dim handled by this kernel
constexpr int NTHREADS = 384;    // 12 warps
constexpr int LCAP = 12;         // group list capacity per thread (pass 1)
constexpr int ZSTAGE = TM * DD * 4, ZATOM = TM * 128;
constexpr int ESTAGE = CN * DD * 4, EATOM = CN * 128;
constexpr int HIST_MAX = 1024;

constexpr int OFF_Z = 0;
constexpr int OFF_E = OFF_Z + 2 * ZSTAGE;
constexpr int OFF_B = OFF_E + 2 * ESTAGE;
constexpr int OFF_LIST = OFF_B + 2 * CN * 4;
constexpr int OFF_XMIN = OFF_LIST + 256 * LCAP * 8;      // int nc[256]: compacted list lengths
constexpr int OFF_XBD = OFF_XMIN + 256 * 4;              // float[256]
constexpr int OFF_XBK = OFF_XBD + 256 * 4;               // int[256]
constexpr int OFF_HIST = OFF_XBK + 256 * 4;
constexpr int OFF_BAR = OFF_HIST + HIST_MAX * 4;
constexpr int OFF_TMEM = OFF_BAR + 16 * 8;
constexpr int OFF_RED = OFF_TMEM + 64;
constexpr int SMEM_TOTAL = OFF_RED + 64;
constexpr int SMEM_ALLOC = SMEM_TOTAL + 1024;   // slack for the manual 1024-byte alignment
static_assert(SMEM_ALLOC <= 227 * 1024, "shared memory budget");

enum { Z_FULL = 0, Q_FULL = 2, E_FULL = 4, E_EMPTY = 6, T_FULL = 8, T_EMPTY = 10 };

struct VqTcParams {
    const float *E;        // (K, 64) codebook
    const float *bn;       // (nchunks*256) canonical ||e_k||^2, +inf past K
    const float *scal;     // [0] = upper bound of max ||e_k||, [1] = non-finite flag (int)
    long long N;
    int K, nchunks;
    long long *idx;
    double *partials;
    unsigned *pending;     // number of SSE partials this launch leaves in `partials` (read by vqb_vq_reduce_sse_f32)
    int *hist;
    float *dbg;            // optional (N, nchunks*256) raw approximate scores
    int flags;             // perf-experiment knobs (env VQB_TC_FLAGS), 0 in production
    int zq_bf16;           // VQB_BF16 pipeline: z_q leaves as bf16 rows (the decoder's first conv reads bf16)
};

__device__ __forceinline__ bool vq_better(float dn, int kn, float db, int kb) {
    const bool nn = dn != dn, nb = db != db;            // torch.argmin: NaN is the minimum
    if (nn || nb) return nn && (!nb || kn < kb);
    return dn < db || (dn == db && kn < kb);
}

__device__ __forceinline__ void tma_store_2d(const CUtensorMap *m, uint32_t src, int c0, int c1) {
    asm volatile("cp.async.bulk.tensor.2d.global.shared::cta.bulk_group [%0, {%2, %3}], [%1];" ::
                     "l"(reinterpret_cast<uint64_t>(m)), "r"(src), "r"(c0), "r"(c1) : "memory");
}
__device__ __forceinline__ void bulk_commit() { asm volatile("cp.async.bulk.commit_group;" ::: "memory"); }
__device__ __forceinline__ void bulk_wait_read0() { asm volatile("cp.async.bulk.wait_group.read 0;" ::: "memory"); }
__device__ __forceinline__ void bulk_wait_all0() { asm volatile("cp.async.bulk.wait_group 0;" ::: "memory"); }

// workspace prep: canonical code norms (quantizer.py:50), their maximum, a non-finite flag
__global__ void vq_tc_prep_kernel(const float *__restrict__ E, int K, int Kpad, float *__restrict__ bn,
                                  unsigned *__restrict__ scal) {
    const int k = blockIdx.x * blockDim.x + threadIdx.x;
    if (k >= Kpad) return;
    if (k >= K) { bn[k] = __int_as_float(0x7f800000); return; }
    const float *e = E + (size_t)k * DD;
    float s = 0.f;
    for (int d = 0; d < DD; ++d) s = __fadd_rn(s, __fmul_rn(e[d], e[d]));
    bn[k] = s;
    if (!(s < __int_as_float(0x7f800000))) atomicOr(&scal[1], 1u);       // inf or NaN
    else atomicMax(&scal[0], __float_as_uint(sqrtf(s) * 1.00001f));       // positive floats order as uints
}

template <bool kDebug>
__global__ void __launch_bounds__(NTHREADS, 1)
vq_tc_kernel(const __grid_constant__ CUtensorMap tmz, const __grid_constant__ CUtensorMap tme,
             const __grid_constant__ CUtensorMap tmq, const VqTcParams p) {
    extern __shared__ unsigned char smem_raw[];
    const uint32_t raw = ptx::smem_u32(smem_raw);
    const uint32_t sbase = (raw + 1023u) & ~1023u;
    unsigned char *sm = smem_raw + (sbase - raw);

    const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
    const int kflags = VQB_DIAG ? p.flags : 0;          // release build: every diagnostic branch folds away
    auto kmark = [&](int i) {          // kernel-level timeline of CTA 0 (slots 32..), VQB_TC_FLAGS & 8
        if ((kflags & 8) && blockIdx.x == 0) {
            unsigned long long t;
            asm volatile("mov.u64 %0, %%globaltimer;" : "=l"(t));
            g_vqb_trace_vq[32 + i] = t;
        }
    };
    if (tid == 128) kmark(0);
    if (tid == 128 && (kflags & 8) && blockIdx.x < 256) {
        unsigned long long t;
        asm volatile("mov.u64 %0, %%globaltimer;" : "=l"(t));
        g_vqb_cta_t[2 * blockIdx.x] = t;
    }
    const uint32_t bars = sbase + OFF_BAR;
    auto bar = [&](int i) { return bars + 8u * (uint32_t)i; };
    float *bsm = reinterpret_cast<float *>(sm + OFF_B);
    int *hist_s = reinterpret_cast<int *>(sm + OFF_HIST);
    volatile uint32_t *tmem_holder = reinterpret_cast<volatile uint32_t *>(sm + OFF_TMEM);

    const long long ntiles = (p.N + TM - 1) / TM;
    const int nchunks = p.nchunks;
    const bool resident = nchunks <= 2;
    const bool smem_hist = p.K <= HIST_MAX;

    if (tid == 0) {
        ptx::prefetch_tmap(&tmz);
        ptx::prefetch_tmap(&tme);
        ptx::prefetch_tmap(&tmq);
        ptx::mbar_init(bar(Z_FULL + 0), 1); ptx::mbar_init(bar(Z_FULL + 1), 1);
        ptx::mbar_init(bar(Q_FULL + 0), 8); ptx::mbar_init(bar(Q_FULL + 1), 8);     // 8 epilogue warps
        ptx::mbar_init(bar(E_FULL + 0), 1); ptx::mbar_init(bar(E_FULL + 1), 1);
        ptx::mbar_init(bar(E_EMPTY + 0), 9); ptx::mbar_init(bar(E_EMPTY + 1), 9);   // MMA commit + 8 warps
        ptx::mbar_init(bar(T_FULL + 0), 1); ptx::mbar_init(bar(T_FULL + 1), 1);
        ptx::mbar_init(bar(T_EMPTY + 0), 8); ptx::mbar_init(bar(T_EMPTY + 1), 8);
        ptx::fence_mbar_init();
    }
    if (smem_hist)
        for (int k = tid; k < p.K; k += NTHREADS) hist_s[k] = 0;
    if (warp == 2) ptx::tmem_alloc(sbase + OFF_TMEM, 512);
    ptx::tc_fence_before();
    __syncthreads();
    ptx::tc_fence_after();
    const uint32_t tmem_base = *tmem_holder;
    pdl_launch_dependents();
    pdl_wait();                    // z_e is written by the previous layer

    if (warp == 0) {
        // ===================== TMA producer (+ z_q tile stores) =====================
        if (lane == 0) {
            long long gc = 0;
            int it = 0;
            long long tile = blockIdx.x;
            // a z stage is recycled once its tile's z_q (written in place by the epilogue)
            // has been TMA-stored; Q_FULL also implies the MMAs of that tile are done.
            auto drain = [&](int jt, long long jtile) {
                const int zs = jt & 1;
                ptx::mbar_wait(bar(Q_FULL + zs), (jt >> 1) & 1);
                if (!(kflags & 4)) {
                    const uint32_t src = sbase + OFF_Z + zs * ZSTAGE;
                    tma_store_2d(&tmq, src, 0, (int)(jtile * TM));
                    if (!p.zq_bf16) tma_store_2d(&tmq, src + ZATOM, 32, (int)(jtile * TM));     // (bf16: 64 channels = one atom)
                    bulk_commit();
                    bulk_wait_read0();
                }
            };
            for (; tile < ntiles; tile += gridDim.x, ++it) {
                const int zs = it & 1;
                if (it >= 2) drain(it - 2, tile - 2 * (long long)gridDim.x);
                ptx::mbar_expect_tx(bar(Z_FULL + zs), ZSTAGE);
                const uint32_t zdst = sbase + OFF_Z + zs * ZSTAGE;
                ptx::tma_load_2d(zdst, &tmz, bar(Z_FULL + zs), 0, (int)(tile * TM));
                ptx::tma_load_2d(zdst + ZATOM, &tmz, bar(Z_FULL + zs), 32, (int)(tile * TM));
                if (resident && it > 0) continue;
                for (int c = 0; c < nchunks; ++c, ++gc) {
                    const int es = resident ? c : (int)(gc & 1);
                    const uint32_t par = resident ? 0u : (uint32_t)((gc >> 1) & 1);
                    ptx::mbar_wait(bar(E_EMPTY + es), par ^ 1);
                    ptx::mbar_expect_tx(bar(E_FULL + es), resident ? ESTAGE : ESTAGE + CN * 4);
                    const uint32_t edst = sbase + OFF_E + es * ESTAGE;
                    ptx::tma_load_2d(edst, &tme, bar(E_FULL + es), 0, c * CN);
                    ptx::tma_load_2d(edst + EATOM, &tme, bar(E_FULL + es), 32, c * CN);
                    if (!resident)      // resident codebooks get their norms computed in-kernel (below)
                        ptx::bulk_load_1d(sbase + OFF_B + es * CN * 4, p.bn + (size_t)c * CN, CN * 4, bar(E_FULL + es));
                }
            }
            // `it` tiles were issued; the last (up to) two are still to be stored
            for (int jt = (it >= 2 ? it - 2 : 0); jt < it; ++jt)
                drain(jt, (long long)blockIdx.x + (long long)jt * gridDim.x);
            bulk_wait_all0();
        }
    } else if (warp == 1) {
        // ===================== MMA issuer (converged warp, elected leader lane issues) =====================
        {
            const bool leader = ptx::elect_one();
            constexpr uint32_t idesc = ptx::instr_desc(ptx::FMT_TF32, TM, CN);
            const uint32_t d_hi = ptx::desc_hi_sw128(1024);
            int it = 0;
            for (long long tile = blockIdx.x; tile < ntiles; tile += gridDim.x, ++it) {
                const int zs = it & 1;
                ptx::mbar_wait(bar(Z_FULL + zs), (it >> 1) & 1);
                const uint32_t za_lo = (sbase + OFF_Z + zs * ZSTAGE) >> 4;
                for (int c = 0; c < nchunks; ++c) {
                    const long long gc = (long long)it * nchunks + c;
                    int es;
                    if (resident) {
                        es = c;
                        if (it == 0) ptx::mbar_wait(bar(E_FULL + es), 0);
                    } else {
                        es = (int)(gc & 1);
                        ptx::mbar_wait(bar(E_FULL + es), (uint32_t)((gc >> 1) & 1));
                    }
                    const int ab = (int)(gc & 1);
                    ptx::mbar_wait(bar(T_EMPTY + ab), (uint32_t)(((gc >> 1) & 1) ^ 1));
                    ptx::tc_fence_after();
                    const uint32_t ea_lo = (sbase + OFF_E + es * ESTAGE) >> 4;
#pragma unroll
                    for (int ks = 0; ks < DD / 8; ++ks)
                        if (leader)
                            ptx::mma_tf32_w(tmem_base + ab * CN, za_lo + (ks >> 2) * (ZATOM >> 4) + (ks & 3) * 2, d_hi,
                                            ea_lo + (ks >> 2) * (EATOM >> 4) + (ks & 3) * 2, d_hi, idesc, ks > 0 ? 1u : 0u);
                    if (leader) {
                        ptx::tc_commit(bar(T_FULL + ab));
                        if (!resident) ptx::tc_commit(bar(E_EMPTY + es));
                    }
                    __syncwarp();
                }
            }
        }
    } else if (warp >= 4) {
        // ===================== epilogue =====================
        const int et = tid - 128;               // 0..255
        const int q = warp & 3;                 // TMEM lane quadrant this warp may read
        const int h = (warp - 4) >> 2;          // column half of every chunk
        const int row = q * 32 + lane;          // accumulator row = TMEM lane
        const int rsw = row & 7;
        const float INF = __int_as_float(0x7f800000);
        float Emax;
        bool bad_codebook;
        if (resident) {
            // The <= 512 resident codes: canonical ||e_k||^2 (quantizer.py:50), their maximum and a
            // non-finite flag, computed once per CTA from the shared-memory copy (no prep launch).
            float *xred = reinterpret_cast<float *>(sm + OFF_XBD);      // scratch before the first tile
            float mymax = 0.f;
            unsigned mybad = 0u;
            for (int c = 0; c < nchunks; ++c) ptx::mbar_wait(bar(E_FULL + c), 0);
            for (int k = et; k < nchunks * CN; k += 256) {
                float sn = INF;
                if (k < p.K) {
                    const unsigned char *er = sm + OFF_E + (k >> 8) * ESTAGE + (k & 255) * 128;
                    sn = 0.f;
#pragma unroll
                    for (int a = 0; a < 2; ++a)
#pragma unroll
                        for (int c16 = 0; c16 < 8; ++c16) {
                            const float4 v = *reinterpret_cast<const float4 *>(er + a * EATOM + ((c16 ^ (k & 7)) << 4));
                            sn = __fadd_rn(sn, __fmul_rn(v.x, v.x)); sn = __fadd_rn(sn, __fmul_rn(v.y, v.y));
                            sn = __fadd_rn(sn, __fmul_rn(v.z, v.z)); sn = __fadd_rn(sn, __fmul_rn(v.w, v.w));
                        }
                    if (!(sn < INF)) mybad = 1u;
                    else mymax = fmaxf(mymax, sqrtf(sn) * 1.00001f);
                }
                bsm[k] = sn;
            }
#pragma unroll
            for (int off = 16; off >= 1; off >>= 1) {
                mymax = fmaxf(mymax, __shfl_xor_sync(0xffffffffu, mymax, off));
                mybad |= __shfl_xor_sync(0xffffffffu, mybad, off);
            }
            if (lane == 0) { xred[warp - 4] = mymax; xred[8 + warp - 4] = __uint_as_float(mybad); }
            ptx::named_bar_sync(5, 256);
            float mx = 0.f;
            unsigned bad = 0u;
            for (int w = 0; w < 8; ++w) { mx = fmaxf(mx, xred[w]); bad |= __float_as_uint(xred[8 + w]); }
            ptx::named_bar_sync(5, 256);            // xred (= xbd) is reused by the tile loop
            Emax = mx;
            bad_codebook = bad != 0u;
        } else {
            Emax = __uint_as_float(reinterpret_cast<const unsigned *>(p.scal)[0]);
            bad_codebook = reinterpret_cast<const unsigned *>(p.scal)[1] != 0u;
        }
        float2 *lists = reinterpret_cast<float2 *>(sm + OFF_LIST);
        int *xnc = reinterpret_cast<int *>(sm + OFF_XMIN);
        float *xbd = reinterpret_cast<float *>(sm + OFF_XBD);
        int *xbk = reinterpret_cast<int *>(sm + OFF_XBK);
        const int Kpad = nchunks * CN;
        const uint32_t lane_taddr = tmem_base + ((uint32_t)(q * 32) << 16);
        double sse = 0.0;

        auto code_ptr_smem = [&](int k) -> const unsigned char * {
            return sm + OFF_E + (k >> 8) * ESTAGE + (k & 255) * 128;
        };

        if (tid == 128) kmark(1);          // setup + codebook norms done
        int it = 0;
        for (long long tile = blockIdx.x; tile < ntiles; tile += gridDim.x, ++it) {
            const int zs = it & 1;
            unsigned char *zrow = sm + OFF_Z + zs * ZSTAGE + row * 128;
            const bool tr = kDebug == false && (kflags & 8) && blockIdx.x == 0 && tid == 128;
            vq_mark(tr, it, 0);
            ptx::mbar_wait(bar(Z_FULL + zs), (it >> 1) & 1);
            vq_mark(tr, it, 1);

            // ---- A_i = sum_d fl(z^2), canonical left-to-right order (quantizer.py:49) ----
            float A = 0.f;
#pragma unroll
            for (int a = 0; a < 2; ++a)
#pragma unroll
                for (int c16 = 0; c16 < 8; ++c16) {
                    const float4 v = *reinterpret_cast<const float4 *>(zrow + a * ZATOM + ((c16 ^ rsw) << 4));
                    A = __fadd_rn(A, __fmul_rn(v.x, v.x)); A = __fadd_rn(A, __fmul_rn(v.y, v.y));
                    A = __fadd_rn(A, __fmul_rn(v.z, v.z)); A = __fadd_rn(A, __fmul_rn(v.w, v.w));
                }
            // S >= sum_d |z_d e_kd| for every k (Cauchy-Schwarz, rounded up)
            const float S = sqrtf(A) * 1.00001f * Emax;
            // 2 x (tf32 truncation of both operands on 2M: 2*2^-9*S, + fp32 accumulation and
            // the canonical formula's own rounding), with margin.
            const float tau = S * (0.0078125f + 0.0009765625f) + (A + Emax * Emax + S) * 1.9073486e-6f;
            bool slow_row = bad_codebook || !(A < INF) || !(tau < INF);

            // ---- pass 1: approximate scores -> group minima -> candidate list ----
            float run_min = INF, thr = INF;
            int cnt = 0;
            for (int c = 0; c < nchunks; ++c) {
                const long long gc = (long long)it * nchunks + c;
                const int ab = (int)(gc & 1);
                const int es = resident ? c : (int)(gc & 1);
                vq_mark(tr, it, 2 + 2 * (c & 1));
                ptx::mbar_wait(bar(T_FULL + ab), (uint32_t)((gc >> 1) & 1));
                ptx::tc_fence_after();
                vq_mark(tr, it, 3 + 2 * (c & 1));
                const float *bch = bsm + es * CN + h * 128;
                const uint32_t tcol = lane_taddr + (uint32_t)(ab * CN + h * 128);
                const int gbase = (c * CN + h * 128) / 8;
                float va[32], vb[32], warm[8];
                auto process = [&](const float (&v)[32], int j) {
#pragma unroll
                    for (int g = 0; g < 4; ++g) {
                        const float4 b0 = *reinterpret_cast<const float4 *>(bch + j * 32 + g * 8);
                        const float4 b1 = *reinterpret_cast<const float4 *>(bch + j * 32 + g * 8 + 4);
                        const float s0 = fmaf(v[g * 8 + 0], -2.f, b0.x), s1 = fmaf(v[g * 8 + 1], -2.f, b0.y);
                        const float s2 = fmaf(v[g * 8 + 2], -2.f, b0.z), s3 = fmaf(v[g * 8 + 3], -2.f, b0.w);
                        const float s4 = fmaf(v[g * 8 + 4], -2.f, b1.x), s5 = fmaf(v[g * 8 + 5], -2.f, b1.y);
                        const float s6 = fmaf(v[g * 8 + 6], -2.f, b1.z), s7 = fmaf(v[g * 8 + 7], -2.f, b1.w);
                        if (kDebug && p.dbg) {
                            const long long grow = tile * TM + row;
                            if (grow < p.N) {
                                float *dst = p.dbg + (size_t)grow * Kpad + c * CN + h * 128 + j * 32 + g * 8;
                                dst[0] = s0; dst[1] = s1; dst[2] = s2; dst[3] = s3;
                                dst[4] = s4; dst[5] = s5; dst[6] = s6; dst[7] = s7;
                            }
                        }
                        const float gm = ptx::fmin3(ptx::fmin3(s0, s1, s2), ptx::fmin3(s3, s4, s5), fminf(s6, s7));
                        if (c == 0 && j < 2) {
                            // warm-up: the first 8 group minima only seed the running minimum; they are
                            // pushed afterwards against min8 + tau.  A scan that pushes every record low
                            // lists H(64) ~ 4.7 groups per thread with a long tail; list overflows (-> exact
                            // scan of all K codes by the whole warp) cost about half of the kernel time.
                            warm[j * 4 + g] = gm;
                            run_min = fminf(run_min, gm);
                        } else {
                            // branch-free push: always write slot min(cnt, LCAP-1), keep it when in range
                            lists[min(cnt, LCAP - 1) * 256 + et] = make_float2(gm, __int_as_float(gbase + j * 4 + g));
                            cnt += (gm <= thr) ? 1 : 0;
                            run_min = fminf(run_min, gm);
                            thr = run_min + tau;
                        }
                    }
                    if (c == 0 && j == 1) {
                        thr = run_min + tau;
#pragma unroll
                        for (int i = 0; i < 8; ++i) {
                            lists[min(cnt, LCAP - 1) * 256 + et] = make_float2(warm[i], __int_as_float(gbase + i));
                            cnt += (warm[i] <= thr) ? 1 : 0;
                        }
                    }
                };
                if (!(kflags & 2)) {
                    ptx::tmem_ld32(tcol, va);
                    ptx::tmem_ld_wait32(va);
                    ptx::tmem_ld32(tcol + 32, vb);
                    process(va, 0);
                    ptx::tmem_ld_wait32(vb);
                    ptx::tmem_ld32(tcol + 64, va);
                    process(vb, 1);
                    ptx::tmem_ld_wait32(va);
                    ptx::tmem_ld32(tcol + 96, vb);
                    process(va, 2);
                    ptx::tmem_ld_wait32(vb);
                    process(vb, 3);
                }
                ptx::tc_fence_before();
                __syncwarp();
                if (lane == 0) {
                    ptx::mbar_arrive(bar(T_EMPTY + ab));
                    if (!resident) ptx::mbar_arrive(bar(E_EMPTY + es));
                }
            }

            vq_mark(tr, it, 6);
            // ---- exchange with the partner thread (same row, other column half).  The running minimum travels
            // through the scratch slot of this thread's own list (slot LCAP-1, never a valid entry). ----
            lists[(LCAP - 1) * 256 + et] = make_float2(run_min, __int_as_float((slow_row || cnt >= LCAP) ? -1 : 0));
            ptx::named_bar_sync(1 + q, 64);
            {
                const float2 pinfo = lists[(LCAP - 1) * 256 + (et ^ 128)];
                thr = fminf(run_min, pinfo.x) + tau;        // approximate minimum of the whole row + tau
                if (cnt >= LCAP || __float_as_int(pinfo.y) < 0) slow_row = true;   // a full list -> exact scan of every code
            }
            // ---- compact the list in place down to the groups that can hold the canonical winner; the partner
            // reads them straight from shared memory (no cap on their number) ----
            int nc = 0;
            if (!slow_row) {
                for (int sidx = 0; sidx < cnt; ++sidx) {
                    const float2 ent = lists[sidx * 256 + et];
                    if (ent.x <= thr) { lists[nc * 256 + et] = ent; ++nc; }
                }
            }
            xnc[et] = slow_row ? -1 : nc;
            ptx::named_bar_sync(1 + q, 64);
            const int pnc = xnc[et ^ 128];
            if (pnc < 0) slow_row = true;

            vq_mark(tr, it, 7);
            // ---- z row -> registers (needed for the exact chains and for z_q) ----
            float zr[DD];
#pragma unroll
            for (int a = 0; a < 2; ++a)
#pragma unroll
                for (int c16 = 0; c16 < 8; ++c16) {
                    const float4 v = *reinterpret_cast<const float4 *>(zrow + a * ZATOM + ((c16 ^ rsw) << 4));
                    zr[a * 32 + c16 * 4 + 0] = v.x; zr[a * 32 + c16 * 4 + 1] = v.y;
                    zr[a * 32 + c16 * 4 + 2] = v.z; zr[a * 32 + c16 * 4 + 3] = v.w;
                }

            vq_mark(tr, it, 8);
            // ---- exact canonical re-scoring ----
            float bd = 0.f;
            int bk = -1;
            auto consider = [&](float M, int k) {
                if (k < p.K) {
                    const float bnk = resident ? bsm[k] : __ldg(p.bn + k);
                    const float dist = __fsub_rn(__fadd_rn(A, bnk), __fmul_rn(2.0f, M));   // quantizer.py:49-51
                    if (bk < 0 || vq_better(dist, k, bd, bk)) { bd = dist; bk = k; }
                }
            };
            // This thread's four codes of an 8-code group.  Roles are rotated per lane so that at
            // every load the 32 lanes of a warp spread evenly over the eight 16-byte bank groups
            // of the swizzled rows (row k keeps chunk c16 at position c16 ^ (k & 7)); the partner
            // thread (same lane, other column half) takes the complementary four codes.
            const int jbase = 4 * (h ^ (lane & 1)), jrot = lane >> 1;
            const int j0 = jbase + ((jrot + 0) & 3), j1 = jbase + ((jrot + 1) & 3);
            const int j2 = jbase + ((jrot + 2) & 3), j3 = jbase + ((jrot + 3) & 3);
            auto rescore_group = [&](int g) {
                const int kg = g * 8;
                float M0 = 0.f, M1 = 0.f, M2 = 0.f, M3 = 0.f;     // four independent sequential chains
                if (resident) {
                    const unsigned char *eg = code_ptr_smem(kg);
                    const unsigned char *r0 = eg + j0 * 128, *r1 = eg + j1 * 128, *r2 = eg + j2 * 128, *r3 = eg + j3 * 128;
#pragma unroll
                    for (int a = 0; a < 2; ++a)
#pragma unroll
                        for (int c16 = 0; c16 < 8; ++c16) {
                            const float4 e0 = *reinterpret_cast<const float4 *>(r0 + a * EATOM + ((c16 ^ j0) << 4));
                            const float4 e1 = *reinterpret_cast<const float4 *>(r1 + a * EATOM + ((c16 ^ j1) << 4));
                            const float4 e2 = *reinterpret_cast<const float4 *>(r2 + a * EATOM + ((c16 ^ j2) << 4));
                            const float4 e3 = *reinterpret_cast<const float4 *>(r3 + a * EATOM + ((c16 ^ j3) << 4));
                            const float z0 = zr[a * 32 + c16 * 4 + 0], z1 = zr[a * 32 + c16 * 4 + 1];
                            const float z2 = zr[a * 32 + c16 * 4 + 2], z3 = zr[a * 32 + c16 * 4 + 3];
                            M0 = __fmaf_rn(z0, e0.x, M0); M1 = __fmaf_rn(z0, e1.x, M1); M2 = __fmaf_rn(z0, e2.x, M2); M3 = __fmaf_rn(z0, e3.x, M3);
                            M0 = __fmaf_rn(z1, e0.y, M0); M1 = __fmaf_rn(z1, e1.y, M1); M2 = __fmaf_rn(z1, e2.y, M2); M3 = __fmaf_rn(z1, e3.y, M3);
                            M0 = __fmaf_rn(z2, e0.z, M0); M1 = __fmaf_rn(z2, e1.z, M1); M2 = __fmaf_rn(z2, e2.z, M2); M3 = __fmaf_rn(z2, e3.z, M3);
                            M0 = __fmaf_rn(z3, e0.w, M0); M1 = __fmaf_rn(z3, e1.w, M1); M2 = __fmaf_rn(z3, e2.w, M2); M3 = __fmaf_rn(z3, e3.w, M3);
                        }
                } else {
                    const int kl = p.K - 1;     // clamp the address; consider() drops k >= K
                    const float4 *r0 = reinterpret_cast<const float4 *>(p.E + (size_t)min(kg + j0, kl) * DD);
                    const float4 *r1 = reinterpret_cast<const float4 *>(p.E + (size_t)min(kg + j1, kl) * DD);
                    const float4 *r2 = reinterpret_cast<const float4 *>(p.E + (size_t)min(kg + j2, kl) * DD);
                    const float4 *r3 = reinterpret_cast<const float4 *>(p.E + (size_t)min(kg + j3, kl) * DD);
#pragma unroll
                    for (int c16 = 0; c16 < 16; ++c16) {
                        const float4 e0 = __ldg(r0 + c16), e1 = __ldg(r1 + c16), e2 = __ldg(r2 + c16), e3 = __ldg(r3 + c16);
                        const float z0 = zr[c16 * 4 + 0], z1 = zr[c16 * 4 + 1], z2 = zr[c16 * 4 + 2], z3 = zr[c16 * 4 + 3];
                        M0 = __fmaf_rn(z0, e0.x, M0); M1 = __fmaf_rn(z0, e1.x, M1); M2 = __fmaf_rn(z0, e2.x, M2); M3 = __fmaf_rn(z0, e3.x, M3);
                        M0 = __fmaf_rn(z1, e0.y, M0); M1 = __fmaf_rn(z1, e1.y, M1); M2 = __fmaf_rn(z1, e2.y, M2); M3 = __fmaf_rn(z1, e3.y, M3);
                        M0 = __fmaf_rn(z2, e0.z, M0); M1 = __fmaf_rn(z2, e1.z, M1); M2 = __fmaf_rn(z2, e2.z, M2); M3 = __fmaf_rn(z2, e3.z, M3);
                        M0 = __fmaf_rn(z3, e0.w, M0); M1 = __fmaf_rn(z3, e1.w, M1); M2 = __fmaf_rn(z3, e2.w, M2); M3 = __fmaf_rn(z3, e3.w, M3);
                    }
                }
                consider(M0, kg + j0); consider(M1, kg + j1); consider(M2, kg + j2); consider(M3, kg + j3);
            };
            if (kflags & 1) {
                bk = 0;                                         // timing experiment only
            } else if (!slow_row) {
                // Both threads of the row walk BOTH compacted lists (their own, then the partner's, straight from
                // shared memory); each takes 4 of the 8 codes of a group.  No cap on the number of candidate groups: a cap of 4 per half sent about
                // one row in 10^5 to the exact scan of all K codes, which cost the whole kernel 50 us
                // (19 -> 70 us for one cfg2 batch in three).
                const int total = nc + pnc;
                for (int t = 0; t < total; ++t) {
                    const int g = __float_as_int(t < nc ? lists[t * 256 + et].y : lists[(t - nc) * 256 + (et ^ 128)].y);
                    if (g * 8 < p.K) rescore_group(g);
                }
            } else {
                // non-finite data or overflowing lists (e.g. many duplicated codes): every code, exactly
                for (int g = 0; g * 8 < p.K; ++g) rescore_group(g);
            }
            vq_mark(tr, it, 9);
            xbd[et] = bd;
            xbk[et] = bk;
            ptx::named_bar_sync(1 + q, 64);
            {
                const float od = xbd[et ^ 128];
                const int ok = xbk[et ^ 128];
                if (ok >= 0 && (bk < 0 || vq_better(od, ok, bd, bk))) { bd = od; bk = ok; }
            }

            vq_mark(tr, it, 10);
            // ---- gather e_idx, straight-through z_q (in place over the z tile), SSE, histogram ----
            const long long grow = tile * TM + row;
            {
                auto emit_bf16 = [&](const float *zh) {
                    // bf16 rows (128 B) written over the first atom of the z tile: this thread's 32 channels are the
                    // four 16-byte pieces 4h .. 4h+3 (both threads of the row hold z in registers by now)
#pragma unroll
                    for (int c8 = 0; c8 < 4; ++c8) {
                        float o[8];
#pragma unroll
                        for (int u = 0; u < 2; ++u) {
                            const int c16 = 2 * c8 + u;
                            float4 e4;
                            if (resident)
                                e4 = *reinterpret_cast<const float4 *>(code_ptr_smem(bk) + h * EATOM + ((c16 ^ (bk & 7)) << 4));
                            else
                                e4 = __ldg(reinterpret_cast<const float4 *>(p.E + (size_t)bk * DD + h * 32) + c16);
                            float4 df;
                            df.x = __fsub_rn(e4.x, zh[c16 * 4 + 0]); df.y = __fsub_rn(e4.y, zh[c16 * 4 + 1]);
                            df.z = __fsub_rn(e4.z, zh[c16 * 4 + 2]); df.w = __fsub_rn(e4.w, zh[c16 * 4 + 3]);
                            o[4 * u + 0] = __fadd_rn(zh[c16 * 4 + 0], df.x); o[4 * u + 1] = __fadd_rn(zh[c16 * 4 + 1], df.y);   // quantizer.py:67
                            o[4 * u + 2] = __fadd_rn(zh[c16 * 4 + 2], df.z); o[4 * u + 3] = __fadd_rn(zh[c16 * 4 + 3], df.w);
                            if (grow < p.N)
                                sse += (double)df.x * df.x + (double)df.y * df.y + (double)df.z * df.z + (double)df.w * df.w;
                        }
                        const __nv_bfloat162 h0 = __floats2bfloat162_rn(o[0], o[1]), h1 = __floats2bfloat162_rn(o[2], o[3]);
                        const __nv_bfloat162 h2 = __floats2bfloat162_rn(o[4], o[5]), h3 = __floats2bfloat162_rn(o[6], o[7]);
                        *reinterpret_cast<uint4 *>(zrow + (((4 * h + c8) ^ rsw) << 4)) =
                            make_uint4(*reinterpret_cast<const uint32_t *>(&h0), *reinterpret_cast<const uint32_t *>(&h1),
                                       *reinterpret_cast<const uint32_t *>(&h2), *reinterpret_cast<const uint32_t *>(&h3));
                    }
                };
                auto emit = [&](const float *zh) {
                    if (p.zq_bf16) { emit_bf16(zh); return; }
#pragma unroll
                    for (int c16 = 0; c16 < 8; ++c16) {
                        float4 e4;
                        if (resident)
                            e4 = *reinterpret_cast<const float4 *>(code_ptr_smem(bk) + h * EATOM + ((c16 ^ (bk & 7)) << 4));
                        else
                            e4 = __ldg(reinterpret_cast<const float4 *>(p.E + (size_t)bk * DD + h * 32) + c16);
                        float4 df, o;
                        df.x = __fsub_rn(e4.x, zh[c16 * 4 + 0]); df.y = __fsub_rn(e4.y, zh[c16 * 4 + 1]);
                        df.z = __fsub_rn(e4.z, zh[c16 * 4 + 2]); df.w = __fsub_rn(e4.w, zh[c16 * 4 + 3]);
                        o.x = __fadd_rn(zh[c16 * 4 + 0], df.x); o.y = __fadd_rn(zh[c16 * 4 + 1], df.y);   // quantizer.py:67
                        o.z = __fadd_rn(zh[c16 * 4 + 2], df.z); o.w = __fadd_rn(zh[c16 * 4 + 3], df.w);
                        *reinterpret_cast<float4 *>(zrow + h * ZATOM + ((c16 ^ rsw) << 4)) = o;
                        if (grow < p.N)
                            sse += (double)df.x * df.x + (double)df.y * df.y + (double)df.z * df.z + (double)df.w * df.w;
                    }
                };
                if (h == 0) {
                    emit(zr);
                    if (grow < p.N) {
                        p.idx[grow] = bk;
                        if (smem_hist) atomicAdd(&hist_s[bk], 1);
                        else atomicAdd(&p.hist[bk], 1);
                    }
                } else {
                    emit(zr + 32);
                }
            }
            ptx::fence_proxy_async();          // generic-proxy writes -> visible to the TMA store
            __syncwarp();
            if (lane == 0) ptx::mbar_arrive(bar(Q_FULL + zs));
            vq_mark(tr, it, 11);
        }

        if (tid == 128) { kmark(2); g_vqb_trace_vq[40] = (unsigned long long)it; }   // all tiles done
        // ---- CTA reduction of the SSE partial, histogram flush ----
#pragma unroll
        for (int off = 16; off >= 1; off >>= 1) sse += __shfl_xor_sync(0xffffffffu, sse, off);
        double *red = reinterpret_cast<double *>(sm + OFF_RED);
        if (lane == 0) red[warp - 4] = sse;
        ptx::named_bar_sync(5, 256);
        if (et == 0) {
            double s = 0.0;
            for (int w = 0; w < 8; ++w) s += red[w];
            p.partials[blockIdx.x] = s;
            if (blockIdx.x == 0) *p.pending = gridDim.x;
        }
        if (smem_hist)
            for (int k = et; k < p.K; k += 256) {
                const int cval = hist_s[k];
                if (cval) atomicAdd(&p.hist[k], cval);
            }
    }

    if (tid == 128) kmark(3);              // epilogue finished (histogram flushed)
    ptx::tc_fence_before();
    __syncthreads();
    if (tid == 128) kmark(4);              // every warp done (producer drained its TMA stores)
    if (tid == 128 && (kflags & 8) && blockIdx.x < 256) {
        unsigned long long t;
        unsigned sm_id;
        asm volatile("mov.u64 %0, %%globaltimer;" : "=l"(t));
        asm volatile("mov.u32 %0, %%smid;" : "=r"(sm_id));
        g_vqb_cta_t[2 * blockIdx.x + 1] = (t << 10) | (sm_id & 1023);      // end time with the SM id in the low bits
    }
    if (warp == 2) ptx::tmem_dealloc(tmem_base, 512);
}

// n < 0: the count is read from *pending (deferred reduction; 0 = sse is already final)
__global__ void vq_tc_sum_partials(const double *__restrict__ partials, int n, const unsigned *__restrict__ pending,
                                   double *__restrict__ out) {
    __shared__ double sh[256];
    if (n < 0) n = (int)*pending;
    if (n <= 0) return;
    double s = 0.0;
    for (int i = threadIdx.x; i < n; i += 256) s += partials[i];
    sh[threadIdx.x] = s;
    __syncthreads();
    for (int off = 128; off >= 1; off >>= 1) {
        if (threadIdx.x < off) sh[threadIdx.x] += sh[threadIdx.x + off];
        __syncthreads();
    }
    if (threadIdx.x == 0) *out = sh[0];
}

size_t align256(size_t x) { return (x + 255) / 256 * 256; }

}  // namespace

// shared with vq2.cu (same workspace layout)
void vq_tc_prep(const float *E, int K, int Kpad, float *bn, unsigned *scal, cudaStream_t s) {
    vq_tc_prep_kernel<<<(Kpad + 127) / 128, 128, 0, s>>>(E, K, Kpad, bn, scal);
}
void vq_tc_sum(const double *partials, int n, double *out, cudaStream_t s) {
    vq_tc_sum_partials<<<1, 256, 0, s>>>(partials, n, nullptr, out);
}

// workspace: [bn: Kpad floats][scal: 256 B][partials: 256 doubles]
size_t vq_tc_workspace_bytes(int K) {
    const size_t Kpad = (size_t)((K + CN - 1) / CN) * CN;
    return align256(Kpad * 4) + 256 + 256 * sizeof(double);
}

bool vq_tc_supported(long long N, int K, int D) {
    return D == DD && N >= 1 && N < (1LL << 31) && K >= 1 && K <= (1 << 20);
}

// the deferred-reduction marker lives behind both kernels' workspace layouts
size_t vq_ws_marker_offset(int K);

int launch_vq_reduce_sse(const void *ws, int K, double *sse, cudaStream_t s) {
    const int Kpad = (K + CN - 1) / CN * CN;
    const unsigned char *w = reinterpret_cast<const unsigned char *>(ws);
    const double *partials = reinterpret_cast<const double *>(w + align256((size_t)Kpad * 4) + 256);
    vq_tc_sum_partials<<<1, 256, 0, s>>>(partials, -1, reinterpret_cast<const unsigned *>(w + vq_ws_marker_offset(K)), sse);
    VQB_COUNT_LAUNCH(1);
    return vqb_cuda_status(cudaGetLastError());
}

// defer != 0: the SSE partials stay in the workspace (vqb_vq_reduce_sse_f32 sums them later, e.g. on a side stream)
int launch_vq_tc(const float *z, const float *E, long long N, int K, int D, long long *idx, void *zq, double *sse,
                 int *hist, void *ws, float *dbg, int defer, int zq_bf16, cudaStream_t s) {
    if (!vq_tc_supported(N, K, D)) return VQB_ERR_UNSUPPORTED;
    const int nchunks = (K + CN - 1) / CN;
    const int Kpad = nchunks * CN;
    unsigned char *w = reinterpret_cast<unsigned char *>(ws);
    float *bn = reinterpret_cast<float *>(w);
    unsigned *scal = reinterpret_cast<unsigned *>(w + align256((size_t)Kpad * 4));
    double *partials = reinterpret_cast<double *>(w + align256((size_t)Kpad * 4) + 256);

    CUtensorMap tmz, tme, tmq;
    int rc = vqb_encode_tmap_2d(&tmz, CU_TENSOR_MAP_DATA_TYPE_FLOAT32, z, DD, (uint64_t)N, DD * 4, 32, TM,
                                CU_TENSOR_MAP_SWIZZLE_128B);
    if (rc) return rc;
    rc = vqb_encode_tmap_2d(&tme, CU_TENSOR_MAP_DATA_TYPE_FLOAT32, E, DD, (uint64_t)K, DD * 4, 32, CN,
                            CU_TENSOR_MAP_SWIZZLE_128B);
    if (rc) return rc;
    rc = zq_bf16 ? vqb_encode_tmap_2d(&tmq, CU_TENSOR_MAP_DATA_TYPE_BFLOAT16, zq, DD, (uint64_t)N, DD * 2, 64, TM,
                                      CU_TENSOR_MAP_SWIZZLE_128B)
                 : vqb_encode_tmap_2d(&tmq, CU_TENSOR_MAP_DATA_TYPE_FLOAT32, zq, DD, (uint64_t)N, DD * 4, 32, TM,
                                      CU_TENSOR_MAP_SWIZZLE_128B);
    if (rc) return rc;

    cudaError_t e = cudaMemsetAsync(hist, 0, sizeof(int) * (size_t)K, s);
    if (e != cudaSuccess) return (int)e;
    int nlaunch = 2;
    if (nchunks > 2) {          // streamed codebook: norms / max / flag from a prep launch
        e = cudaMemsetAsync(scal, 0, 256, s);
        if (e != cudaSuccess) return (int)e;
        vq_tc_prep_kernel<<<(Kpad + 127) / 128, 128, 0, s>>>(E, K, Kpad, bn, scal);
        nlaunch = 3;
    }

    static bool attr_set = false;
    if (!attr_set) {
        e = cudaFuncSetAttribute(vq_tc_kernel<false>, cudaFuncAttributeMaxDynamicSharedMemorySize, SMEM_ALLOC);
        if (e != cudaSuccess) return (int)e;
        e = cudaFuncSetAttribute(vq_tc_kernel<true>, cudaFuncAttributeMaxDynamicSharedMemorySize, SMEM_ALLOC);
        if (e != cudaSuccess) return (int)e;
        attr_set = true;
    }
    int dev = 0, sms = 148;
    cudaGetDevice(&dev);
    cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, dev);
    const long long ntiles = (N + TM - 1) / TM;
    int grid = (int)(ntiles < sms ? ntiles : sms);
    if (grid > 256) grid = 256;
    if (const char *ge = vqb_getenv("VQB_TC_GRID")) { const int g = atoi(ge); if (g > 0 && g < grid) grid = g; }   // experiments
    VqTcParams p;
    p.E = E; p.bn = bn; p.scal = reinterpret_cast<const float *>(scal);
    p.N = N; p.K = K; p.nchunks = nchunks;
    p.idx = idx; p.partials = partials; p.hist = hist; p.dbg = dbg; p.zq_bf16 = zq_bf16;
    p.pending = reinterpret_cast<unsigned *>(w + vq_ws_marker_offset(K));
    {
        const char *fl = vqb_getenv("VQB_TC_FLAGS");
        p.flags = fl ? atoi(fl) : 0;
        const char *tt = vqb_getenv("VQB_TC_TRACE_TILE");
        if (tt && (p.flags & 8)) {
            const int v = atoi(tt);
            cudaMemcpyToSymbolAsync(g_vqb_trace_tile, &v, sizeof(int), 0, cudaMemcpyHostToDevice, s);
        }
    }
    if (dbg) vq_tc_kernel<true><<<grid, NTHREADS, SMEM_ALLOC, s>>>(tmz, tme, tmq, p);
    else if (cudaError_t le = vqb_launch(vq_tc_kernel<false>, dim3((unsigned)grid), dim3(NTHREADS), (size_t)SMEM_ALLOC, s, tmz, tme, tmq, p)) return (int)le;
    if (!defer) vq_tc_sum_partials<<<1, 256, 0, s>>>(partials, grid, nullptr, sse);
    VQB_COUNT_LAUNCH(defer ? nlaunch - 1 : nlaunch);
    return vqb_cuda_status(cudaGetLastError());
}
