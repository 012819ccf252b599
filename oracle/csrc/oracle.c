/*
 * oracle.c -- CPU restatement of the MishaLaskin/vqvae inference hot path.
 *
 * TEST INFRASTRUCTURE ONLY.  Nothing under vqvae_b200/ or models/ may import,
 * link or execute this file; only tests/, __graft_entry__.smoke() and bench.py's
 * cpu_baseline / --impl reference legs use it, and only as the checker.
 *
 * Every function cites the reference lines it restates (paths under
 * /root/reference).  The arithmetic lives in PyTorch (pinned torch==1.1.0 in
 * requirements.txt:10; container has torch 2.11.0), which is not vendored, so the
 * published definitions of Conv2d / ConvTranspose2d / argmin are restated here and
 * the restatement is pinned against outputs of the unmodified reference run in the
 * authoring container (oracle/make_golden.py -> tests/golden/).
 *
 * Canonical fp32 arithmetic of the VQ step (what "bit-exact" means for the CUDA
 * path; see DESIGN.md "VQ arithmetic contract"):
 *   A_i   = sum_d fl(z_id * z_id)        left-to-right fp32 adds of rounded squares
 *   B_k   = sum_d fl(e_kd * e_kd)        same
 *   M_ik  = fma chain over d = 0..D-1    (one accumulator per (i,k), like a GEMM
 *                                          micro-kernel / cuBLAS SGEMM thread)
 *   d_ik  = fl( fl(A_i + B_k) - fl(2 * M_ik) )         quantizer.py:49-51
 *   idx_i = first k attaining the minimum; a NaN distance wins (torch.argmin)
 */
#include <math.h>
#include <stdint.h>
#include <string.h>

#if defined(__GNUC__)
#pragma STDC FP_CONTRACT OFF
#endif

/* ---- VectorQuantizer.forward, models/quantizer.py:45-76 ------------------ */

static float sumsq_f32(const float *v, int D) {
    float s = 0.0f;
    for (int d = 0; d < D; ++d) {
        float sq = v[d] * v[d]; /* quantizer.py:49  z_flattened ** 2 (rounded) */
        s = s + sq;             /* torch.sum(dim=1), canonical left-to-right   */
    }
    return s;
}

/* argmin ordering: quantizer.py:54 torch.argmin -- lowest index among equal
 * minima, and a NaN compares as the minimum (first NaN wins).               */
static int better(float dn, float dbest) {
    if (isnan(dbest)) return 0;
    if (isnan(dn)) return 1;
    return dn < dbest;
}

/*
 * z      : (N, D) rows = z.permute(0,2,3,1).view(-1, e_dim)   quantizer.py:45-46
 * E      : (K, D) embedding.weight                             quantizer.py:26
 * idx    : (N) int64 min_encoding_indices                      quantizer.py:54
 * zq     : (N, D) straight-through value fl(z + fl(e - z))     quantizer.py:60,67
 * sse    : sum over all elements of fl(e - z)^2 in double (the numerator of
 *          both mean() terms of quantizer.py:63-64)
 * hist   : (K) int32 code counts = column sums of the one-hot  quantizer.py:55-57,70
 * dmin   : optional (N) winning distance (may be NULL)
 */
void oracle_vq_forward_f32(const float *z, const float *E, int64_t N, int K, int D,
                           int64_t *idx, float *zq, double *sse, int32_t *hist,
                           float *dmin) {
    float *B = (float *)__builtin_malloc(sizeof(float) * (size_t)K);
    for (int k = 0; k < K; ++k) B[k] = sumsq_f32(E + (size_t)k * D, D);
    memset(hist, 0, sizeof(int32_t) * (size_t)K);
    double acc_sse = 0.0;
#pragma omp parallel for reduction(+ : acc_sse) schedule(static)
    for (int64_t i = 0; i < N; ++i) {
        const float *zi = z + (size_t)i * D;
        float A = sumsq_f32(zi, D);
        float best = 0.0f;
        int bestk = -1;
        for (int k = 0; k < K; ++k) {
            const float *ek = E + (size_t)k * D;
            float m = 0.0f;
            for (int d = 0; d < D; ++d) m = fmaf(zi[d], ek[d], m); /* quantizer.py:51 */
            float ab = A + B[k];
            float two_m = 2.0f * m;
            float dist = ab - two_m;
            if (bestk < 0 || better(dist, best)) {
                best = dist;
                bestk = k;
            }
        }
        idx[i] = bestk;
        if (dmin) dmin[i] = best;
        const float *eb = E + (size_t)bestk * D;
        double row_sse = 0.0;
        for (int d = 0; d < D; ++d) {
            float diff = eb[d] - zi[d];            /* (z_q - z), quantizer.py:63-64,67 */
            zq[(size_t)i * D + d] = zi[d] + diff;  /* z + (z_q - z).detach()            */
            row_sse += (double)diff * (double)diff;
        }
        acc_sse += row_sse;
    }
    for (int64_t i = 0; i < N; ++i) hist[idx[i]] += 1;
    *sse = acc_sse;
    __builtin_free(B);
}

/* loss and perplexity scalars, quantizer.py:63-64 and :70-71, fp32 like torch */
void oracle_vq_finish_f32(double sse, const int32_t *hist, int64_t N, int K, int D,
                          float beta, float *loss, float *perplexity) {
    float mse = (float)(sse / ((double)N * (double)D));
    *loss = mse + beta * mse;
    float ent = 0.0f;
    for (int k = 0; k < K; ++k) {
        float p = (float)hist[k] / (float)N; /* torch.mean(min_encodings, dim=0) */
        ent = ent + p * logf(p + 1e-10f);
    }
    *perplexity = expf(-ent);
}

/* ---- nn.Conv2d (NCHW), encoder.py:29-36, residual.py:20-24, vqvae.py:16-17 --
 * x (B,Cin,H,W); w (Cout,Cin,kh,kw); bias (Cout) or NULL; y (B,Cout,OH,OW)
 * OH = (H + 2*pad - kh)/stride + 1.  Accumulates in double and rounds once: the
 * most accurate fp32 answer; the CUDA path is compared with a tolerance.       */
void oracle_conv2d_f32(const float *x, const float *w, const float *bias, float *y,
                       int B, int Cin, int H, int W, int Cout, int kh, int kw,
                       int stride, int pad) {
    int OH = (H + 2 * pad - kh) / stride + 1;
    int OW = (W + 2 * pad - kw) / stride + 1;
#pragma omp parallel for collapse(2) schedule(static)
    for (int b = 0; b < B; ++b)
        for (int co = 0; co < Cout; ++co)
            for (int oh = 0; oh < OH; ++oh)
                for (int ow = 0; ow < OW; ++ow) {
                    double acc = bias ? (double)bias[co] : 0.0;
                    for (int ci = 0; ci < Cin; ++ci)
                        for (int r = 0; r < kh; ++r) {
                            int ih = oh * stride - pad + r;
                            if (ih < 0 || ih >= H) continue;
                            for (int s = 0; s < kw; ++s) {
                                int iw = ow * stride - pad + s;
                                if (iw < 0 || iw >= W) continue;
                                acc += (double)x[(((size_t)b * Cin + ci) * H + ih) * W + iw] *
                                       (double)w[(((size_t)co * Cin + ci) * kh + r) * kw + s];
                            }
                        }
                    y[(((size_t)b * Cout + co) * OH + oh) * OW + ow] = (float)acc;
                }
}

/* ---- nn.ConvTranspose2d (NCHW), decoder.py:28-35 ---------------------------
 * x (B,Cin,H,W); w (Cin,Cout,kh,kw); y (B,Cout,OH,OW), OH=(H-1)*stride-2*pad+kh.
 * Written in the scatter form of the definition (every input pixel adds
 * x*w into the outputs it touches), independent of the gather / sub-pixel phase
 * decomposition the CUDA kernels use.                                          */
void oracle_conv_transpose2d_f32(const float *x, const float *w, const float *bias,
                                 float *y, int B, int Cin, int H, int W, int Cout,
                                 int kh, int kw, int stride, int pad) {
    int OH = (H - 1) * stride - 2 * pad + kh;
    int OW = (W - 1) * stride - 2 * pad + kw;
    size_t plane = (size_t)OH * OW;
#pragma omp parallel for collapse(2) schedule(static)
    for (int b = 0; b < B; ++b)
        for (int co = 0; co < Cout; ++co) {
            double *acc = (double *)__builtin_malloc(sizeof(double) * plane);
            for (size_t i = 0; i < plane; ++i) acc[i] = bias ? (double)bias[co] : 0.0;
            for (int ci = 0; ci < Cin; ++ci)
                for (int ih = 0; ih < H; ++ih)
                    for (int iw = 0; iw < W; ++iw) {
                        double xv = (double)x[(((size_t)b * Cin + ci) * H + ih) * W + iw];
                        for (int r = 0; r < kh; ++r) {
                            int oh = ih * stride - pad + r;
                            if (oh < 0 || oh >= OH) continue;
                            for (int s = 0; s < kw; ++s) {
                                int ow = iw * stride - pad + s;
                                if (ow < 0 || ow >= OW) continue;
                                acc[(size_t)oh * OW + ow] +=
                                    xv * (double)w[(((size_t)ci * Cout + co) * kh + r) * kw + s];
                            }
                        }
                    }
            float *yo = y + ((size_t)b * Cout + co) * plane;
            for (size_t i = 0; i < plane; ++i) yo[i] = (float)acc[i];
            __builtin_free(acc);
        }
}

/* nn.ReLU / F.relu, encoder.py:31,34; residual.py:19,22,50; decoder.py:33 */
void oracle_relu_f32(float *x, int64_t n) {
    for (int64_t i = 0; i < n; ++i) x[i] = x[i] > 0.0f ? x[i] : 0.0f;
}

/* x + res_block(x) after the in-place ReLU has already replaced x by relu(x)
 * (residual.py:19,27-29): out = a + b elementwise.                 */
void oracle_add_f32(const float *a, const float *b, float *out, int64_t n) {
    for (int64_t i = 0; i < n; ++i) out[i] = a[i] + b[i];
}
