// EXPERIMENTAL (compiled only with -DDDB_EXPERIMENTAL): fp32-FMA token assembly used together with the single-CTA GEMM
// variant (gemm_variant = 1).  The model path uses patch_gather_kernel + the CTA-pair GEMM + token_extras_kernel.
#pragma once
#include "../elementwise.cuh"

namespace ddb {

// =====================================================================================================
// Token assembly: patch embed (models/uvit.py:221-225) + time token (:95-115, :352-360) + label (:361-364)
// + pos_embed (:365).  One CTA per (sample, patch row); the CTA with patch row 0 also writes the extras.
// x_img [B,C,H,W] fp32 -> tokens [B*L, D] bf16.   Requires W/p == 16 patches per row (true for every config).
// =====================================================================================================
// PD = patch_dim (compile-time so the K loop unrolls and the W loads pipeline).  Each thread owns pairs of embedding
// dims (e, e+1) for all 16 tokens of the patch row; the math is packed fp32x2 over token pairs.  The kernel also
// writes the LayerNorm row statistics (mean, M2) of the bf16-rounded tokens for the first block's norm1, so no
// separate statistics pass follows (shifted single-pass sums; shift = the token's dim-0 value).
template <int PD>
__global__ void __launch_bounds__(256, 2) embed_tokens_kernel(
    const float* __restrict__ x_img, const float* __restrict__ tsteps, const long long* __restrict__ y,
    const float* __restrict__ Wt /*[pd,D]*/, const float* __restrict__ pe_bias, const float* __restrict__ pos /*[L,D]*/,
    const float* __restrict__ label_emb /*[classes,D] or null*/, __nv_bfloat16* __restrict__ tokens,
    float2* __restrict__ stats /*[B*L] (mean, M2) or null*/, int C, int H, int W, int P, int D, int L, int extras,
    int normalize_t) {
    __shared__ __align__(16) float patch[PD * EMB_TOK];  // patchT[k][16 tokens]
    __shared__ float shift[EMB_TOK];
    __shared__ float red[2][8][EMB_TOK + 2];
    const int Hp = H / P;
    const int b = blockIdx.x / Hp, hh = blockIdx.x % Hp;
    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    pdl_launch_dependents();
    pdl_wait();  // x_img / tsteps come from the previous step's kernels
    // gather: k = (c*P + p1)*P + p2 ; token ww ; pixel (c, hh*P+p1, ww*P+p2)
    {
        // C*P*W = PD*16 elements: all global loads are issued before the first shared-memory store
        constexpr int NLD = (PD * EMB_TOK + 255) / 256;
        float gv[NLD];
#pragma unroll
        for (int j = 0; j < NLD; ++j) {
            const int i = threadIdx.x + j * 256;
            const int col = i % W, rp = i / W;  // rp = c*P + p1
            const int c = rp / P, p1 = rp % P;
            gv[j] = (i < PD * EMB_TOK) ? __ldg(x_img + (((size_t)b * C + c) * H + hh * P + p1) * W + col) : 0.f;
        }
#pragma unroll
        for (int j = 0; j < NLD; ++j) {
            const int i = threadIdx.x + j * 256;
            const int col = i % W, rp = i / W;
            const int ww = col / P, p2 = col % P;
            if (i < PD * EMB_TOK) patch[(rp * P + p2) * EMB_TOK + ww] = gv[j];
        }
    }
    __syncthreads();
    float s1[EMB_TOK], s2[EMB_TOK];
#pragma unroll
    for (int t = 0; t < EMB_TOK; ++t) s1[t] = s2[t] = 0.f;
    bool first = true;
    for (int e2 = threadIdx.x; e2 < D / 2; e2 += blockDim.x) {  // uniform trip count per warp; D/2 is a multiple of 128
        const int e = e2 * 2;
        f32x2 a0[EMB_TOK / 2], a1[EMB_TOK / 2];  // dim e / dim e+1, token pairs
        const f32x2 b0 = f2_splat(pe_bias[e]), b1 = f2_splat(pe_bias[e + 1]);
#pragma unroll
        for (int t = 0; t < EMB_TOK / 2; ++t) a0[t] = b0, a1[t] = b1;
        // W rows are prefetched one batch of 4 k ahead of the FMAs that use them
        static_assert(PD % 4 == 0, "patch_dim must be a multiple of 4");
        float2 wn[4];
#pragma unroll
        for (int u = 0; u < 4; ++u) wn[u] = __ldg(reinterpret_cast<const float2*>(Wt + (size_t)u * D + e));
#pragma unroll 1
        for (int kb = 0; kb < PD; kb += 4) {
            float2 wc[4];
#pragma unroll
            for (int u = 0; u < 4; ++u) wc[u] = wn[u];
            if (kb + 4 < PD) {
#pragma unroll
                for (int u = 0; u < 4; ++u)
                    wn[u] = __ldg(reinterpret_cast<const float2*>(Wt + (size_t)(kb + 4 + u) * D + e));
            }
#pragma unroll
            for (int u = 0; u < 4; ++u) {
                const f32x2 w0 = f2_splat(wc[u].x), w1 = f2_splat(wc[u].y);
                const ulonglong2* pr = reinterpret_cast<const ulonglong2*>(patch + (kb + u) * EMB_TOK);
#pragma unroll
                for (int q = 0; q < 4; ++q) {
                    const ulonglong2 pv = pr[q];  // tokens 4q..4q+3 as two packed pairs
                    a0[2 * q] = f2_fma(pv.x, w0, a0[2 * q]);
                    a0[2 * q + 1] = f2_fma(pv.y, w0, a0[2 * q + 1]);
                    a1[2 * q] = f2_fma(pv.x, w1, a1[2 * q]);
                    a1[2 * q + 1] = f2_fma(pv.y, w1, a1[2 * q + 1]);
                }
            }
        }
        float v0[EMB_TOK], v1[EMB_TOK];
#pragma unroll
        for (int t = 0; t < EMB_TOK / 2; ++t) {
            f2_unpack(a0[t], v0[2 * t], v0[2 * t + 1]);
            f2_unpack(a1[t], v1[2 * t], v1[2 * t + 1]);
        }
#pragma unroll
        for (int t = 0; t < EMB_TOK; ++t) {
            const int l = extras + hh * EMB_TOK + t;
            const float2 pp = *reinterpret_cast<const float2*>(pos + (size_t)l * D + e);
            const uint32_t pk = pack_bf16(v0[t] + pp.x, v1[t] + pp.y);
            *reinterpret_cast<uint32_t*>(tokens + ((size_t)b * L + l) * D + e) = pk;
            v0[t] = bf16_lo(pk), v1[t] = bf16_hi(pk);  // statistics of what the consumer will read
        }
        if (stats) {
            if (first) {
                if (threadIdx.x == 0) {
#pragma unroll
                    for (int t = 0; t < EMB_TOK; ++t) shift[t] = v0[t];
                }
                __syncthreads();
                first = false;
            }
#pragma unroll
            for (int t = 0; t < EMB_TOK; ++t) {
                const float c = shift[t];
                const float d0 = v0[t] - c, d1 = v1[t] - c;
                s1[t] += d0 + d1;
                s2[t] = fmaf(d0, d0, fmaf(d1, d1, s2[t]));
            }
        }
    }
    if (stats) {
#pragma unroll
        for (int t = 0; t < EMB_TOK; ++t) {
            float a = s1[t], q = s2[t];
            for (int o = 16; o > 0; o >>= 1) {
                a += __shfl_xor_sync(0xffffffffu, a, o);
                q += __shfl_xor_sync(0xffffffffu, q, o);
            }
            if (lane == 0) red[0][warp][t] = a, red[1][warp][t] = q;
        }
        __syncthreads();
        if (threadIdx.x < EMB_TOK) {
            const int t = threadIdx.x;
            float a = 0.f, q = 0.f;
            for (int w = 0; w < 8; ++w) a += red[0][w][t], q += red[1][w][t];
            const float dm = a / (float)D;
            stats[(size_t)b * L + extras + hh * EMB_TOK + t] = make_float2(shift[t] + dm, fmaf(-a, dm, q));
        }
    }
    if (hh == 0) {
        // time token: [cos(tau f_i) | sin(tau f_i)], f_i = exp(-ln(1e4) i / half); label token in front of it
        const int half = D / 2;
        float tau = tsteps[b];
        if (normalize_t) tau = tau / 1000.f;
        __syncthreads();  // red[] / shift[] are reused below
        for (int row = 0; row < extras; ++row) {
            const bool is_time = (row == extras - 1);
            const float* src = is_time ? nullptr : label_emb + (size_t)y[b] * D;
            float a = 0.f, q = 0.f, c = 0.f;
            for (int e = threadIdx.x; e < D; e += blockDim.x) {
                float v;
                if (is_time) {
                    const int i = (e < half) ? e : e - half;
                    const float f = expf((-9.210340371976184f * (float)i) / (float)half);
                    const float arg = tau * f;
                    v = (e < half) ? cosf(arg) : sinf(arg);
                } else {
                    v = src[e];
                }
                const __nv_bfloat16 hv = __float2bfloat16_rn(v + pos[(size_t)row * D + e]);
                tokens[((size_t)b * L + row) * D + e] = hv;
                const float r = __bfloat162float(hv);
                if (e < (int)blockDim.x) c = r;  // shift = the thread's own first value
                const float d = r - c;
                a += d;
                q = fmaf(d, d, q);
            }
            if (stats) {
                // per-thread shifted sums (shift c_thread, n_thread values) -> merge as (mean, M2) partials (Chan)
                const int n_thr = (D - (int)threadIdx.x + (int)blockDim.x - 1) / (int)blockDim.x;
                float mean = c + a / (float)n_thr, m2 = fmaf(-a, a / (float)n_thr, q), cnt = (float)n_thr;
                for (int o = 16; o > 0; o >>= 1) {
                    const float mean_o = __shfl_xor_sync(0xffffffffu, mean, o);
                    const float m2_o = __shfl_xor_sync(0xffffffffu, m2, o);
                    const float cnt_o = __shfl_xor_sync(0xffffffffu, cnt, o);
                    const float tot = cnt + cnt_o, dlt = mean_o - mean;
                    m2 = m2 + m2_o + dlt * dlt * cnt * cnt_o / tot;
                    mean = mean + dlt * cnt_o / tot;
                    cnt = tot;
                }
                if (lane == 0) red[0][warp][row] = mean, red[1][warp][row] = m2, red[0][warp][EMB_TOK + row] = cnt;
                __syncthreads();
                if (threadIdx.x == 0) {
                    float M = red[0][0][row], Q = red[1][0][row], N = red[0][0][EMB_TOK + row];
                    for (int w = 1; w < 8; ++w) {
                        const float mo = red[0][w][row], qo = red[1][w][row], no = red[0][w][EMB_TOK + row];
                        const float tot = N + no, dlt = mo - M;
                        Q = Q + qo + dlt * dlt * N * no / tot;
                        M = M + dlt * no / tot;
                        N = tot;
                    }
                    stats[(size_t)b * L + row] = make_float2(M, Q);
                }
                __syncthreads();
            }
        }
    }
}

}  // namespace ddb
