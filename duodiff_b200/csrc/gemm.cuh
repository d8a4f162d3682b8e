// Persistent, warp-specialised tcgen05/TMEM GEMM for the U-ViT linears (sm_100a).
//
//   out[M,N] = epilogue( [A0 | A1][M, K0+K1] (bf16, K-major)  x  W[N, K0+K1]^T (bf16, K-major) )
//
// Replaces every nn.Linear on the reference hot path (models/uvit.py:87-90 fc1/fc2, :158 qkv, :166 proj,
// :205 skip_linear over cat([x, skip]), :378 decoder_pred) together with the LayerNorm in front of it
// (:206-207, :377), the bias, the exact-erf GELU (:88), the residual add (:206-207) and -- for the decoder --
// the `extras` slice + un-patchify (:379-381).
//
// Roles (384 threads, 1 CTA / SM):
//   warp 0      TMA producer: A (two K sources for the long skip) and W tiles -> 128B-swizzled smem ring
//   warp 1      MMA issuer:   tcgen05.mma cta_group::1 kind::f16, M=128 x N=BN x K=16, fp32 accum in TMEM
//   warp 2      TMEM allocator (2 accumulator stages x BN columns)
//   warps 4-11  epilogue: two warpgroups, each drains 64-column chunks TMEM -> regs -> math -> smem -> TMA store;
//               the residual tile is TMA-loaded into the same staging buffer and overwritten in place.
//
// LayerNorm is folded: gamma goes into W' = W*diag(gamma), beta into bias' = bias + W*beta, and the per-row
// normalisation commutes with the contraction:  LN(x) W^T = rstd * (x W'^T - mean * colsum(W')) + bias'.
// The epilogue applies that fix-up from per-row (mean, M2) partials, so the A operand is the raw residual stream.
#pragma once
#include "ptx.cuh"

namespace ddb {

enum GemmEpilogue : int {
    EPI_BIAS = 0,     // out = acc + bias
    EPI_LN = 1,       // out = rstd*(acc - mean*colsum) + bias
    EPI_LN_GELU = 2,  // out = gelu(rstd*(acc - mean*colsum) + bias)
    EPI_RES = 3,      // out = residual + acc + bias
    EPI_DECODE = 4,   // LN fix-up + bias, drop `extras` rows, scatter un-patchified fp32 to [B,C,H,W]
};

struct GemmArgs {
    CUtensorMap tmA0;   // [M, K0] bf16, box {64, 128}
    CUtensorMap tmA1;   // [M, K1] bf16 (second K source; unused when K1 == 0)
    CUtensorMap tmB;    // [N, K0+K1] bf16, box {64, BN}
    CUtensorMap tmB2;   // same tensor, box {64, 128}: one CTA's half of the W tile in the CTA-pair kernel
    CUtensorMap tmB3;   // same tensor, box {64, 64}: half of a 128-wide W tile (CTA-pair kernel, N = 512 GEMMs)
    CUtensorMap tmOut;  // [M, N] bf16, box {64, 128}
    CUtensorMap tmRes;  // [M, N] bf16 residual, box {64, 128}
    CUtensorMap tmOut2; // [M, N] bf16, box {32, 128}, SWIZZLE_64B (CTA-pair kernel: 32-column sub-chunks)
    CUtensorMap tmRes2; // [M, N] bf16 residual, box {32, 128}, SWIZZLE_64B
    int M, N, K0, K1;
    const float* bias;    // [N] or nullptr
    const float* colsum;  // [N]           (LN epilogues)
    const float2* stats;  // [M, nparts] partial (mean, M2) over ln_dim/nparts features (LN epilogues)
    int nparts;
    int ln_dim;
    float ln_eps;
    const int* m_dev;  // optional: live row count read from device memory (early-exit compaction)
    int debug;          // bench-only knobs (ddb_set_option "gemm_debug"): 1 = epilogue drains TMEM but skips math/stores,
                        // 2 = MMA issue skipped (barrier traffic only), 4 = no TMA operand loads
    float2* stats_out;  // optional [M, N/64]: per-row (mean, M2) of every 64-column output chunk (gemm2 only)
    // Early-exit probe folded into the producer (gemm2, PROBE instantiation): probe_out[row, N/64] = partial dot products
    // of the output row (fp32, before the bf16 rounding) with probe_w over every 64-column chunk (models/early_exit.py:34-37)
    const float* probe_w;  // [N]
    float* probe_out;      // [M, N/64]
    long long* trace;   // bench-only: cluster 0 / leader records clock64() per tile ([tile][16])
    // Patch-embed mode (CTA-pair kernel): GEMM row r = (sample b = r / 256, patch l = r % 256).  The output goes to
    // token row b*tok_L + tok_extras + l through a 3-D tensor map (tmOut: [B][256][N] view of the token buffer), the
    // "residual" is the positional embedding of patch l (tmRes: [256, N], the same rows for every sample).
    int embed_mode, tok_L, tok_extras;
    // Row blocks are walked from the last to the first (CTA-pair kernel).  Consecutive kernels of a transformer block
    // alternate the direction, so a consumer starts with the rows its producer wrote LAST -- the part of a tensor larger
    // than what the L2 can hold (qkv 101 MB, MLP hidden 135 MB of 126 MB) that is still resident.
    int reverse;
    // L2 eviction priority of the CTA-pair kernel's TMA traffic: bit 0 = A loads evict_first (the operand is dead after
    // this GEMM), bit 1 = output stores evict_last (the next kernel re-reads them)
    int l2_hints;
    // EPI_DECODE scatter geometry
    float* img;  // [B, C, H, W] fp32
    int L, extras, C, P, Wp, H, W, patch_dim;
    // EPI_DECODE, grouped mode (early-exit compaction): one launch decodes every sample that left the batch, each with
    // the output head of ITS exit layer.  A = the leavers' hidden rows kept per original slot, tmA0 = 3-D map
    // [D, L, B] (box {64, 128, 1}), tmB = 3-D map over the stacked head weights [D, 64, depth] (box {64, 64, 1}); tiles =
    // (sample, 128-row block of its L tokens); grp_layer[b] = exit layer of sample b, samples with grp_layer[b] >=
    // grp_depth are skipped (they never left: the full model's head handles them).  bias / colsum are stacked
    // [depth][64]; stats is indexed by b * L + token.
    const int* grp_layer;
    int grp_depth, grp_B;
    // EPI_DECODE after early-exit compaction: compact sample j is un-patchified into image slot dec_slot[j] (its original
    // position in the batch), so the stayers' and the leavers' decodes fill one image buffer
    const int* dec_slot;
};

template <int BN>
struct GemmCfg {
    static constexpr int BM = 128;
    static constexpr int BK = 64;
    static constexpr int STAGES = (BN == 256) ? 3 : 6;
    static constexpr int A_BYTES = BM * BK * 2;
    static constexpr int B_BYTES = BN * BK * 2;
    static constexpr int STAGE_BYTES = A_BYTES + B_BYTES;
    static constexpr int OUT_BUF_BYTES = 128 * 128;  // 128 rows x 64 bf16
    static constexpr int NUM_OUT_BUFS = 4;           // 2 per epilogue warpgroup
    static constexpr int TMEM_COLS = (2 * BN <= 128) ? 128 : (2 * BN <= 256 ? 256 : 512);
    static constexpr int OFF_A = 0;
    static constexpr int OFF_B = OFF_A + STAGES * A_BYTES;
    static constexpr int OFF_OUT = OFF_B + STAGES * B_BYTES;
    static constexpr int OFF_BAR = OFF_OUT + NUM_OUT_BUFS * OUT_BUF_BYTES;
    static constexpr int SMEM_BYTES = OFF_BAR + 256 + 1024;  // + alignment slack
};

__device__ __forceinline__ void ln_row_stats(const float2* __restrict__ stats, int row, int nparts, int ln_dim,
                                             float eps, float& rstd, float& mean_rstd) {
    // merge equal-count (mean, M2) partials (Chan et al.)
    float mean = 0.f;
    for (int i = 0; i < nparts; ++i) mean += __ldg(&stats[(size_t)row * nparts + i]).x;
    mean /= (float)nparts;
    float m2 = 0.f;
    const float cnt = (float)(ln_dim / nparts);
    for (int i = 0; i < nparts; ++i) {
        float2 p = __ldg(&stats[(size_t)row * nparts + i]);
        float d = p.x - mean;
        m2 += p.y + cnt * d * d;
    }
    rstd = rsqrtf(m2 / (float)ln_dim + eps);
    mean_rstd = mean * rstd;
}

// EPI_DECODE store of one token: out[b, c, hh*P+p1, ww*P+p2] = LN-fixed acc[(p1*P+p2)*C + c]; warpgroup g takes the
// patch rows p1 in [g*P/2, (g+1)*P/2) and writes each (c, p1) run of P pixels as one vector.
struct GemmArgs;
template <int C, int P>
__device__ __forceinline__ void decode_store(const uint32_t (&acc)[2][32], const GemmArgs& a, float rstd,
                                             float mean_rstd, int b, int hh, int ww, int g, int voff);

template <int BN, int EPI>
__global__ void __launch_bounds__(384, 1) gemm_tcgen05_kernel(const __grid_constant__ GemmArgs a) {
    using Cfg = GemmCfg<BN>;
    constexpr int STAGES = Cfg::STAGES;
    constexpr bool kTmaStore = (EPI != EPI_DECODE);
    constexpr bool kLN = (EPI == EPI_LN || EPI == EPI_LN_GELU || EPI == EPI_DECODE);

    extern __shared__ uint8_t smem_raw[];
    uint8_t* smem = reinterpret_cast<uint8_t*>((reinterpret_cast<uintptr_t>(smem_raw) + 1023) & ~uintptr_t(1023));
    uint8_t* sA = smem + Cfg::OFF_A;
    uint8_t* sB = smem + Cfg::OFF_B;
    uint8_t* sOut = smem + Cfg::OFF_OUT;
    uint64_t* bars = reinterpret_cast<uint64_t*>(smem + Cfg::OFF_BAR);
    uint64_t* full_bar = bars;                  // [STAGES]
    uint64_t* empty_bar = bars + STAGES;        // [STAGES]
    uint64_t* tfull_bar = bars + 2 * STAGES;    // [2]
    uint64_t* tempty_bar = tfull_bar + 2;       // [2]
    uint64_t* res_bar = tempty_bar + 2;         // [4]  (warpgroup g, buffer b) -> g*2+b
    uint32_t* tmem_holder = reinterpret_cast<uint32_t*>(res_bar + 4);

    const int warp = threadIdx.x >> 5;
    const int lane = threadIdx.x & 31;

    pdl_launch_dependents();
    // grouped decode: tiles = (sample, 128-row block of its tokens)
    const bool kGrouped = (EPI == EPI_DECODE) && a.grp_layer != nullptr;
    const int grp_tps = (a.L + Cfg::BM - 1) / Cfg::BM;
    const int nkb0 = a.K0 / Cfg::BK;
    const int nkb = nkb0 + a.K1 / Cfg::BK;

    if (warp == 0 && lane == 0) {
        tma_prefetch_desc(&a.tmA0);
        tma_prefetch_desc(&a.tmB);
        if (a.K1 > 0) tma_prefetch_desc(&a.tmA1);
        if (kTmaStore) tma_prefetch_desc(&a.tmOut);
        if (EPI == EPI_RES) tma_prefetch_desc(&a.tmRes);
    }
    if (warp == 1 && lane == 0) {
        for (int i = 0; i < STAGES; ++i) {
            mbar_init(&full_bar[i], 1);
            mbar_init(&empty_bar[i], 1);
        }
        for (int i = 0; i < 2; ++i) {
            mbar_init(&tfull_bar[i], 1);
            mbar_init(&tempty_bar[i], 8);  // one elected lane per epilogue warp
        }
        for (int i = 0; i < 4; ++i) mbar_init(&res_bar[i], 1);
        fence_mbar_init();
    }
    if (warp == 2) tmem_alloc<Cfg::TMEM_COLS>(tmem_holder);
    tc_fence_before();
    __syncthreads();
    tc_fence_after();
    const uint32_t tmem_base = *tmem_holder;
    pdl_wait();
    // the live row count (early-exit compaction) is an earlier kernel's output: read after the wait, set-up before it
    const int M = a.m_dev ? ld_state(a.m_dev) : a.M;
    const int nblk_n = a.N / BN;
    const int nblk_m = (M + Cfg::BM - 1) / Cfg::BM;
    const int num_tiles = kGrouped ? a.grp_B * grp_tps : nblk_m * nblk_n;

    if (warp == 0) {
        // ===================================================================== TMA producer
        if (lane == 0) {
            int stage = 0;
            uint32_t phase = 0;
            for (int tile = blockIdx.x; tile < num_tiles; tile += gridDim.x) {
                const int m_blk = tile / nblk_n, n_blk = tile % nblk_n;
                int grp_b = 0, grp_l = 0;
                if (kGrouped) {
                    grp_b = tile / grp_tps;
                    grp_l = ld_state(a.grp_layer + grp_b);
                    if (grp_l >= a.grp_depth) continue;  // every role skips the same tiles
                }
                for (int kb = 0; kb < nkb; ++kb) {
                    mbar_wait(&empty_bar[stage], phase ^ 1);
                    mbar_expect_tx(&full_bar[stage], Cfg::STAGE_BYTES);
                    if (kGrouped) {
                        tma_load_3d(sA + stage * Cfg::A_BYTES, &a.tmA0, &full_bar[stage], kb * Cfg::BK,
                                    (tile % grp_tps) * Cfg::BM, grp_b);
                        tma_load_3d(sB + stage * Cfg::B_BYTES, &a.tmB, &full_bar[stage], kb * Cfg::BK, 0, grp_l);
                        if (++stage == STAGES) {
                            stage = 0;
                            phase ^= 1;
                        }
                        continue;
                    }
                    if (kb < nkb0)
                        tma_load_2d(sA + stage * Cfg::A_BYTES, &a.tmA0, &full_bar[stage], kb * Cfg::BK,
                                    m_blk * Cfg::BM);
                    else
                        tma_load_2d(sA + stage * Cfg::A_BYTES, &a.tmA1, &full_bar[stage], (kb - nkb0) * Cfg::BK,
                                    m_blk * Cfg::BM);
                    tma_load_2d(sB + stage * Cfg::B_BYTES, &a.tmB, &full_bar[stage], kb * Cfg::BK, n_blk * BN);
                    if (++stage == STAGES) {
                        stage = 0;
                        phase ^= 1;
                    }
                }
            }
        }
    } else if (warp == 1) {
        // ===================================================================== MMA issuer (one thread)
        if (lane == 0) {
            constexpr uint32_t idesc = umma_idesc_bf16(Cfg::BM, BN);
            int stage = 0;
            uint32_t phase = 0;
            int it = 0;
            for (int tile = blockIdx.x; tile < num_tiles; tile += gridDim.x) {
                if (kGrouped && ld_state(a.grp_layer + tile / grp_tps) >= a.grp_depth) continue;
                const int as = it & 1;
                const uint32_t aph = (it >> 1) & 1;
                ++it;
                mbar_wait(&tempty_bar[as], aph ^ 1);  // epilogue has drained this accumulator stage
                tc_fence_after();
                const uint32_t d_tmem = tmem_base + as * BN;
                for (int kb = 0; kb < nkb; ++kb) {
                    mbar_wait(&full_bar[stage], phase);
                    tc_fence_after();
                    const uint64_t a_desc = umma_desc_kmajor_sw128(smem_u32(sA + stage * Cfg::A_BYTES));
                    const uint64_t b_desc = umma_desc_kmajor_sw128(smem_u32(sB + stage * Cfg::B_BYTES));
#pragma unroll
                    for (int k = 0; k < Cfg::BK / 16; ++k) {
                        // +32 B per K=16 step inside the 128 B swizzle row (encoded >> 4)
                        umma_f16_ss(d_tmem, a_desc + 2 * k, b_desc + 2 * k, idesc, (kb | k) != 0);
                    }
                    umma_commit(&empty_bar[stage]);  // smem slot reusable once these MMAs retire
                    if (++stage == STAGES) {
                        stage = 0;
                        phase ^= 1;
                    }
                }
                umma_commit(&tfull_bar[as]);  // accumulator complete -> epilogue
            }
        }
    } else if (warp >= 4) {
        // ===================================================================== epilogue (2 warpgroups)
        const int g = (warp - 4) >> 2;         // warpgroup 0/1
        const int quarter = warp & 3;          // TMEM lane quarter this warp may read
        const int et = threadIdx.x - 128 - g * 128;  // 0..127 == row inside the tile
        const int row_in_tile = quarter * 32 + lane;
        const uint32_t bar_id = 1 + g;
        uint8_t* my_bufs = sOut + g * 2 * Cfg::OUT_BUF_BYTES;
        constexpr int CHUNKS_PER_WG = (BN / 64 + 1) / 2;  // BN=256 -> 2 ; BN=64 -> 1 (wg1 idle for chunk math)
        uint32_t q = 0;  // running chunk counter for this warpgroup (TMA-store modes)

        // first residual prefetch
        if (EPI == EPI_RES && et == 0) {
            int tile = blockIdx.x;
            if (tile < num_tiles && g < BN / 64) {
                const int m_blk = tile / nblk_n, n_blk = tile % nblk_n;
                mbar_expect_tx(&res_bar[g * 2 + 0], Cfg::OUT_BUF_BYTES);
                tma_load_2d(my_bufs, &a.tmRes, &res_bar[g * 2 + 0], n_blk * BN + g * 64, m_blk * Cfg::BM);
            }
        }

        int it = 0;
        for (int tile = blockIdx.x; tile < num_tiles; tile += gridDim.x) {
            const int m_blk = tile / nblk_n, n_blk = tile % nblk_n;
            int grp_b = 0, grp_l = 0, grp_tok = 0;
            if (kGrouped) {
                grp_b = tile / grp_tps;
                grp_l = ld_state(a.grp_layer + grp_b);
                if (grp_l >= a.grp_depth) continue;
                grp_tok = (tile % grp_tps) * Cfg::BM + row_in_tile;
            }
            const int as = it & 1;
            const uint32_t aph = (it >> 1) & 1;
            ++it;
            const int row = kGrouped ? grp_b * a.L + grp_tok : m_blk * Cfg::BM + row_in_tile;
            const bool row_ok = kGrouped ? grp_tok < a.L : row < M;
            float rstd = 1.f, mean_rstd = 0.f;
            if (kLN && row_ok) ln_row_stats(a.stats, row, a.nparts, a.ln_dim, a.ln_eps, rstd, mean_rstd);

            mbar_wait(&tfull_bar[as], aph);
            tc_fence_after();
            const uint32_t t_row = tmem_base + (uint32_t(quarter * 32) << 16) + as * BN;

            if constexpr (kTmaStore) {
#pragma unroll 1
                for (int cc = 0; cc < CHUNKS_PER_WG; ++cc) {
                    const int c = g + 2 * cc;  // 64-column chunk inside the tile
                    const int buf = q & 1;
                    uint8_t* sbuf = my_bufs + buf * Cfg::OUT_BUF_BYTES;
                    const int col0 = n_blk * BN + c * 64;

                    uint32_t acc[2][32];
                    tmem_ld_32x32b_x32(t_row + c * 64, acc[0]);
                    tmem_ld_32x32b_x32(t_row + c * 64 + 32, acc[1]);

                    if constexpr (EPI == EPI_RES) {
                        mbar_wait(&res_bar[g * 2 + buf], (q >> 1) & 1);  // residual chunk landed in sbuf
                    } else {
                        // staging buffer `buf` was last used by chunk q-2: its TMA store must have read it
                        if (et == 0) tma_store_wait_read<1>();
                        named_bar_sync(bar_id, 128);
                    }
                    tmem_ld_wait();
                    if (cc == CHUNKS_PER_WG - 1) {
                        // all TMEM reads of this tile by this warp are done -> release accumulator stage early
                        tc_fence_before();
                        __syncwarp();
                        if (lane == 0) mbar_arrive(&tempty_bar[as]);
                    }

                    uint8_t* srow = sbuf + row_in_tile * 128;
#pragma unroll
                    for (int j = 0; j < 8; ++j) {  // 8 x (8 columns = 16 B)
                        const int cbase = col0 + j * 8;
                        float v[8];
#pragma unroll
                        for (int e = 0; e < 8; ++e) v[e] = __uint_as_float(acc[j >> 2][(j & 3) * 8 + e]);
                        float bb[8];
                        if (a.bias) {
                            const float4 b0 = __ldg(reinterpret_cast<const float4*>(a.bias + cbase));
                            const float4 b1 = __ldg(reinterpret_cast<const float4*>(a.bias + cbase + 4));
                            bb[0] = b0.x, bb[1] = b0.y, bb[2] = b0.z, bb[3] = b0.w;
                            bb[4] = b1.x, bb[5] = b1.y, bb[6] = b1.z, bb[7] = b1.w;
                        } else {
#pragma unroll
                            for (int e = 0; e < 8; ++e) bb[e] = 0.f;
                        }
                        if constexpr (kLN) {
                            const float4 c0 = __ldg(reinterpret_cast<const float4*>(a.colsum + cbase));
                            const float4 c1 = __ldg(reinterpret_cast<const float4*>(a.colsum + cbase + 4));
                            const float cs[8] = {c0.x, c0.y, c0.z, c0.w, c1.x, c1.y, c1.z, c1.w};
#pragma unroll
                            for (int e = 0; e < 8; ++e) v[e] = fmaf(v[e], rstd, fmaf(-mean_rstd, cs[e], bb[e]));
                        } else {
#pragma unroll
                            for (int e = 0; e < 8; ++e) v[e] += bb[e];
                        }
                        if constexpr (EPI == EPI_LN_GELU) {
#pragma unroll
                            for (int e = 0; e < 8; ++e) v[e] = gelu_erf(v[e]);
                        }
                        uint4* sp = reinterpret_cast<uint4*>(srow + ((j ^ (row_in_tile & 7)) << 4));
                        if constexpr (EPI == EPI_RES) {
                            const uint4 r = *sp;
                            v[0] += bf16_lo(r.x), v[1] += bf16_hi(r.x);
                            v[2] += bf16_lo(r.y), v[3] += bf16_hi(r.y);
                            v[4] += bf16_lo(r.z), v[5] += bf16_hi(r.z);
                            v[6] += bf16_lo(r.w), v[7] += bf16_hi(r.w);
                        }
                        uint4 o;
                        o.x = pack_bf16(v[0], v[1]);
                        o.y = pack_bf16(v[2], v[3]);
                        o.z = pack_bf16(v[4], v[5]);
                        o.w = pack_bf16(v[6], v[7]);
                        *sp = o;
                    }
                    fence_proxy_async_smem();
                    named_bar_sync(bar_id, 128);
                    if (et == 0) {
                        tma_store_2d(&a.tmOut, sbuf, col0, m_blk * Cfg::BM);
                        tma_store_commit();
                        if constexpr (EPI == EPI_RES) {
                            // prefetch the residual of chunk q+1 into the other buffer (last used by chunk q-1)
                            int ncc = cc + 1, ntile = tile;
                            if (ncc == CHUNKS_PER_WG) {
                                ncc = 0;
                                ntile = tile + gridDim.x;
                            }
                            if (ntile < num_tiles) {
                                tma_store_wait_read<1>();
                                const int nm = ntile / nblk_n, nn = ntile % nblk_n;
                                uint64_t* rb = &res_bar[g * 2 + (buf ^ 1)];
                                mbar_expect_tx(rb, Cfg::OUT_BUF_BYTES);
                                tma_load_2d(my_bufs + (buf ^ 1) * Cfg::OUT_BUF_BYTES, &a.tmRes, rb,
                                            nn * BN + (g + 2 * ncc) * 64, nm * Cfg::BM);
                            }
                        }
                    }
                    ++q;
                }
            } else {
                // ---------------------------------------------------------------- EPI_DECODE (BN == 64)
                // every thread reads its token's whole (zero-padded) patch vector; warpgroup g stores the patch rows
                // p1 in its half as P-wide vectors (un-patchify: feature (p1, p2, c), channel innermost)
                uint32_t acc[2][32];
                tmem_ld_32x32b_x32(t_row, acc[0]);
                tmem_ld_32x32b_x32(t_row + 32, acc[1]);
                tmem_ld_wait();
                tc_fence_before();
                __syncwarp();
                if (lane == 0) mbar_arrive(&tempty_bar[as]);
                if (row_ok) {
                    const int b = kGrouped ? grp_b : (a.dec_slot ? ld_state(a.dec_slot + row / a.L) : row / a.L);
                    const int l = kGrouped ? grp_tok : row % a.L;
                    const int voff = grp_l * 64;  // grouped: this sample's head in the stacked bias / colsum vectors
                    if (l >= a.extras) {
                        const int n = l - a.extras;
                        const int hh = n / a.Wp, ww = n % a.Wp;
                        if (a.C == 3 && a.P == 4)
                            decode_store<3, 4>(acc, a, rstd, mean_rstd, b, hh, ww, g, voff);
                        else if (a.C == 3 && a.P == 2)
                            decode_store<3, 2>(acc, a, rstd, mean_rstd, b, hh, ww, g, voff);
                        else if (a.C == 4 && a.P == 2)
                            decode_store<4, 2>(acc, a, rstd, mean_rstd, b, hh, ww, g, voff);
                        else
                            __trap();  // plan_decode_geometry() rejects other (in_chans, patch_size) pairs
                    }
                }
            }
        }
        if (kTmaStore && et == 0) tma_store_wait_all<0>();
    }

    tc_fence_before();
    __syncthreads();
    if (warp == 2) {
        tc_fence_after();
        tmem_dealloc<Cfg::TMEM_COLS>(tmem_base);
    }
}

template <int C, int P>
__device__ __forceinline__ void decode_store(const uint32_t (&acc)[2][32], const GemmArgs& a, float rstd,
                                             float mean_rstd, int b, int hh, int ww, int g, int voff) {
    static_assert(P == 2 || P == 4, "patch size");
#pragma unroll
    for (int pr = 0; pr < P / 2; ++pr) {
        const int p1 = g * (P / 2) + pr;
#pragma unroll
        for (int ch = 0; ch < C; ++ch) {
            float v[P];
#pragma unroll
            for (int p2 = 0; p2 < P; ++p2) {
                // p1 is only known at run time through g: select between the two compile-time candidates
                const int j0 = ((pr)*P + p2) * C + ch, j1 = ((P / 2 + pr) * P + p2) * C + ch;
                const float raw = __uint_as_float(g ? acc[j1 >> 5][j1 & 31] : acc[j0 >> 5][j0 & 31]);
                const int j = g ? j1 : j0;
                v[p2] = fmaf(raw, rstd,
                             fmaf(-mean_rstd, __ldg(a.colsum + voff + j), a.bias ? __ldg(a.bias + voff + j) : 0.f));
            }
            float* dst = a.img + (((size_t)b * C + ch) * a.H + hh * P + p1) * a.W + ww * P;
            if constexpr (P == 4)
                *reinterpret_cast<float4*>(dst) = make_float4(v[0], v[1], v[2], v[3]);
            else
                *reinterpret_cast<float2*>(dst) = make_float2(v[0], v[1]);
        }
    }
}

}  // namespace ddb
