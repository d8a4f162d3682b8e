// CTA-pair (cta_group::2) variant of the U-ViT GEMM: a cluster of two CTAs computes a 256 x 256 output tile.
//
// Why: with one CTA per 128x256 tile every k-block pulls 48 KB of operands into shared memory for 128x256x64 MACs
// (85 FLOP per byte of smem fill).  ncu on the round-1 kernel showed the mainloop pinned at the L2->SM delivery
// cap (~6300 B/clk chip-wide = 10.8 TB/s => ~910 TFLOP/s).  In a CTA pair each CTA stages its own 128x64 A tile and
// only HALF of the 256x64 W tile (the tensor cores of both SMs read both halves), i.e. 32 KB per k-block for the
// same MACs (128 FLOP/B), which lifts the cap above the tensor peak under the power limit.
//
// Protocol (per cluster; CTA rank 0 = leader):
//   both CTAs : TMA producer loads A[rank] and W-half[rank]; completion bytes are credited to the LEADER's full
//               barrier (cp.async.bulk.tensor ... .cta_group::2)
//   leader    : one thread issues tcgen05.mma.cta_group::2 (M=256,N=256,K=16); tcgen05.commit multicasts the
//               "slot free" arrive to both CTAs' empty barriers and "accumulator ready" to both tfull barriers
//   both CTAs : 8 epilogue warps drain their own 128 TMEM lanes; "accumulator drained" arrives go to the leader
//
// Epilogues are those of gemm.cuh plus: GELU through tanh.approx (one MUFU per element instead of erff), and an
// optional per-row partial LayerNorm statistic (mean, M2 over each 64-column chunk) written for the next consumer,
// which removes the standalone ln_stats pass.
#pragma once
#include "gemm.cuh"

namespace ddb {

struct Gemm2Cfg {
    static constexpr int BM = 128;  // rows per CTA (256 per pair)
    static constexpr int BN = 256;
    static constexpr int BK = 64;
    static constexpr int STAGES = 5;
    static constexpr int A_BYTES = BM * BK * 2;        // 16 KB
    static constexpr int B_BYTES = (BN / 2) * BK * 2;  // 16 KB: this CTA's half of W
    static constexpr int STAGE_BYTES = A_BYTES + B_BYTES;
    static constexpr int OUT_BUF_BYTES = 128 * 128;
    static constexpr int NUM_OUT_BUFS = 4;
    static constexpr int OFF_A = 0;
    static constexpr int OFF_B = OFF_A + STAGES * A_BYTES;
    static constexpr int OFF_OUT = OFF_B + STAGES * B_BYTES;
    static constexpr int OFF_BAR = OFF_OUT + NUM_OUT_BUFS * OUT_BUF_BYTES;
    static constexpr int SMEM_BYTES = OFF_BAR + 256 + 1024;
};

template <int EPI, bool STATS>
__global__ void __cluster_dims__(2, 1, 1) __launch_bounds__(384, 1)
    gemm2_tcgen05_kernel(const __grid_constant__ GemmArgs a) {
    using Cfg = Gemm2Cfg;
    constexpr int STAGES = Cfg::STAGES;
    constexpr int BN = Cfg::BN;
    constexpr bool kLN = (EPI == EPI_LN || EPI == EPI_LN_GELU);

    extern __shared__ uint8_t smem_raw[];
    uint8_t* smem = reinterpret_cast<uint8_t*>((reinterpret_cast<uintptr_t>(smem_raw) + 1023) & ~uintptr_t(1023));
    uint8_t* sA = smem + Cfg::OFF_A;
    uint8_t* sB = smem + Cfg::OFF_B;
    uint8_t* sOut = smem + Cfg::OFF_OUT;
    uint64_t* bars = reinterpret_cast<uint64_t*>(smem + Cfg::OFF_BAR);
    uint64_t* full_bar = bars;                // [STAGES]  (used on the leader)
    uint64_t* empty_bar = bars + STAGES;      // [STAGES]  (per CTA, multicast-arrived by the leader's commits)
    uint64_t* tfull_bar = bars + 2 * STAGES;  // [2]       (per CTA)
    uint64_t* tempty_bar = tfull_bar + 2;     // [2]       (used on the leader; 16 arrivals)
    uint64_t* res_bar = tempty_bar + 2;       // [4]
    uint32_t* tmem_holder = reinterpret_cast<uint32_t*>(res_bar + 4);

    const int warp = threadIdx.x >> 5;
    const int lane = threadIdx.x & 31;
    const uint32_t rank = cluster_ctarank();
    const bool leader = (rank == 0);
    const int cluster_id = blockIdx.x >> 1;
    const int num_clusters = gridDim.x >> 1;

    const int M = a.m_dev ? *a.m_dev : a.M;
    const int nblk_n = a.N / BN;
    const int nblk_m = (M + 255) / 256;
    const int num_tiles = nblk_m * nblk_n;
    const int nkb0 = a.K0 / Cfg::BK;
    const int nkb = nkb0 + a.K1 / Cfg::BK;

    if (warp == 0 && lane == 0) {
        tma_prefetch_desc(&a.tmA0);
        tma_prefetch_desc(&a.tmB2);
        if (a.K1 > 0) tma_prefetch_desc(&a.tmA1);
        tma_prefetch_desc(&a.tmOut);
        if (EPI == EPI_RES) tma_prefetch_desc(&a.tmRes);
    }
    if (warp == 1 && lane == 0) {
        for (int i = 0; i < STAGES; ++i) {
            mbar_init(&full_bar[i], 1);
            mbar_init(&empty_bar[i], 1);
        }
        for (int i = 0; i < 2; ++i) {
            mbar_init(&tfull_bar[i], 1);
            mbar_init(&tempty_bar[i], 16);  // 8 epilogue warps x 2 CTAs
        }
        for (int i = 0; i < 4; ++i) mbar_init(&res_bar[i], 1);
        fence_mbar_init();
    }
    if (warp == 2) tmem_alloc_2cta<512>(tmem_holder);
    tc_fence_before();
    cluster_sync_all();  // barriers of both CTAs initialised before any remote arrive / multicast commit
    tc_fence_after();
    const uint32_t tmem_base = *tmem_holder;

    if (warp == 0) {
        // ===================================================================== TMA producer (both CTAs)
        if (lane == 0) {
            int stage = 0;
            uint32_t phase = 0;
            for (int tile = cluster_id; tile < num_tiles; tile += num_clusters) {
                const int m_blk = tile / nblk_n, n_blk = tile % nblk_n;
                const int row0 = m_blk * 256 + (int)rank * 128;
                const int wrow0 = n_blk * BN + (int)rank * 128;
                for (int kb = 0; kb < nkb; ++kb) {
                    mbar_wait(&empty_bar[stage], phase ^ 1);
                    const uint32_t fb = leader_smem_addr(&full_bar[stage]);
                    if (leader) mbar_expect_tx(&full_bar[stage], 2 * Cfg::STAGE_BYTES);
                    if (kb < nkb0)
                        tma_load_2d_2cta(sA + stage * Cfg::A_BYTES, &a.tmA0, fb, kb * Cfg::BK, row0);
                    else
                        tma_load_2d_2cta(sA + stage * Cfg::A_BYTES, &a.tmA1, fb, (kb - nkb0) * Cfg::BK, row0);
                    tma_load_2d_2cta(sB + stage * Cfg::B_BYTES, &a.tmB2, fb, kb * Cfg::BK, wrow0);
                    if (++stage == STAGES) {
                        stage = 0;
                        phase ^= 1;
                    }
                }
            }
        }
    } else if (warp == 1) {
        // ===================================================================== MMA issuer (leader CTA, one thread)
        if (leader && lane == 0) {
            constexpr uint32_t idesc = umma_idesc_bf16(256, BN);
            int stage = 0;
            uint32_t phase = 0;
            int it = 0;
            for (int tile = cluster_id; tile < num_tiles; tile += num_clusters, ++it) {
                const int as = it & 1;
                const uint32_t aph = (it >> 1) & 1;
                mbar_wait(&tempty_bar[as], aph ^ 1);
                tc_fence_after();
                const uint32_t d_tmem = tmem_base + as * BN;
                for (int kb = 0; kb < nkb; ++kb) {
                    mbar_wait(&full_bar[stage], phase);
                    tc_fence_after();
                    const uint64_t a_desc = umma_desc_kmajor_sw128(smem_u32(sA + stage * Cfg::A_BYTES));
                    const uint64_t b_desc = umma_desc_kmajor_sw128(smem_u32(sB + stage * Cfg::B_BYTES));
#pragma unroll
                    for (int k = 0; k < Cfg::BK / 16; ++k)
                        umma_f16_ss_2cta(d_tmem, a_desc + 2 * k, b_desc + 2 * k, idesc, (kb | k) != 0);
                    umma_commit_2cta(&empty_bar[stage], 0x3);
                    if (++stage == STAGES) {
                        stage = 0;
                        phase ^= 1;
                    }
                }
                umma_commit_2cta(&tfull_bar[as], 0x3);
            }
        }
    } else if (warp >= 4) {
        // ===================================================================== epilogue (2 warpgroups per CTA)
        const int g = (warp - 4) >> 2;
        const int quarter = warp & 3;
        const int et = threadIdx.x - 128 - g * 128;
        const int row_in_tile = quarter * 32 + lane;
        const uint32_t bar_id = 1 + g;
        uint8_t* my_bufs = sOut + g * 2 * Cfg::OUT_BUF_BYTES;
        constexpr int CHUNKS_PER_WG = 2;
        uint32_t q = 0;

        if (EPI == EPI_RES && et == 0) {
            const int tile = cluster_id;
            if (tile < num_tiles) {
                const int m_blk = tile / nblk_n, n_blk = tile % nblk_n;
                mbar_expect_tx(&res_bar[g * 2 + 0], Cfg::OUT_BUF_BYTES);
                tma_load_2d(my_bufs, &a.tmRes, &res_bar[g * 2 + 0], n_blk * BN + g * 64,
                            m_blk * 256 + (int)rank * 128);
            }
        }

        int it = 0;
        for (int tile = cluster_id; tile < num_tiles; tile += num_clusters, ++it) {
            const int m_blk = tile / nblk_n, n_blk = tile % nblk_n;
            const int as = it & 1;
            const uint32_t aph = (it >> 1) & 1;
            const int row0 = m_blk * 256 + (int)rank * 128;
            const int row = row0 + row_in_tile;
            float rstd = 1.f, mean_rstd = 0.f;
            if (kLN && row < M) ln_row_stats(a.stats, row, a.nparts, a.ln_dim, a.ln_eps, rstd, mean_rstd);

            mbar_wait(&tfull_bar[as], aph);
            tc_fence_after();
            const uint32_t t_row = tmem_base + (uint32_t(quarter * 32) << 16) + as * BN;

#pragma unroll 1
            for (int cc = 0; cc < CHUNKS_PER_WG; ++cc) {
                const int c = g + 2 * cc;
                const int buf = q & 1;
                uint8_t* sbuf = my_bufs + buf * Cfg::OUT_BUF_BYTES;
                const int col0 = n_blk * BN + c * 64;

                uint32_t acc[2][32];
                tmem_ld_32x32b_x32(t_row + c * 64, acc[0]);
                tmem_ld_32x32b_x32(t_row + c * 64 + 32, acc[1]);

                if constexpr (EPI == EPI_RES) {
                    mbar_wait(&res_bar[g * 2 + buf], (q >> 1) & 1);
                } else {
                    if (et == 0) tma_store_wait_read<1>();
                    named_bar_sync(bar_id, 128);
                }
                tmem_ld_wait();
                if (cc == CHUNKS_PER_WG - 1) {
                    tc_fence_before();
                    __syncwarp();
                    if (lane == 0) mbar_arrive_cluster(leader_smem_addr(&tempty_bar[as]));
                }

                uint8_t* srow = sbuf + row_in_tile * 128;
                // shifted single-pass statistics of this thread's 64 outputs: s1 = sum(v - v0), s2 = sum((v - v0)^2)
                float s1 = 0.f, s2 = 0.f, shift = 0.f;
#pragma unroll
                for (int j = 0; j < 8; ++j) {
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
                        for (int e = 0; e < 8; ++e) v[e] = gelu_fast(v[e]);
                    }
                    uint4* sp = reinterpret_cast<uint4*>(srow + ((j ^ (row_in_tile & 7)) << 4));
                    if constexpr (EPI == EPI_RES) {
                        const uint4 r = *sp;
                        v[0] += bf16_lo(r.x), v[1] += bf16_hi(r.x);
                        v[2] += bf16_lo(r.y), v[3] += bf16_hi(r.y);
                        v[4] += bf16_lo(r.z), v[5] += bf16_hi(r.z);
                        v[6] += bf16_lo(r.w), v[7] += bf16_hi(r.w);
                    }
                    if constexpr (STATS) {
                        if (j == 0) shift = v[0];
#pragma unroll
                        for (int e = 0; e < 8; ++e) {
                            const float d = v[e] - shift;
                            s1 += d;
                            s2 = fmaf(d, d, s2);
                        }
                    }
                    uint4 o;
                    o.x = pack_bf16(v[0], v[1]);
                    o.y = pack_bf16(v[2], v[3]);
                    o.z = pack_bf16(v[4], v[5]);
                    o.w = pack_bf16(v[6], v[7]);
                    *sp = o;
                }
                if constexpr (STATS) {
                    // partial LayerNorm statistic (mean, M2) of this 64-column chunk for the next consumer
                    const float dm = s1 * (1.f / 64.f);
                    if (row < M)
                        a.stats_out[(size_t)row * (a.N >> 6) + (col0 >> 6)] =
                            make_float2(shift + dm, fmaf(-s1, dm, s2));
                }
                fence_proxy_async_smem();
                named_bar_sync(bar_id, 128);
                if (et == 0) {
                    tma_store_2d(&a.tmOut, sbuf, col0, row0);
                    tma_store_commit();
                    if constexpr (EPI == EPI_RES) {
                        int ncc = cc + 1, ntile = tile;
                        if (ncc == CHUNKS_PER_WG) {
                            ncc = 0;
                            ntile = tile + num_clusters;
                        }
                        if (ntile < num_tiles) {
                            tma_store_wait_read<1>();
                            const int nm = ntile / nblk_n, nn = ntile % nblk_n;
                            uint64_t* rb = &res_bar[g * 2 + (buf ^ 1)];
                            mbar_expect_tx(rb, Cfg::OUT_BUF_BYTES);
                            tma_load_2d(my_bufs + (buf ^ 1) * Cfg::OUT_BUF_BYTES, &a.tmRes, rb,
                                        nn * BN + (g + 2 * ncc) * 64, nm * 256 + (int)rank * 128);
                        }
                    }
                }
                ++q;
            }
        }
        if (et == 0) tma_store_wait_all<0>();
    }

    __syncwarp();  // warps 0/1 ran single-lane role loops: reconverge before the .aligned cluster barrier
    tc_fence_before();
    cluster_sync_all();  // peer may still multicast into / read from this CTA until both are done
    if (warp == 2) {
        tc_fence_after();
        tmem_dealloc_2cta<512>(tmem_base);
    }
}

}  // namespace ddb
