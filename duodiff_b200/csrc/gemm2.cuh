// CTA-pair (cta_group::2) variant of the U-ViT GEMM: a cluster of two CTAs computes a 256 x 256 output tile.
//
// Why: with one CTA per 128x256 tile every k-block pulls 48 KB of operands into shared memory for 128x256x64 MACs
// (85 FLOP per byte of smem fill).  ncu on the round-1 kernel showed the mainloop pinned at the L2->SM delivery
// cap (~6300 B/clk chip-wide = 10.8 TB/s => ~910 TFLOP/s).  In a CTA pair each CTA stages its own 128x64 A tile and
// only HALF of the 256x64 W tile (the tensor cores of both SMs read both halves), i.e. 32 KB per k-block for the
// same MACs (128 FLOP/B).
//
// Warp roles (384 threads; the two single-thread control loops sit at high warp ids):
//   warps 0-7   epilogue: two warpgroups, each drains 64-column chunks TMEM -> regs -> packed-fp32 math -> smem ->
//               TMA store; the residual tile is TMA-loaded into the same staging buffer and overwritten in place
//   warp  8     TMA producer (both CTAs): A[rank] and W-half[rank]; completion bytes are credited to the LEADER's
//               full barrier (cp.async.bulk.tensor ... .cta_group::2)
//   warp  9     MMA issuer (leader CTA only): tcgen05.mma.cta_group::2 M=256 N=256 K=16; tcgen05.commit multicasts
//               "slot free" to both CTAs' empty barriers and "accumulator ready" to both tfull barriers
//   warps 10-11 aux: one tile ahead of the epilogue they merge the per-row LayerNorm partials into (rstd, mean*rstd)
//               and stage the tile's bias / colsum vectors in shared memory, so the epilogue never waits on L2
//
// Epilogues are those of gemm.cuh plus: GELU through tanh.approx (one MUFU per element instead of erff), FFMA2
// packed math, and an optional per-row partial LayerNorm statistic (mean, M2 over each 64-column chunk) written for
// the next consumer, which removes the standalone ln_stats pass.
#pragma once
#include "gemm.cuh"

namespace ddb {

template <int STAGES_, int NBUF_, bool LN_, int BN_ = 256, bool PROBE_ = false>
struct Gemm2Cfg {
    static constexpr int BM = 128;  // rows per CTA (256 per pair)
    static constexpr int BN = BN_;  // 256, or 128 for the N = 512 GEMMs (wave quantisation: 258 -> 516 tiles on 74 pairs)
    static constexpr int BK = 64;
    static constexpr int STAGES = STAGES_;
    static constexpr int NBUF = NBUF_;                 // staging buffers per epilogue warpgroup
    static constexpr int A_BYTES = BM * BK * 2;        // 16 KB
    static constexpr int B_BYTES = (BN / 2) * BK * 2;  // 16 / 8 KB: this CTA's half of W
    static constexpr int STAGE_BYTES = A_BYTES + B_BYTES;
    static constexpr int OUT_BUF_BYTES = 128 * 128;
    static constexpr int AUX_BYTES = LN_ ? 3072 : 1024;  // per buffer: [rowstats 1 KB][bias 1 KB][colsum 1 KB]
    static_assert(!PROBE_ || !LN_, "the probe epilogue is a residual epilogue");
    static constexpr int OFF_A = 0;
    static constexpr int OFF_B = OFF_A + STAGES * A_BYTES;
    static constexpr int OFF_OUT = OFF_B + STAGES * B_BYTES;
    static constexpr int OFF_AUX = OFF_OUT + 2 * NBUF * OUT_BUF_BYTES;
    static constexpr int OFF_BAR = OFF_AUX + 2 * AUX_BYTES;
    static constexpr int NUM_BARS = 2 * STAGES + 4 + 2 * NBUF + 4;
    static constexpr int SMEM_BYTES = OFF_BAR + NUM_BARS * 8 + 16;
    static_assert(SMEM_BYTES <= 232448, "exceeds 227 KB of shared memory");
};

template <int EPI, bool STATS, int STAGES, int NBUF, int BN_ = 256, bool PROBE = false>
__global__ void __cluster_dims__(2, 1, 1) __launch_bounds__(384, 1)
    gemm2_tcgen05_kernel(const __grid_constant__ GemmArgs a) {
    constexpr bool kLN = (EPI == EPI_LN || EPI == EPI_LN_GELU);
    static_assert(!PROBE || (STATS && !kLN && BN_ == 256), "the probe rides on the statistics epilogue");
    using Cfg = Gemm2Cfg<STAGES, NBUF, kLN, BN_, PROBE>;
    constexpr int BN = Cfg::BN;
    static_assert(BN == 256 || BN == 128, "tile width");

    extern __shared__ __align__(1024) uint8_t smem[];
    uint8_t* sA = smem + Cfg::OFF_A;
    uint8_t* sB = smem + Cfg::OFF_B;
    uint8_t* sOut = smem + Cfg::OFF_OUT;
    uint8_t* sAux = smem + Cfg::OFF_AUX;
    uint64_t* bars = reinterpret_cast<uint64_t*>(smem + Cfg::OFF_BAR);
    uint64_t* full_bar = bars;                  // [STAGES]  (used on the leader)
    uint64_t* empty_bar = full_bar + STAGES;    // [STAGES]  (per CTA, multicast-arrived by the leader's commits)
    uint64_t* tfull_bar = empty_bar + STAGES;   // [2]       (per CTA)
    uint64_t* tempty_bar = tfull_bar + 2;       // [2]       (used on the leader; 16 arrivals)
    uint64_t* res_bar = tempty_bar + 2;         // [2*NBUF]
    uint64_t* aux_full = res_bar + 2 * NBUF;    // [2]  64 arrivals (aux threads)
    uint64_t* aux_empty = aux_full + 2;         // [2]  8 arrivals (epilogue warps)
    uint32_t* tmem_holder = reinterpret_cast<uint32_t*>(aux_empty + 2);

    const int warp = threadIdx.x >> 5;
    const int lane = threadIdx.x & 31;
    const uint32_t rank = cluster_ctarank();
    const bool leader = (rank == 0);
    const int cluster_id = blockIdx.x >> 1;
    const int num_clusters = gridDim.x >> 1;

    // PDL: the set-up below (barriers, TMEM, descriptor prefetch, cluster sync) overlaps the predecessor's tail
    pdl_launch_dependents();
    const int nkb0 = a.K0 / Cfg::BK;
    const int nkb = nkb0 + a.K1 / Cfg::BK;

    constexpr int kProducerWarp = 8, kMmaWarp = 9, kAllocWarp = 10;
    if (threadIdx.x == 0 && (smem_u32(smem) & 1023u)) __trap();  // swizzled tiles need 1024-byte alignment
    if (warp == kProducerWarp && lane == 0) {
        tma_prefetch_desc(&a.tmA0);
        tma_prefetch_desc(BN == 256 ? &a.tmB2 : &a.tmB3);
        if (a.K1 > 0) tma_prefetch_desc(&a.tmA1);
        tma_prefetch_desc(&a.tmOut);
        if (EPI == EPI_RES) tma_prefetch_desc(&a.tmRes);
    }
    if (warp == kMmaWarp && lane == 0) {
        for (int i = 0; i < STAGES; ++i) {
            mbar_init(&full_bar[i], 1);
            mbar_init(&empty_bar[i], 1);
        }
        for (int i = 0; i < 2; ++i) {
            mbar_init(&tfull_bar[i], 1);
            mbar_init(&tempty_bar[i], 16);  // 8 epilogue warps x 2 CTAs
            mbar_init(&aux_full[i], 64);
            mbar_init(&aux_empty[i], 8);
        }
        for (int i = 0; i < 2 * NBUF; ++i) mbar_init(&res_bar[i], 1);
        fence_mbar_init();
    }
    if (warp == kAllocWarp) tmem_alloc_2cta<512>(tmem_holder);
    tc_fence_before();
    cluster_sync_all();  // barriers of both CTAs initialised before any remote arrive / multicast commit
    tc_fence_after();
    const uint32_t tmem_base = *tmem_holder;
    pdl_wait();  // everything below reads or overwrites buffers of earlier kernels
    // (the live row count of the early-exit compaction is one of them: read it only now, so that the set-up above still
    // overlaps the predecessor's tail -- reading it first cost every kernel of the compacted step ~1 us)
    const int M = a.m_dev ? ld_state(a.m_dev) : a.M;
    const int nblk_n = a.N / BN;
    const int nblk_m = (M + 255) / 256;
    const int num_tiles = nblk_m * nblk_n;

    if (warp == kProducerWarp) {
        // ===================================================================== TMA producer (both CTAs)
        if (lane == 0) {
            const uint64_t pol_first = l2_policy_evict_first();
            int stage = 0;
            uint32_t phase = 0;
            for (int tile = cluster_id; tile < num_tiles; tile += num_clusters) {
                const int m_blk = a.reverse ? nblk_m - 1 - tile / nblk_n : tile / nblk_n, n_blk = tile % nblk_n;
                const int row0 = m_blk * 256 + (int)rank * 128;
                const int wrow0 = n_blk * BN + (int)rank * (BN / 2);
                for (int kb = 0; kb < nkb; ++kb) {
                    mbar_wait(&empty_bar[stage], phase ^ 1);
                    const uint32_t fb = leader_smem_addr(&full_bar[stage]);
                    if (a.debug & 4) {
                        if (leader) mbar_arrive(&full_bar[stage]);
                    } else {
                        if (leader) mbar_expect_tx(&full_bar[stage], 2 * Cfg::STAGE_BYTES);
                        if (kb < nkb0) {
                            if (a.l2_hints & 1)
                                tma_load_2d_2cta_hint(sA + stage * Cfg::A_BYTES, &a.tmA0, fb, kb * Cfg::BK, row0, pol_first);
                            else
                                tma_load_2d_2cta(sA + stage * Cfg::A_BYTES, &a.tmA0, fb, kb * Cfg::BK, row0);
                        } else
                            tma_load_2d_2cta(sA + stage * Cfg::A_BYTES, &a.tmA1, fb, (kb - nkb0) * Cfg::BK, row0);
                        tma_load_2d_2cta(sB + stage * Cfg::B_BYTES, BN == 256 ? &a.tmB2 : &a.tmB3, fb, kb * Cfg::BK,
                                         wrow0);
                    }
                    if (++stage == STAGES) {
                        stage = 0;
                        phase ^= 1;
                    }
                }
            }
        }
    } else if (warp == kMmaWarp) {
        // ===================================================================== MMA issuer (leader CTA, one thread)
        if (leader && lane == 0) {
            constexpr uint32_t idesc = umma_idesc_bf16(256, BN);
            int stage = 0;
            uint32_t phase = 0;
            int it = 0;
            for (int tile = cluster_id; tile < num_tiles; tile += num_clusters, ++it) {
                const int as = it & 1;
                const uint32_t aph = (it >> 1) & 1;
                long long* tr = (a.trace && cluster_id == 0 && it < 16) ? a.trace + it * 16 : nullptr;  // bench-only
                if (tr) tr[0] = clock64();
                mbar_wait(&tempty_bar[as], aph ^ 1);
                tc_fence_after();
                if (tr) tr[1] = clock64();
                const uint32_t d_tmem = tmem_base + as * BN;
                for (int kb = 0; kb < nkb; ++kb) {
                    mbar_wait(&full_bar[stage], phase);
                    tc_fence_after();
                    if (tr && kb == 0) tr[2] = clock64();
                    const uint64_t a_desc = umma_desc_kmajor_sw128(smem_u32(sA + stage * Cfg::A_BYTES));
                    const uint64_t b_desc = umma_desc_kmajor_sw128(smem_u32(sB + stage * Cfg::B_BYTES));
                    if (!(a.debug & 2)) {
#pragma unroll
                        for (int k = 0; k < Cfg::BK / 16; ++k)
                            umma_f16_ss_2cta(d_tmem, a_desc + 2 * k, b_desc + 2 * k, idesc, (kb | k) != 0);
                    }
                    umma_commit_2cta(&empty_bar[stage], 0x3);
                    if (++stage == STAGES) {
                        stage = 0;
                        phase ^= 1;
                    }
                }
                umma_commit_2cta(&tfull_bar[as], 0x3);
                if (tr) tr[3] = clock64();
            }
        }
    } else if (warp >= kAllocWarp) {
        // ===================================================================== aux warps: per-tile vectors -> smem
        const int t = threadIdx.x - kAllocWarp * 32;  // 0..63
        int it = 0;
        for (int tile = cluster_id; tile < num_tiles; tile += num_clusters, ++it) {
            const int m_blk = a.reverse ? nblk_m - 1 - tile / nblk_n : tile / nblk_n, n_blk = tile % nblk_n;
            const int buf = it & 1;
            const uint32_t ph = (it >> 1) & 1;
            uint8_t* ab = sAux + buf * Cfg::AUX_BYTES;
            mbar_wait(&aux_empty[buf], ph ^ 1);
            const int col = n_blk * BN + (t * 4) % BN;  // BN = 128: threads 32-63 duplicate (harmless)
            if constexpr (kLN) {
                float2* srow = reinterpret_cast<float2*>(ab);
#pragma unroll
                for (int rr = 0; rr < 2; ++rr) {
                    const int r = t + rr * 64;
                    const int row = m_blk * 256 + (int)rank * 128 + r;
                    float rstd = 1.f, mr = 0.f;
                    if (row < M) ln_row_stats(a.stats, row, a.nparts, a.ln_dim, a.ln_eps, rstd, mr);
                    srow[r] = make_float2(rstd, mr);
                }
                *reinterpret_cast<float4*>(ab + 1024 + ((t * 16) % (BN * 4))) =
                    a.bias ? __ldg(reinterpret_cast<const float4*>(a.bias + col)) : make_float4(0.f, 0.f, 0.f, 0.f);
                *reinterpret_cast<float4*>(ab + 2048 + ((t * 16) % (BN * 4))) =
                    __ldg(reinterpret_cast<const float4*>(a.colsum + col));
            } else {
                *reinterpret_cast<float4*>(ab + ((t * 16) % (BN * 4))) =
                    a.bias ? __ldg(reinterpret_cast<const float4*>(a.bias + col)) : make_float4(0.f, 0.f, 0.f, 0.f);
            }
            mbar_arrive(&aux_full[buf]);  // release: the st.shared above are visible to the waiting epilogue warps
        }
    } else {
        // ===================================================================== epilogue (2 warpgroups per CTA)
        const int g = warp >> 2;
        const int quarter = warp & 3;
        const int et = threadIdx.x - g * 128;
        const int row_in_tile = quarter * 32 + lane;
        const uint32_t bar_id = 1 + g;
        uint8_t* my_bufs = sOut + g * NBUF * Cfg::OUT_BUF_BYTES;
        uint64_t* my_res = res_bar + g * NBUF;
        constexpr int CHUNKS_PER_WG = BN / 128;  // 64-column chunks per warpgroup and tile
        const bool traffic = !(a.debug & 8);
        uint32_t q = 0;

        // chunk sequence number s (0,1,2,...) of this warpgroup -> (tile, chunk-in-tile); residual prefetch
        auto prefetch_res = [&](uint32_t s) {
            const int tile = cluster_id + (int)(s / CHUNKS_PER_WG) * num_clusters;
            if (tile >= num_tiles) return;
            const int m_blk = a.reverse ? nblk_m - 1 - tile / nblk_n : tile / nblk_n, n_blk = tile % nblk_n;
            const int c = g + 2 * (int)(s % CHUNKS_PER_WG);
            uint64_t* rb = &my_res[s % NBUF];
            mbar_expect_tx(rb, Cfg::OUT_BUF_BYTES);
            tma_load_2d(my_bufs + (s % NBUF) * Cfg::OUT_BUF_BYTES, &a.tmRes, rb, n_blk * BN + c * 64,
                        (a.embed_mode ? 0 : m_blk * 256) + (int)rank * 128);
        };
        if (EPI == EPI_RES && et == 0 && traffic && !(a.debug & 1)) {
            for (uint32_t s = 0; s + 1 < NBUF; ++s) prefetch_res(s);
        }

        int it = 0;
        for (int tile = cluster_id; tile < num_tiles; tile += num_clusters, ++it) {
            const int m_blk = a.reverse ? nblk_m - 1 - tile / nblk_n : tile / nblk_n, n_blk = tile % nblk_n;
            const int as = it & 1;
            const uint32_t aph = (it >> 1) & 1;
            const int row0 = m_blk * 256 + (int)rank * 128;
            const int row = row0 + row_in_tile;
            const uint8_t* ab = sAux + as * Cfg::AUX_BYTES;
            const float* sbias = reinterpret_cast<const float*>(ab + (kLN ? 1024 : 0));
            const float* scs = reinterpret_cast<const float*>(ab + 2048);
            // PROBE: the 5-stage residual configuration has no shared memory left for a third per-tile vector; the probe
            // weights come through L1 (every thread of the CTA reads the same 32 bytes: one broadcast line per request)
            const float* spw = PROBE ? a.probe_w + n_blk * BN : nullptr;

            long long* tr = (a.trace && cluster_id == 0 && leader && threadIdx.x == 0 && it < 16) ? a.trace + it * 16
                                                                                                   : nullptr;
            if (tr) tr[4] = clock64();
            mbar_wait(&aux_full[as], aph);
            if (tr) tr[5] = clock64();
            float rstd = 1.f, mean_rstd = 0.f;
            if constexpr (kLN) {
                const float2 rs = reinterpret_cast<const float2*>(ab)[row_in_tile];
                rstd = rs.x, mean_rstd = rs.y;
            }
            mbar_wait(&tfull_bar[as], aph);
            tc_fence_after();
            if (tr) tr[6] = clock64();
            const uint32_t t_row = tmem_base + (uint32_t(quarter * 32) << 16) + as * BN;

            if (a.debug & 1) {
                uint32_t acc[32];
                tmem_ld_32x32b_x32(t_row, acc);
                tmem_ld_wait();
                tc_fence_before();
                __syncwarp();
                if (lane == 0) {
                    mbar_arrive_cluster(leader_smem_addr(&tempty_bar[as]));
                    mbar_arrive(&aux_empty[as]);
                }
                continue;
            }
#pragma unroll 1
            for (int cc = 0; cc < CHUNKS_PER_WG; ++cc) {
                const int c = g + 2 * cc;
                const int buf = q % NBUF;
                uint8_t* sbuf = my_bufs + buf * Cfg::OUT_BUF_BYTES;
                const int col0 = n_blk * BN + c * 64;

                uint32_t acc[2][32];
                tmem_ld_32x32b_x32(t_row + c * 64, acc[0]);
                tmem_ld_32x32b_x32(t_row + c * 64 + 32, acc[1]);

                if (EPI == EPI_RES && traffic) {
                    mbar_wait(&my_res[buf], (q / NBUF) & 1);  // residual chunk landed in sbuf
                } else {
                    // staging buffer `buf` was last used by chunk q-NBUF: its TMA store must have read it
                    if (et == 0) tma_store_wait_read<NBUF - 1>();
                    named_bar_sync(bar_id, 128);
                }
                tmem_ld_wait();
                if (tr && cc < 2) tr[7 + cc * 4] = clock64();
                if (cc == CHUNKS_PER_WG - 1) {
                    tc_fence_before();
                    __syncwarp();
                    if (lane == 0) mbar_arrive_cluster(leader_smem_addr(&tempty_bar[as]));
                }

                uint8_t* srow = sbuf + row_in_tile * 128;
                const f32x2 rstd2 = f2_splat(rstd), nmr2 = f2_splat(-mean_rstd);
                f32x2 s1 = f2_splat(0.f), s2 = f2_splat(0.f), nshift = f2_splat(0.f);
                // per-column vectors of the next 8 columns are fetched one iteration ahead (the issue order of a warp is
                // the program order: without this every j starts with an exposed shared-memory round trip)
                float4 nb0 = *reinterpret_cast<const float4*>(sbias + c * 64);
                float4 nb1 = *reinterpret_cast<const float4*>(sbias + c * 64 + 4);
                float4 nc0 = make_float4(0.f, 0.f, 0.f, 0.f), nc1 = nc0;
                if constexpr (kLN) {
                    nc0 = *reinterpret_cast<const float4*>(scs + c * 64);
                    nc1 = *reinterpret_cast<const float4*>(scs + c * 64 + 4);
                }
                if constexpr (PROBE) {  // the probe weights travel in the (unused) colsum registers
                    nc0 = __ldg(reinterpret_cast<const float4*>(spw + c * 64));
                    nc1 = __ldg(reinterpret_cast<const float4*>(spw + c * 64 + 4));
                }
                f32x2 pdot = f2_splat(0.f);
#pragma unroll
                for (int j = 0; j < 8; ++j) {
                    const int cl = c * 64 + j * 8;  // column inside the tile
                    f32x2 v[4];
#pragma unroll
                    for (int e = 0; e < 4; ++e)
                        v[e] = f2_pack_u(acc[j >> 2][(j & 3) * 8 + 2 * e], acc[j >> 2][(j & 3) * 8 + 2 * e + 1]);
                    const float4 b0 = nb0, b1 = nb1, c0 = nc0, c1 = nc1;
                    if (j < 7) {
                        nb0 = *reinterpret_cast<const float4*>(sbias + cl + 8);
                        nb1 = *reinterpret_cast<const float4*>(sbias + cl + 12);
                        if constexpr (kLN) {
                            nc0 = *reinterpret_cast<const float4*>(scs + cl + 8);
                            nc1 = *reinterpret_cast<const float4*>(scs + cl + 12);
                        }
                        if constexpr (PROBE) {
                            nc0 = __ldg(reinterpret_cast<const float4*>(spw + cl + 8));
                            nc1 = __ldg(reinterpret_cast<const float4*>(spw + cl + 12));
                        }
                    }
                    const f32x2 bb[4] = {f2_pack(b0.x, b0.y), f2_pack(b0.z, b0.w), f2_pack(b1.x, b1.y),
                                         f2_pack(b1.z, b1.w)};
                    if constexpr (kLN) {
                        const f32x2 cs[4] = {f2_pack(c0.x, c0.y), f2_pack(c0.z, c0.w), f2_pack(c1.x, c1.y),
                                             f2_pack(c1.z, c1.w)};
#pragma unroll
                        for (int e = 0; e < 4; ++e) v[e] = f2_fma(v[e], rstd2, f2_fma(nmr2, cs[e], bb[e]));
                    } else {
#pragma unroll
                        for (int e = 0; e < 4; ++e) v[e] = f2_add(v[e], bb[e]);
                    }
                    if constexpr (EPI == EPI_LN_GELU) {
                        // (no run-time switches inside this loop: a branch per j splits the unrolled body into basic
                        // blocks and the scheduler can no longer overlap the LDS / MUFU latency of one j with the math
                        // of the next -- with only two epilogue warps per scheduler that halves the issue rate)
#pragma unroll
                        for (int e = 0; e < 4; ++e) v[e] = gelu_fast2(v[e]);
                    }
                    uint4* sp = reinterpret_cast<uint4*>(srow + ((j ^ (row_in_tile & 7)) << 4));
                    if constexpr (EPI == EPI_RES) {
                        const uint4 r = *sp;
                        v[0] = f2_add(v[0], f2_pack_u(r.x << 16, r.x & 0xFFFF0000u));
                        v[1] = f2_add(v[1], f2_pack_u(r.y << 16, r.y & 0xFFFF0000u));
                        v[2] = f2_add(v[2], f2_pack_u(r.z << 16, r.z & 0xFFFF0000u));
                        v[3] = f2_add(v[3], f2_pack_u(r.w << 16, r.w & 0xFFFF0000u));
                    }
                    if constexpr (PROBE) {
                        pdot = f2_fma(v[0], f2_pack(c0.x, c0.y), pdot);
                        pdot = f2_fma(v[1], f2_pack(c0.z, c0.w), pdot);
                        pdot = f2_fma(v[2], f2_pack(c1.x, c1.y), pdot);
                        pdot = f2_fma(v[3], f2_pack(c1.z, c1.w), pdot);
                    }
                    if constexpr (STATS) {
                        // shifted single-pass statistics: s1 = sum(v - v0), s2 = sum((v - v0)^2); lanes = even/odd cols
                        if (j == 0) {
                            float lo, hi;
                            f2_unpack(v[0], lo, hi);
                            nshift = f2_splat(-lo);
                        }
#pragma unroll
                        for (int e = 0; e < 4; ++e) {
                            const f32x2 d = f2_add(v[e], nshift);
                            s1 = f2_add(s1, d);
                            s2 = f2_fma(d, d, s2);
                        }
                    }
                    uint4 o;
                    o.x = f2_to_bf16x2(v[0]);
                    o.y = f2_to_bf16x2(v[1]);
                    o.z = f2_to_bf16x2(v[2]);
                    o.w = f2_to_bf16x2(v[3]);
                    *sp = o;
                }
                if constexpr (STATS) {
                    // partial LayerNorm statistic (mean, M2) of this 64-column chunk for the next consumer
                    float s1a, s1b, s2a, s2b, ns, ns_;
                    f2_unpack(s1, s1a, s1b);
                    f2_unpack(s2, s2a, s2b);
                    f2_unpack(nshift, ns, ns_);
                    const float t1 = s1a + s1b, t2 = s2a + s2b;
                    const float dm = t1 * (1.f / 64.f);
                    if (row < M) {
                        const size_t orow = a.embed_mode
                                                ? (size_t)m_blk * a.tok_L + a.tok_extras + (int)rank * 128 + row_in_tile
                                                : (size_t)row;
                        a.stats_out[orow * (a.N >> 6) + (col0 >> 6)] = make_float2(dm - ns, fmaf(-t1, dm, t2));
                        if constexpr (PROBE) {
                            float pa, pb;
                            f2_unpack(pdot, pa, pb);
                            a.probe_out[orow * (a.N >> 6) + (col0 >> 6)] = pa + pb;
                        }
                    }
                }
                if (cc == CHUNKS_PER_WG - 1) {
                    __syncwarp();
                    if (lane == 0) mbar_arrive(&aux_empty[as]);  // this warp is done with the tile's smem vectors
                }
                if (tr && cc < 2) tr[8 + cc * 4] = clock64();
                fence_proxy_async_smem();
                named_bar_sync(bar_id, 128);
                if (tr && cc < 2) tr[9 + cc * 4] = clock64();
                if (et == 0 && traffic) {
                    if (a.embed_mode)
                        tma_store_3d(&a.tmOut, sbuf, col0, (int)rank * 128, m_blk);  // sample m_blk, patches rank*128..
                    else if (a.l2_hints & 2)
                        tma_store_2d_hint(&a.tmOut, sbuf, col0, row0, l2_policy_evict_last());
                    else
                        tma_store_2d(&a.tmOut, sbuf, col0, row0);
                    tma_store_commit();
                    if constexpr (EPI == EPI_RES) {
                        // buffer (q + NBUF - 1) % NBUF was last used by chunk q-1: wait until its store has read it,
                        // then prefetch the residual of chunk q + NBUF - 1 into it
                        tma_store_wait_read<1>();
                        prefetch_res(q + NBUF - 1);
                    }
                }
                if (tr && cc < 2) tr[10 + cc * 4] = clock64();
                ++q;
            }
        }
        if (et == 0) tma_store_wait_all<0>();
    }

    __syncwarp();  // warps 8/9 ran single-lane role loops: reconverge before the .aligned cluster barrier
    tc_fence_before();
    cluster_sync_all();  // peer may still multicast into / read from this CTA until both are done
    if (warp == kAllocWarp) {
        tc_fence_after();
        tmem_dealloc_2cta<512>(tmem_base);
    }
}

}  // namespace ddb
