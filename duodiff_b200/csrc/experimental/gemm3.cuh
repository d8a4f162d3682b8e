// CTA-pair GEMM with the A operand RESIDENT IN TENSOR MEMORY (K <= 512: qkv, fc1, proj, patch embed).
//
// Why (profiles/r01_gemm_trace.txt): in the SS form the 256x256x16 tcgen05.mma of gemm2.cuh is paced
// by its shared-memory operand reads (A 4 KB + W-half 4 KB per CTA and MMA => ~169 clk instead of the 128-clk tensor
// rate), and the TMA fill of the same bytes plus the epilogue staging traffic compete for the same shared memory.
// Here the 128 x K bf16 A panel of a CTA is copied once per 256-row block into TMEM columns [256, 256 + K/2)
// (tcgen05.cp, shared memory -> TMEM) and every MMA of every N tile of that block reads A from TMEM (TS form); only W
// streams through shared memory.  The other 256 TMEM columns hold ONE accumulator; the 16-warp epilogue loads its
// 64 columns per warp into registers first and releases the accumulator before doing any math, so the next tile's
// MMAs start ~500 clk after the previous tile's last MMA.
//
// Tiles are walked M-major and every CTA pair owns a CONTIGUOUS range of them (balanced to +-1 tile), so an A panel is
// loaded 2-3 times per kernel and pair.  One ring of 16 KB slots carries both operand kinds in program order: the 8
// A k-blocks of a new row block, then 8 W-half k-blocks per tile.
//
// Warp roles (640 threads): warps 0-15 epilogue (as in the 16-warp variant of gemm2: warpgroup g drains the 64-column
// chunk g as two 32-column sub-chunks through two 8 KB SWIZZLE_64B staging buffers), warp 16 TMA producer (both CTAs),
// warp 17 MMA / copy issuer (leader CTA), warps 18-19 aux (per-tile row statistics, bias, colsum -> shared memory).
#pragma once
#include "../gemm2.cuh"

namespace ddb {

constexpr int G3_EPI_WARPS = 16;                       // 4 epilogue warpgroups
constexpr int G3_THREADS = (G3_EPI_WARPS + 4) * 32;     // + producer, MMA issuer, 2 aux warps

// smem -> TMEM copy of a 128-row x 256-bit (16 bf16) slice per CTA of the pair; same matrix descriptor as an MMA operand
__device__ __forceinline__ void tmem_cp_128x256b_2cta(uint32_t taddr, uint64_t sdesc) {
    asm volatile("tcgen05.cp.cta_group::2.128x256b [%0], %1;" ::"r"(taddr), "l"(sdesc) : "memory");
}
// CTA-pair MMA with A from TMEM (each CTA's TMEM holds its own 128 rows), B from shared memory
__device__ __forceinline__ void umma_f16_ts_2cta(uint32_t d_tmem, uint32_t a_tmem, uint64_t b_desc, uint32_t idesc,
                                                 uint32_t accumulate) {
    asm volatile(
        "{\n\t.reg .pred p;\n\t"
        "setp.ne.b32 p, %4, 0;\n\t"
        "tcgen05.mma.cta_group::2.kind::f16 [%0], [%1], %2, %3, p;\n\t}" ::"r"(d_tmem),
        "r"(a_tmem), "l"(b_desc), "r"(idesc), "r"(accumulate)
        : "memory");
}

template <int RING_, bool LN_>
struct Gemm3Cfg {
    static constexpr int BM = 128;  // rows per CTA (256 per pair)
    static constexpr int BN = 256;
    static constexpr int BK = 64;
    static constexpr int RING = RING_;                   // 16 KB slots: A k-block [128 x 64] or W-half k-block [128 x 64]
    static constexpr int SLOT_BYTES = BM * BK * 2;
    static constexpr int OUT_BUF_BYTES = 128 * 64;       // 128 rows x 32 bf16 (one 32-column sub-chunk), 64B-swizzled
    static constexpr int AUX_BYTES = LN_ ? 3072 : 1024;  // per buffer: [rowstats 1 KB][bias 1 KB][colsum 1 KB]
    static constexpr int OFF_RING = 0;
    static constexpr int OFF_OUT = OFF_RING + RING * SLOT_BYTES;
    static constexpr int OFF_AUX = OFF_OUT + 8 * OUT_BUF_BYTES;  // 2 staging buffers per epilogue warpgroup
    static constexpr int OFF_BAR = OFF_AUX + 2 * AUX_BYTES;
    static constexpr int NUM_BARS = 2 * RING + 2 + 8 + 4;
    static constexpr int SMEM_BYTES = OFF_BAR + NUM_BARS * 8 + 16;
    static constexpr uint32_t TMEM_A = 256;              // first TMEM column of the A panel (K/2 columns)
    static_assert(SMEM_BYTES <= 232448, "exceeds 227 KB of shared memory");
};

template <int EPI, bool STATS, int RING>
__global__ void __cluster_dims__(2, 1, 1) __launch_bounds__(G3_THREADS, 1)
    gemm3_tcgen05_kernel(const __grid_constant__ GemmArgs a) {
    constexpr bool kLN = (EPI == EPI_LN || EPI == EPI_LN_GELU);
    using Cfg = Gemm3Cfg<RING, kLN>;
    constexpr int BN = Cfg::BN;
    constexpr int STAGES = RING;

    extern __shared__ __align__(1024) uint8_t smem[];
    uint8_t* sRing = smem + Cfg::OFF_RING;
    uint8_t* sOut = smem + Cfg::OFF_OUT;
    uint8_t* sAux = smem + Cfg::OFF_AUX;
    uint64_t* bars = reinterpret_cast<uint64_t*>(smem + Cfg::OFF_BAR);
    uint64_t* full_bar = bars;                  // [STAGES]  (used on the leader)
    uint64_t* empty_bar = full_bar + STAGES;    // [STAGES]  (per CTA, multicast-arrived by the leader's commits)
    uint64_t* tfull_bar = empty_bar + STAGES;   // [1]       (per CTA) accumulator complete
    uint64_t* tempty_bar = tfull_bar + 1;       // [1]       (used on the leader; 32 arrivals) accumulator in registers
    uint64_t* res_bar = tempty_bar + 1;         // [8]       residual sub-chunk landed in (warpgroup g, buffer b): g*2+b
    uint64_t* aux_full = res_bar + 8;           // [2]  64 arrivals (aux threads)
    uint64_t* aux_empty = aux_full + 2;         // [2]  16 arrivals (epilogue warps)
    uint32_t* tmem_holder = reinterpret_cast<uint32_t*>(aux_empty + 2);

    const int warp = threadIdx.x >> 5;
    const int lane = threadIdx.x & 31;
    const uint32_t rank = cluster_ctarank();
    const bool leader = (rank == 0);
    const int cluster_id = blockIdx.x >> 1;
    const int num_clusters = gridDim.x >> 1;

    // PDL: the set-up below (barriers, TMEM, descriptor prefetch, cluster sync) overlaps the predecessor's tail
    pdl_launch_dependents();
    if (a.m_dev) pdl_wait();  // the live row count is written by an earlier kernel of the step
    const int M = a.m_dev ? *a.m_dev : a.M;
    const int nblk_n = a.N / BN;
    const int nblk_m = (M + 255) / 256;
    const int num_tiles = nblk_m * nblk_n;
    const int nkb = a.K0 / Cfg::BK;  // <= 8 (K <= 512); no second K source in this kernel
    // contiguous, balanced tile range of this CTA pair (M-major tile order: consecutive tiles share the A panel)
    const int tile_lo = (int)(((long long)cluster_id * num_tiles) / num_clusters);
    const int tile_hi = (int)(((long long)(cluster_id + 1) * num_tiles) / num_clusters);

    constexpr int kProducerWarp = G3_EPI_WARPS, kMmaWarp = G3_EPI_WARPS + 1, kAllocWarp = G3_EPI_WARPS + 2;
    if (threadIdx.x == 0 && (smem_u32(smem) & 1023u)) __trap();  // swizzled tiles need 1024-byte alignment
    if (warp == kProducerWarp && lane == 0) {
        tma_prefetch_desc(&a.tmA0);
        tma_prefetch_desc(&a.tmB2);
        tma_prefetch_desc(&a.tmOut2);
        if (EPI == EPI_RES) tma_prefetch_desc(&a.tmRes2);
    }
    if (warp == kMmaWarp && lane == 0) {
        for (int i = 0; i < STAGES; ++i) {
            mbar_init(&full_bar[i], 1);
            mbar_init(&empty_bar[i], 1);
        }
        mbar_init(&tfull_bar[0], 1);
        mbar_init(&tempty_bar[0], 2 * G3_EPI_WARPS);  // epilogue warps x 2 CTAs
        for (int i = 0; i < 2; ++i) {
            mbar_init(&aux_full[i], 64);
            mbar_init(&aux_empty[i], G3_EPI_WARPS);
        }
        for (int i = 0; i < 8; ++i) mbar_init(&res_bar[i], 1);
        fence_mbar_init();
    }
    if (warp == kAllocWarp) tmem_alloc_2cta<512>(tmem_holder);
    tc_fence_before();
    cluster_sync_all();  // barriers of both CTAs initialised before any remote arrive / multicast commit
    tc_fence_after();
    const uint32_t tmem_base = *tmem_holder;
    pdl_wait();  // everything below reads or overwrites buffers of earlier kernels
    auto gtime = []() { unsigned long long t; asm volatile("mov.u64 %0, %%globaltimer;" : "=l"(t)); return (long long)t; };
    if (a.trace && leader && threadIdx.x == 0) a.trace[256 + cluster_id * 4 + 0] = gtime();

    if (warp == kProducerWarp) {
        // ===================================================================== TMA producer (both CTAs)
        if (lane == 0) {
            int slot = 0;
            uint32_t phase = 0;
            auto load_slot = [&](const CUtensorMap* tm, int c0, int c1) {
                mbar_wait(&empty_bar[slot], phase ^ 1);
                const uint32_t fb = leader_smem_addr(&full_bar[slot]);
                if (leader) mbar_expect_tx(&full_bar[slot], 2 * Cfg::SLOT_BYTES);
                tma_load_2d_2cta(sRing + slot * Cfg::SLOT_BYTES, tm, fb, c0, c1);
                if (++slot == RING) {
                    slot = 0;
                    phase ^= 1;
                }
            };
            for (int tile = tile_lo; tile < tile_hi; ++tile) {
                const int m_blk = tile / nblk_n, n_blk = tile % nblk_n;
                if (tile == tile_lo || n_blk == 0)  // new 256-row block: its A panel goes first
                    for (int kb = 0; kb < nkb; ++kb) load_slot(&a.tmA0, kb * Cfg::BK, m_blk * 256 + (int)rank * 128);
                for (int kb = 0; kb < nkb; ++kb) load_slot(&a.tmB2, kb * Cfg::BK, n_blk * BN + (int)rank * 128);
            }
        }
    } else if (warp == kMmaWarp) {
        // ===================================================================== MMA / copy issuer (leader CTA)
        if (leader && lane == 0) {
            constexpr uint32_t idesc = umma_idesc_bf16(256, BN);
            const uint32_t a_tmem = tmem_base + Cfg::TMEM_A;
            int slot = 0;
            uint32_t phase = 0;
            int it = 0;
            for (int tile = tile_lo; tile < tile_hi; ++tile, ++it) {
                const int n_blk = tile % nblk_n;
                long long* tr = (a.trace && cluster_id == 0 && it < 16) ? a.trace + it * 16 : nullptr;
                if (tr) tr[0] = clock64();
                if (tile == tile_lo || n_blk == 0) {
                    // A panel of the new row block: shared memory -> TMEM, 16 K-elements (8 columns) per copy.  The
                    // copies queue behind the previous tile's MMAs in the tcgen05 pipeline, which still read the old
                    // panel.
                    for (int kb = 0; kb < nkb; ++kb) {
                        mbar_wait(&full_bar[slot], phase);
                        tc_fence_after();
                        const uint64_t desc = umma_desc_kmajor_sw128(smem_u32(sRing + slot * Cfg::SLOT_BYTES));
#pragma unroll
                        for (int k = 0; k < Cfg::BK / 16; ++k)
                            tmem_cp_128x256b_2cta(a_tmem + kb * 32 + k * 8, desc + 2 * k);
                        umma_commit_2cta(&empty_bar[slot], 0x3);
                        if (++slot == RING) {
                            slot = 0;
                            phase ^= 1;
                        }
                    }
                }
                mbar_wait(&tempty_bar[0], (it & 1) ^ 1);  // the epilogue holds the previous tile in registers
                tc_fence_after();
                if (tr) tr[1] = clock64();
                for (int kb = 0; kb < nkb; ++kb) {
                    mbar_wait(&full_bar[slot], phase);
                    tc_fence_after();
                    if (tr && kb == 0) tr[2] = clock64();
                    const uint64_t b_desc = umma_desc_kmajor_sw128(smem_u32(sRing + slot * Cfg::SLOT_BYTES));
                    if (!(a.debug & 2)) {
#pragma unroll
                        for (int k = 0; k < Cfg::BK / 16; ++k)
                            umma_f16_ts_2cta(tmem_base, a_tmem + kb * 32 + k * 8, b_desc + 2 * k, idesc, (kb | k) != 0);
                    }
                    umma_commit_2cta(&empty_bar[slot], 0x3);
                    if (++slot == RING) {
                        slot = 0;
                        phase ^= 1;
                    }
                }
                umma_commit_2cta(&tfull_bar[0], 0x3);
                if (tr) tr[3] = clock64();
            }
        }
    } else if (warp >= kAllocWarp) {
        // ===================================================================== aux warps: per-tile vectors -> smem
        const int t = threadIdx.x - kAllocWarp * 32;  // 0..63
        int it = 0;
        for (int tile = tile_lo; tile < tile_hi; ++tile, ++it) {
            const int m_blk = tile / nblk_n, n_blk = tile % nblk_n;
            const int buf = it & 1;
            const uint32_t ph = (it >> 1) & 1;
            uint8_t* ab = sAux + buf * Cfg::AUX_BYTES;
            mbar_wait(&aux_empty[buf], ph ^ 1);
            const int col = n_blk * BN + t * 4;
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
                *reinterpret_cast<float4*>(ab + 1024 + t * 16) =
                    a.bias ? __ldg(reinterpret_cast<const float4*>(a.bias + col)) : make_float4(0.f, 0.f, 0.f, 0.f);
                *reinterpret_cast<float4*>(ab + 2048 + t * 16) = __ldg(reinterpret_cast<const float4*>(a.colsum + col));
            } else {
                *reinterpret_cast<float4*>(ab + t * 16) =
                    a.bias ? __ldg(reinterpret_cast<const float4*>(a.bias + col)) : make_float4(0.f, 0.f, 0.f, 0.f);
            }
            mbar_arrive(&aux_full[buf]);  // release: the st.shared above are visible to the waiting epilogue warps
        }
    } else {
        // ===================================================================== epilogue (4 warpgroups per CTA)
        // Warpgroup g owns the 64-column chunk g of every tile and drains it as two 32-column sub-chunks through two
        // 8 KB staging buffers (TMA SWIZZLE_64B), so the store of one sub-chunk overlaps the math of the next.
        const int g = warp >> 2;        // warpgroup == 64-column chunk of the tile
        const int quarter = warp & 3;   // TMEM lane quarter this warp may access
        const int et = threadIdx.x & 127;
        const int row_in_tile = quarter * 32 + lane;
        const uint32_t bar_id = 1 + g;
        uint8_t* my_bufs = sOut + g * 2 * Cfg::OUT_BUF_BYTES;
        uint64_t* my_res = res_bar + g * 2;
        const bool traffic = !(a.debug & 8);
        const uint32_t swz = (row_in_tile >> 1) & 3;  // 64B swizzle: 16-byte chunk index ^= (row / 2) % 4

        // residual of (tile, half h) -> staging buffer h (issued by thread 0 of the warpgroup once the buffer's
        // previous TMA store has been read)
        auto prefetch_res = [&](int tile, int h) {
            if (tile >= tile_hi) return;
            const int m_blk = tile / nblk_n, n_blk = tile % nblk_n;
            uint64_t* rb = &my_res[h];
            mbar_expect_tx(rb, Cfg::OUT_BUF_BYTES);
            tma_load_2d(my_bufs + h * Cfg::OUT_BUF_BYTES, &a.tmRes2, rb, n_blk * BN + g * 64 + h * 32,
                        m_blk * 256 + (int)rank * 128);
        };
        const bool res_on = (EPI == EPI_RES) && traffic && !(a.debug & 1);
        if (res_on && et == 0) prefetch_res(tile_lo, 0);

        int it = 0;
        for (int tile = tile_lo; tile < tile_hi; ++tile, ++it) {
            const int m_blk = tile / nblk_n, n_blk = tile % nblk_n;
            const int as = it & 1;              // aux buffer
            const uint32_t aph = (it >> 1) & 1;
            const int row0 = m_blk * 256 + (int)rank * 128;
            const int row = row0 + row_in_tile;
            const int col0 = n_blk * BN + g * 64;
            const uint8_t* ab = sAux + as * Cfg::AUX_BYTES;
            const float* sbias = reinterpret_cast<const float*>(ab + (kLN ? 1024 : 0)) + g * 64;
            const float* scs = reinterpret_cast<const float*>(ab + 2048) + g * 64;

            long long* tr = (a.trace && cluster_id == 0 && leader && threadIdx.x == 0 && it < 16) ? a.trace + it * 16 : nullptr;
            if (tr) tr[4] = clock64();
            if (res_on && et == 0) {
                // buffer 1 was last stored by the previous tile's second half: once that store has been read, pull
                // this tile's second-half residual into it (hidden behind the wait for the accumulator)
                tma_store_wait_read<0>();
                prefetch_res(tile, 1);
            }
            mbar_wait(&aux_full[as], aph);
            if (tr) tr[5] = clock64();
            float rstd = 1.f, mean_rstd = 0.f;
            if constexpr (kLN) {
                const float2 rs = reinterpret_cast<const float2*>(ab)[row_in_tile];
                rstd = rs.x, mean_rstd = rs.y;
            }
            mbar_wait(&tfull_bar[0], it & 1);
            tc_fence_after();
            if (tr) tr[6] = clock64();
            const uint32_t t_row = tmem_base + (uint32_t(quarter * 32) << 16) + g * 64;

            uint32_t acc[2][32];
            tmem_ld_32x32b_x32(t_row, acc[0]);
            tmem_ld_32x32b_x32(t_row + 32, acc[1]);
            tmem_ld_wait();
            if (tr) tr[7] = clock64();
            tc_fence_before();
            __syncwarp();
            if (lane == 0) mbar_arrive_cluster(leader_smem_addr(&tempty_bar[0]));  // accumulator is in registers
            if (a.debug & 1) {
                if (lane == 0) mbar_arrive(&aux_empty[as]);
                continue;
            }

            const f32x2 rstd2 = f2_splat(rstd), nmr2 = f2_splat(-mean_rstd);
            f32x2 s1 = f2_splat(0.f), s2 = f2_splat(0.f), nshift = f2_splat(0.f);
#pragma unroll
            for (int h = 0; h < 2; ++h) {
                const uint32_t q = 2u * (uint32_t)it + (uint32_t)h;
                uint8_t* sbuf = my_bufs + h * Cfg::OUT_BUF_BYTES;  // q & 1 == h
                if (res_on) {
                    mbar_wait(&my_res[h], (q >> 1) & 1);  // residual sub-chunk landed in sbuf
                } else {
                    // buffer h was last used by sub-chunk q-2: its TMA store must have finished reading it
                    if (et == 0) tma_store_wait_read<1>();
                    named_bar_sync(bar_id, 128);
                }
                if (tr) tr[8 + h * 4] = clock64();
                uint8_t* srow = sbuf + row_in_tile * 64;
#pragma unroll
                for (int j = 0; j < 4; ++j) {
                    const int cl = h * 32 + j * 8;  // column inside the 64-column chunk
                    f32x2 v[4];
#pragma unroll
                    for (int e = 0; e < 4; ++e) v[e] = f2_pack_u(acc[h][j * 8 + 2 * e], acc[h][j * 8 + 2 * e + 1]);
                    const float4 b0 = *reinterpret_cast<const float4*>(sbias + cl);
                    const float4 b1 = *reinterpret_cast<const float4*>(sbias + cl + 4);
                    const f32x2 bb[4] = {f2_pack(b0.x, b0.y), f2_pack(b0.z, b0.w), f2_pack(b1.x, b1.y),
                                         f2_pack(b1.z, b1.w)};
                    if constexpr (kLN) {
                        const float4 c0 = *reinterpret_cast<const float4*>(scs + cl);
                        const float4 c1 = *reinterpret_cast<const float4*>(scs + cl + 4);
                        const f32x2 cs[4] = {f2_pack(c0.x, c0.y), f2_pack(c0.z, c0.w), f2_pack(c1.x, c1.y),
                                             f2_pack(c1.z, c1.w)};
#pragma unroll
                        for (int e = 0; e < 4; ++e) v[e] = f2_fma(v[e], rstd2, f2_fma(nmr2, cs[e], bb[e]));
                    } else {
#pragma unroll
                        for (int e = 0; e < 4; ++e) v[e] = f2_add(v[e], bb[e]);
                    }
                    if constexpr (EPI == EPI_LN_GELU) {  // (no run-time switches in this loop, see gemm2.cuh)
#pragma unroll
                        for (int e = 0; e < 4; ++e) v[e] = gelu_fast2(v[e]);
                    }
                    uint4* sp = reinterpret_cast<uint4*>(srow + (((uint32_t)j ^ swz) << 4));
                    if constexpr (EPI == EPI_RES) {
                        const uint4 r = *sp;
                        v[0] = f2_add(v[0], f2_pack_u(r.x << 16, r.x & 0xFFFF0000u));
                        v[1] = f2_add(v[1], f2_pack_u(r.y << 16, r.y & 0xFFFF0000u));
                        v[2] = f2_add(v[2], f2_pack_u(r.z << 16, r.z & 0xFFFF0000u));
                        v[3] = f2_add(v[3], f2_pack_u(r.w << 16, r.w & 0xFFFF0000u));
                    }
                    if constexpr (STATS) {
                        // shifted single-pass statistics: s1 = sum(v - v0), s2 = sum((v - v0)^2); lanes = even/odd cols
                        if (h == 0 && j == 0) {
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
                if (h == 1) {
                    if constexpr (STATS) {
                        // partial LayerNorm statistic (mean, M2) of this 64-column chunk for the next consumer
                        float s1a, s1b, s2a, s2b, ns, ns_;
                        f2_unpack(s1, s1a, s1b);
                        f2_unpack(s2, s2a, s2b);
                        f2_unpack(nshift, ns, ns_);
                        const float t1 = s1a + s1b, t2 = s2a + s2b;
                        const float dm = t1 * (1.f / 64.f);
                        if (row < M)
                            a.stats_out[(size_t)row * (a.N >> 6) + (col0 >> 6)] =
                                make_float2(dm - ns, fmaf(-t1, dm, t2));
                    }
                    __syncwarp();
                    if (lane == 0) mbar_arrive(&aux_empty[as]);  // this warp is done with the tile's smem vectors
                }
                if (tr) tr[9 + h * 4] = clock64();
                fence_proxy_async_smem();
                named_bar_sync(bar_id, 128);
                if (tr) tr[10 + h * 4] = clock64();
                if (et == 0 && traffic) {
                    tma_store_2d(&a.tmOut2, sbuf, col0 + h * 32, row0);
                    tma_store_commit();
                    if (res_on && h == 1) {
                        // buffer 0 (first half of this tile) has been read by its store: next tile's first half
                        tma_store_wait_read<1>();
                        prefetch_res(tile + 1, 0);
                    }
                }
            }
        }
        if (a.trace && leader && threadIdx.x == 0) a.trace[256 + cluster_id * 4 + 1] = gtime();
        if (et == 0) tma_store_wait_all<0>();
        if (a.trace && leader && threadIdx.x == 0) a.trace[256 + cluster_id * 4 + 2] = gtime();
    }

    __syncwarp();  // the producer / MMA warps ran single-lane role loops: reconverge before the .aligned cluster barrier
    tc_fence_before();
    cluster_sync_all();  // peer may still multicast into / read from this CTA until both are done
    if (warp == kAllocWarp) {
        tc_fence_after();
        tmem_dealloc_2cta<512>(tmem_base);
    }
}

}  // namespace ddb
