// Variant of attention_tcgen05_kernel (attention.cuh) with TWO softmax threads per query row.
//
// Why (profiles/r01_attention_microbench.txt): the one-thread-per-row kernel is paced by the serial chain of a tile --
// S -> row max -> exp -> P.V -> normalise, ~8 000 clk per item with two tiles per SM in flight (the TMEM limit) -- while
// the MUFU pipe is 52 % busy and moving exponentials to the FMA pipes changes nothing.  Splitting every row's 256 score
// columns over two threads (16 softmax warps) halves the two passes and the epilogue of that chain; the row maximum and
// row sum cross between the two threads through shared memory under a 256-thread named barrier per tile.
// Everything else (operand staging, MMA issue order, extras rows on mma.sync, P layout in TMEM) is unchanged.
//
// MEASURED SLOWER (B = 128, L = 257, H = 8: 53.8 us against 43.4 us; same 1.87e-3 rel-L2) and therefore NOT the model
// path (ddb_op_attention variant 3 / ddb_set_option "attn_x2" only).  The clock64 trace shows why: with sixteen softmax
// warps both tiles run their exp pass at the same time (the two-stage operand ring makes the tiles lock-step), four
// warps per scheduler then share the MUFU pipe and pass 2 takes ~4 500 clk instead of 3 600; the extras warp on
// mma.sync (spilling under the 102-register cap of 640 threads) releases the operand stage late, so the next item's
// operands -- and with them everything else -- arrive thousands of clocks late.  Moving half of the exponentials to the
// FMA pipes makes it slower still (63 us): issue slots, not the MUFU pipe, are then short.  The one-thread-per-row kernel
// wins because its two tiles run half a period apart and fill each other's gaps.
#pragma once
#include "../attention.cuh"

namespace ddb {

constexpr uint32_t ATT4_QK_BYTES = 16384 + 16384 + 32768 + 2048 + 2048;  // Q0, Q1, K, Kx, Qx on one barrier


constexpr int ATT4_THREADS = 640;  // 16 softmax warps + producer, 2 MMA issuers, extras warp
constexpr int ATT4_SMEM = ATT3_SMEM + 4096;

__global__ void __launch_bounds__(ATT4_THREADS, 1) attention_tcgen05_x2_kernel(const __grid_constant__ AttnArgs a) {
    extern __shared__ __align__(1024) uint8_t smem[];
    uint64_t* bars = reinterpret_cast<uint64_t*>(smem + ATT3_OFF_BAR);
    uint64_t* qk_full = bars + 0;      // [2] per operand stage
    uint64_t* v_full = bars + 2;       // [2]
    uint64_t* stage_empty = bars + 4;  // [2] 5 arrivals: each tile's MMAs retired + its TMA store read, extras warp
    uint64_t* s_full = bars + 6;       // [2] per query tile
    uint64_t* p_full = bars + 8;       // [2]
    uint64_t* o_full = bars + 10;      // [2]
    uint64_t* tmem_free = bars + 12;   // [2]
    uint64_t* p_half = bars + 14;      // [2] P of keys [0, 128) written
    uint32_t* tmem_holder = reinterpret_cast<uint32_t*>(bars + 16);
    float* xch = reinterpret_cast<float*>(smem + ATT3_OFF_BAR + 256);  // [max | sum][tile][half][128] exchange

    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    const int D = a.H * 64;
    pdl_launch_dependents();   // PDL: the set-up below overlaps the predecessor's tail
    if (a.b_dev) pdl_wait();   // the live batch size is written by an earlier kernel of the step
    const int n_items = (a.b_dev ? *a.b_dev : a.B) * a.H;
    const int my_items =
        (n_items > (int)blockIdx.x) ? (n_items - (int)blockIdx.x + (int)gridDim.x - 1) / (int)gridDim.x : 0;

    if (threadIdx.x == 0 && (smem_u32(smem) & 1023u)) __trap();  // swizzled tiles need 1024-byte alignment
    if (warp == 16 && lane == 0) {
        tma_prefetch_desc(&a.tmQKV);
        tma_prefetch_desc(&a.tmKV);
        tma_prefetch_desc(&a.tmX);
        tma_prefetch_desc(&a.tmOut);
        for (int i = 0; i < 2; ++i) {
            mbar_init(&qk_full[i], 1);
            mbar_init(&v_full[i], 1);
            mbar_init(&stage_empty[i], 5);
            mbar_init(&s_full[i], 1);
            mbar_init(&p_full[i], 4);
            mbar_init(&o_full[i], 1);
            mbar_init(&tmem_free[i], 8);
            mbar_init(&p_half[i], 4);
        }
        fence_mbar_init();
    }
    if (warp == 17) tmem_alloc<512>(tmem_holder);
    tc_fence_before();
    __syncthreads();
    tc_fence_after();
    const uint32_t tmem = *tmem_holder;
    pdl_wait();  // qkv is the predecessor's output

    if (warp == 16) {
        // ================================================================= TMA producer
        if (lane == 0) {
            for (int it = 0; it < my_items; ++it) {
                const int item = blockIdx.x + it * gridDim.x;
                const int b = item / a.H, h = item % a.H;
                const int s = it & 1;
                uint8_t* st = smem + s * ATT3_STAGE;
                mbar_wait(&stage_empty[s], ((it >> 1) & 1) ^ 1);
                mbar_expect_tx(&qk_full[s], ATT4_QK_BYTES);
                tma_load_3d(st + ATT3_OFF_K, &a.tmKV, &qk_full[s], D + h * 64, a.extras, b);
                tma_load_3d(st, &a.tmQKV, &qk_full[s], h * 64, a.extras, b);
                tma_load_3d(st + 16384, &a.tmQKV, &qk_full[s], h * 64, a.extras + 128, b);
                tma_load_3d(st + ATT3_OFF_KX, &a.tmX, &qk_full[s], D + h * 64, 0, b);
                tma_load_3d(st + ATT3_OFF_QX, &a.tmX, &qk_full[s], h * 64, 0, b);
                mbar_expect_tx(&v_full[s], ATT3_V_BYTES);
                tma_load_3d(st + ATT3_OFF_V, &a.tmKV, &v_full[s], 2 * D + h * 64, a.extras, b);
                tma_load_3d(st + ATT3_OFF_VX, &a.tmX, &v_full[s], 2 * D + h * 64, 0, b);
            }
        }
    } else if (warp == 17 || warp == 19) {
        // ================================================================= MMA issuers: one thread per query tile
        // (two independent in-order streams, so neither tile ever waits behind the other tile's barrier)
        if (lane == 0) {
            const int t = (warp == 17) ? 0 : 1;
            constexpr uint32_t idesc_s = umma_idesc_bf16(128, 256);
            constexpr uint32_t idesc_o = umma_idesc_bf16(128, 64, 0, 1);
            for (int it = 0; it < my_items; ++it) {
                const int s = it & 1;
                const uint8_t* st = smem + s * ATT3_STAGE;
                // S_t(it): needs the operands of the item and the tile's TMEM columns (O_t(it-1) drained)
                mbar_wait(&qk_full[s], (it >> 1) & 1);
                mbar_wait(&tmem_free[t], (it & 1) ^ 1);
                // start tile 1 half a period late so the two softmax warpgroups do not fight over the MUFU pipe
                if (t == 1 && it == 0) mbar_wait(&p_full[0], 0);
                tc_fence_after();
                const uint64_t dq = umma_desc_kmajor_sw128(smem_u32(st + t * 16384));
                const uint64_t dk = umma_desc_kmajor_sw128(smem_u32(st + ATT3_OFF_K));
#pragma unroll
                for (int k = 0; k < 4; ++k) umma_f16_ss(tmem + t * 256, dq + 2 * k, dk + 2 * k, idesc_s, k != 0);
                umma_commit(&s_full[t]);
                // O_t(it) = P_t [V; V_x]: 16 keys per k-step (P columns +8, V rows +16 = 2048 B); the first 8
                // k-steps start as soon as the first half of P is written
                long long* tr = (a.trace && blockIdx.x == 0) ? a.trace + (it * 2 + t) * 16 + 8 : nullptr;
                if (tr) tr[0] = clock64();
                mbar_wait(&v_full[s], (it >> 1) & 1);
                const uint64_t dv = umma_desc_mnmajor_sw128(smem_u32(st + ATT3_OFF_V));
                mbar_wait(&p_half[t], it & 1);
                tc_fence_after();
                if (tr) tr[1] = clock64();
#pragma unroll
                for (int k = 0; k < 8; ++k)
                    umma_f16_ts(tmem + t * 256 + 64, tmem + t * 256 + 8 * k, dv + (uint64_t)(k * (2048 >> 4)),
                                idesc_o, k != 0);
                if (tr) tr[2] = clock64();
                mbar_wait(&p_full[t], it & 1);
                tc_fence_after();
                if (tr) tr[3] = clock64();
#pragma unroll
                for (int k = 8; k < 17; ++k)
                    umma_f16_ts(tmem + t * 256 + 64, tmem + t * 256 + 64 + 8 * k, dv + (uint64_t)(k * (2048 >> 4)),
                                idesc_o, 1u);
                umma_commit(&o_full[t]);
                umma_commit(&stage_empty[s]);  // this tile's MMAs no longer read the operand stage once retired
                if (tr) {
                    tr[4] = clock64();
                    mbar_wait(&o_full[t], it & 1);
                    tr[5] = clock64();
                }
            }
        }
    } else if (warp == 18) {
        // ================================================================= extras query rows on mma.sync
        const int g = lane >> 2, t = lane & 3;
        for (int it = 0; it < my_items; ++it) {
            const int item = blockIdx.x + it * gridDim.x;
            const int b = item / a.H, h = item % a.H;
            const int s = it & 1;
            const uint8_t* st = smem + s * ATT3_STAGE;
            mbar_wait(&qk_full[s], (it >> 1) & 1);
            // A fragments: rows g (token g of the sample; only g < extras is kept), rows g+8 are zero
            uint32_t qf[4][4];
#pragma unroll
            for (int ks = 0; ks < 4; ++ks) {
                qf[ks][0] = *reinterpret_cast<const uint32_t*>(st + ATT3_OFF_QX + att_swz(g, ks * 2) + 4 * t);
                qf[ks][2] = *reinterpret_cast<const uint32_t*>(st + ATT3_OFF_QX + att_swz(g, ks * 2 + 1) + 4 * t);
                qf[ks][1] = qf[ks][3] = 0u;
            }
            AttRowState rs;
            rs.init();
            mbar_wait(&v_full[s], (it >> 1) & 1);
            // extras keys: the first 8 tokens of the X tile, of which [0, extras) are valid
            att_mma_block<true>(smem_u32(st + ATT3_OFF_KX), smem_u32(st + ATT3_OFF_VX), 0, 1, a.extras, qf, rs,
                                a.scale_log2e, lane);
            const uint32_t sK_u = smem_u32(st + ATT3_OFF_K), sV_u = smem_u32(st + ATT3_OFF_V);
#pragma unroll 1
            for (int kb0 = 0; kb0 < 256; kb0 += 64)
                att_mma_block<true>(sK_u, sV_u, kb0, 8, 256, qf, rs, a.scale_log2e, lane);
            __syncwarp();
            if (lane == 0) mbar_arrive(&stage_empty[s]);  // this warp no longer reads the stage
            float l0 = rs.l0;
            l0 += __shfl_xor_sync(0xffffffffu, l0, 1);
            l0 += __shfl_xor_sync(0xffffffffu, l0, 2);
            if (g < a.extras) {
                const float inv0 = 1.f / l0;
                __nv_bfloat16* o0 = a.out + ((size_t)b * a.L + g) * D + h * 64 + 2 * t;
#pragma unroll
                for (int i = 0; i < 8; ++i)
                    *reinterpret_cast<uint32_t*>(o0 + i * 8) = pack_bf16(rs.o[i][0] * inv0, rs.o[i][1] * inv0);
            }
        }
    } else {
        // ================================================================= softmax + epilogue: TWO threads per query row
        // warps 8t .. 8t+3 take keys [0, 128) (and the extras keys) of tile t, warps 8t+4 .. 8t+7 keys [128, 256); a
        // warp may only touch the TMEM lanes 32 (warp % 4) .., so warp w and w + 4 share the same 32 rows.  Row maximum
        // and row sum are exchanged through shared memory under a 256-thread named barrier per tile.
        const int t = warp >> 3;
        const int half = (warp >> 2) & 1;
        const int quarter = warp & 3;
        const int r = quarter * 32 + lane;
        const int et = threadIdx.x & 255;
        const uint32_t t_row = tmem + (uint32_t(quarter * 32) << 16) + t * 256;
        const uint32_t t_col = t_row + half * 128;  // this thread's 128 score columns
        const float c = a.scale_log2e;
        const int q0 = a.extras + t * 128;
        float* xmax = xch + t * 256;        // [half][128]
        float* xsum = xch + 512 + t * 256;  // [half][128]
        const uint32_t bar_id = 1 + t;

        // scores of query row r against the extras keys (tokens [0, extras)) of item `it`, on the CUDA cores
        auto extras_scores = [&](int it, float& se0, float& se1) {
            const int s = it & 1;
            const uint8_t* sQ = smem + s * ATT3_STAGE + t * 16384;
            const uint8_t* sKx = smem + s * ATT3_STAGE + ATT3_OFF_KX;
            mbar_wait(&qk_full[s], (it >> 1) & 1);
            se0 = 0.f, se1 = 0.f;
#pragma unroll
            for (int j = 0; j < 8; ++j) {
                const uint4 q = *reinterpret_cast<const uint4*>(sQ + r * 128 + ((j ^ (r & 7)) << 4));
                const float qf[8] = {bf16_lo(q.x), bf16_hi(q.x), bf16_lo(q.y), bf16_hi(q.y),
                                     bf16_lo(q.z), bf16_hi(q.z), bf16_lo(q.w), bf16_hi(q.w)};
                const uint4 k0 = *reinterpret_cast<const uint4*>(sKx + att_swz(0, j));
                const float kf[8] = {bf16_lo(k0.x), bf16_hi(k0.x), bf16_lo(k0.y), bf16_hi(k0.y),
                                     bf16_lo(k0.z), bf16_hi(k0.z), bf16_lo(k0.w), bf16_hi(k0.w)};
#pragma unroll
                for (int e = 0; e < 8; ++e) se0 = fmaf(qf[e], kf[e], se0);
                if (a.extras == 2) {
                    const uint4 k1 = *reinterpret_cast<const uint4*>(sKx + att_swz(1, j));
                    const float kg[8] = {bf16_lo(k1.x), bf16_hi(k1.x), bf16_lo(k1.y), bf16_hi(k1.y),
                                         bf16_lo(k1.z), bf16_hi(k1.z), bf16_lo(k1.w), bf16_hi(k1.w)};
#pragma unroll
                    for (int e = 0; e < 8; ++e) se1 = fmaf(qf[e], kg[e], se1);
                }
            }
            if (a.extras != 2) se1 = -INFINITY;
        };

        float se0 = -INFINITY, se1 = -INFINITY;  // the extras keys belong to the low half
        if (half == 0 && my_items > 0) extras_scores(0, se0, se1);
        for (int it = 0; it < my_items; ++it) {
            const int item = blockIdx.x + it * gridDim.x;
            const int b = item / a.H, h = item % a.H;
            const int s = it & 1;
            const uint32_t ph = it & 1;
            uint8_t* sQ = smem + s * ATT3_STAGE + t * 16384;  // this tile's Q; later its output staging buffer

            long long* tr = (a.trace && blockIdx.x == 0 && r == 0 && half == 0) ? a.trace + (it * 2 + t) * 16 : nullptr;
            if (tr) tr[0] = clock64();
            mbar_wait(&s_full[t], ph);
            tc_fence_after();
            if (tr) tr[1] = clock64();
            if (et == 0 && it > 0) {
                tma_store_wait_read<0>();
                mbar_arrive(&stage_empty[s ^ 1]);
            }
            if (tr) tr[2] = clock64();
            uint32_t va[32], vb[32];
            // ---- pass 1: maximum over this thread's 128 columns (load of chunk j+1 in flight during chunk j)
            float m0 = se0, m1 = se1, m2 = -INFINITY, m3 = -INFINITY;
            tmem_ld_32x32b_x32(t_col, va);
            tmem_ld_wait();
#pragma unroll
            for (int j = 0; j < 4; ++j) {
                uint32_t(&cur)[32] = (j & 1) ? vb : va;
                uint32_t(&nxt)[32] = (j & 1) ? va : vb;
                if (j < 3) tmem_ld_32x32b_x32(t_col + (j + 1) * 32, nxt);
#pragma unroll
                for (int e = 0; e < 32; e += 8) {
                    m0 = fmax3(m0, __uint_as_float(cur[e + 0]), __uint_as_float(cur[e + 1]));
                    m1 = fmax3(m1, __uint_as_float(cur[e + 2]), __uint_as_float(cur[e + 3]));
                    m2 = fmax3(m2, __uint_as_float(cur[e + 4]), __uint_as_float(cur[e + 5]));
                    m3 = fmax3(m3, __uint_as_float(cur[e + 6]), __uint_as_float(cur[e + 7]));
                }
                if (j < 3) tmem_ld_wait();
            }
            const float mloc = fmaxf(fmaxf(m0, m1), fmaxf(m2, m3));
            xmax[half * 128 + r] = mloc;
            named_bar_sync(bar_id, 256);
            const float mc = fmaxf(mloc, xmax[(half ^ 1) * 128 + r]) * c;
            if (tr) tr[3] = clock64();
            const float pe0 = (half == 0) ? ex2_approx(fmaf(se0, c, -mc)) : 0.f;
            const float pe1 = (half == 0 && a.extras == 2) ? ex2_approx(fmaf(se1, c, -mc)) : 0.f;
            // ---- pass 2: P = exp2(s*c - m*c) -> bf16, written over already-consumed score columns of the own half:
            // low half -> [0, 64), high half -> [128, 192)
            const f32x2 c2 = f2_splat(c), nmc2 = f2_splat(-mc);
            f32x2 sum2 = f2_pack(pe0, pe1), sum2b = f2_splat(0.f);
            tmem_ld_32x32b_x32(t_col, va);
            tmem_ld_wait();
#pragma unroll
            for (int j = 0; j < 4; ++j) {
                uint32_t(&cur)[32] = (j & 1) ? vb : va;
                uint32_t(&nxt)[32] = (j & 1) ? va : vb;
                if (j < 3) tmem_ld_32x32b_x32(t_col + (j + 1) * 32, nxt);
                uint32_t pk[16];
#pragma unroll
                for (int e = 0; e < 16; ++e) {
                    float x0, x1;
                    f2_unpack(f2_fma(f2_pack_u(cur[2 * e], cur[2 * e + 1]), c2, nmc2), x0, x1);
                    const f32x2 p = f2_pack(ex2_approx(x0), ex2_approx(x1));
                    if (e & 1)
                        sum2b = f2_add(sum2b, p);
                    else
                        sum2 = f2_add(sum2, p);
                    pk[e] = f2_to_bf16x2(p);
                }
                if (j < 3) tmem_ld_wait();
                tmem_st_32x32b_x16(t_col + j * 16, pk);
            }
            if (half == 0) {
                // 17th k-step: keys = tokens 0..15 of the sample, non-zero weight only for the extras tokens
                const uint32_t px[8] = {pack_bf16(pe0, pe1), 0u, 0u, 0u, 0u, 0u, 0u, 0u};
                tmem_st_32x32b_x8(t_row + 192, px);
            }
            tmem_st_wait();
            tc_fence_before();
            __syncwarp();
            if (lane == 0) mbar_arrive(half == 0 ? &p_half[t] : &p_full[t]);
            if (tr) tr[4] = clock64();
            {
                float sa, sb, sc, sd;
                f2_unpack(sum2, sa, sb);
                f2_unpack(sum2b, sc, sd);
                xsum[half * 128 + r] = (sa + sb) + (sc + sd);
            }
            // while the tensor core finishes O: the next item's extras-key scores (its operands landed long ago)
            if (half == 0 && it + 1 < my_items) extras_scores(it + 1, se0, se1);
            if (tr) tr[5] = clock64();

            // ---- epilogue: this thread's 32 columns of the O row, then release the tile's TMEM columns
            mbar_wait(&o_full[t], ph);
            tc_fence_after();
            if (tr) tr[6] = clock64();
            tmem_ld_32x32b_x32(t_row + 64 + half * 32, va);
            named_bar_sync(bar_id, 256);  // both partial sums are in shared memory
            const float inv = 1.f / (xsum[r] + xsum[128 + r]);
            tmem_ld_wait();
            tc_fence_before();
            __syncwarp();
            if (lane == 0) mbar_arrive(&tmem_free[t]);
            // row r of the Q tile is only touched by its two threads and by the (retired) S MMA: reuse it as staging
            uint8_t* srow = sQ + r * 128;
            const f32x2 inv2 = f2_splat(inv);
#pragma unroll
            for (int j = 0; j < 4; ++j) {
                const uint32_t* o = &va[j * 8];
                uint4 w;
                w.x = f2_to_bf16x2(f2_mul(f2_pack_u(o[0], o[1]), inv2));
                w.y = f2_to_bf16x2(f2_mul(f2_pack_u(o[2], o[3]), inv2));
                w.z = f2_to_bf16x2(f2_mul(f2_pack_u(o[4], o[5]), inv2));
                w.w = f2_to_bf16x2(f2_mul(f2_pack_u(o[6], o[7]), inv2));
                *reinterpret_cast<uint4*>(srow + (((half * 4 + j) ^ (r & 7)) << 4)) = w;
            }
            fence_proxy_async_smem();
            named_bar_sync(bar_id, 256);
            if (et == 0) {
                tma_store_3d(&a.tmOut, sQ, h * 64, q0, b);
                tma_store_commit();
            }
            if (tr) tr[7] = clock64();
        }
        if (et == 0 && my_items > 0) {
            tma_store_wait_read<0>();
            mbar_arrive(&stage_empty[(my_items - 1) & 1]);
            tma_store_wait_all<0>();
        }
    }
    __syncwarp();
    tc_fence_before();
    __syncthreads();
    if (warp == 17) {
        tc_fence_after();
        tmem_dealloc<512>(tmem);
    }
}

}  // namespace ddb
