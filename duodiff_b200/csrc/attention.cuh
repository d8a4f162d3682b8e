// Fused non-causal attention for U-ViT (models/uvit.py:155-166): softmax(q k^T / sqrt(64)) v per (sample, head).
// Sequence length is 257/258 in every config and head_dim is 64, so K and V of one (sample, head) live in
// shared memory for the whole CTA and the score matrix never touches HBM.
//
// Input  qkv [B*L, 3*D] bf16, feature index = k*(H*64) + h*64 + d  (k in {q,k,v}; models/uvit.py:159-161)
// Output o   [B*L, D]   bf16, feature index = h*64 + d             (models/uvit.py:164)
//
// v1: mma.sync.m16n8k16 (bf16 -> fp32) with online softmax; one CTA per (sample, head), 8 warps, each warp
// owns 16-query-row blocks.
#pragma once
#include "ptx.cuh"

namespace ddb {

__device__ __forceinline__ void cp_async16(void* smem_dst, const void* gsrc) {
    asm volatile("cp.async.cg.shared.global [%0], [%1], 16;" ::"r"(smem_u32(smem_dst)), "l"(gsrc) : "memory");
}
__device__ __forceinline__ void cp_async_wait_all() { asm volatile("cp.async.wait_all;" ::: "memory"); }

__device__ __forceinline__ void ldmatrix_x4(uint32_t addr, uint32_t& r0, uint32_t& r1, uint32_t& r2, uint32_t& r3) {
    asm volatile("ldmatrix.sync.aligned.m8n8.x4.shared.b16 {%0,%1,%2,%3}, [%4];"
                 : "=r"(r0), "=r"(r1), "=r"(r2), "=r"(r3)
                 : "r"(addr));
}
__device__ __forceinline__ void ldmatrix_x4_trans(uint32_t addr, uint32_t& r0, uint32_t& r1, uint32_t& r2,
                                                  uint32_t& r3) {
    asm volatile("ldmatrix.sync.aligned.m8n8.x4.trans.shared.b16 {%0,%1,%2,%3}, [%4];"
                 : "=r"(r0), "=r"(r1), "=r"(r2), "=r"(r3)
                 : "r"(addr));
}
__device__ __forceinline__ void mma_bf16_16816(float (&c)[4], const uint32_t (&a)[4], uint32_t b0, uint32_t b1) {
    asm volatile(
        "mma.sync.aligned.m16n8k16.row.col.f32.bf16.bf16.f32 {%0,%1,%2,%3}, {%4,%5,%6,%7}, {%8,%9}, {%0,%1,%2,%3};"
        : "+f"(c[0]), "+f"(c[1]), "+f"(c[2]), "+f"(c[3])
        : "r"(a[0]), "r"(a[1]), "r"(a[2]), "r"(a[3]), "r"(b0), "r"(b1));
}

constexpr int ATT_THREADS = 256;

// smem row = one key (64 bf16 = 128 B), 16-byte chunks XOR-swizzled by (row & 7)
__device__ __forceinline__ uint32_t att_swz(int row, int chunk) { return row * 128 + ((chunk ^ (row & 7)) << 4); }

__global__ void __launch_bounds__(ATT_THREADS, 2) attention_mma_kernel(const __nv_bfloat16* __restrict__ qkv,
                                                                   __nv_bfloat16* __restrict__ out, int L, int H,
                                                                   float scale_log2e) {
    extern __shared__ __align__(128) uint8_t att_smem[];
    const int Lp = (L + 15) & ~15;
    uint8_t* sK = att_smem;
    uint8_t* sV = att_smem + Lp * 128;
    const int b = blockIdx.x / H, h = blockIdx.x % H;
    const int D = H * 64;
    const size_t row_stride = (size_t)3 * D;
    const __nv_bfloat16* base = qkv + (size_t)b * L * row_stride + h * 64;

    // ---- stage K and V (zero the padded keys)
    for (int i = threadIdx.x; i < Lp * 8; i += ATT_THREADS) {
        const int r = i >> 3, c = i & 7;
        if (r < L) {
            cp_async16(sK + att_swz(r, c), base + (size_t)r * row_stride + D + c * 8);
            cp_async16(sV + att_swz(r, c), base + (size_t)r * row_stride + 2 * D + c * 8);
        } else {
            *reinterpret_cast<uint4*>(sK + att_swz(r, c)) = make_uint4(0, 0, 0, 0);
            *reinterpret_cast<uint4*>(sV + att_swz(r, c)) = make_uint4(0, 0, 0, 0);
        }
    }
    cp_async_wait_all();
    __syncthreads();

    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    const int g = lane >> 2, t = lane & 3;
    const uint32_t sK_u = smem_u32(sK), sV_u = smem_u32(sV);
    const int num_qblk = (L + 15) >> 4;

    for (int qb = warp; qb < num_qblk; qb += ATT_THREADS / 32) {
        const int r0 = min(qb * 16 + g, L - 1), r1 = min(qb * 16 + g + 8, L - 1);
        // Q fragments for the 4 k-steps over head_dim
        uint32_t qf[4][4];
        {
            const __nv_bfloat16* q0 = base + (size_t)r0 * row_stride;
            const __nv_bfloat16* q1 = base + (size_t)r1 * row_stride;
#pragma unroll
            for (int ks = 0; ks < 4; ++ks) {
                qf[ks][0] = __ldg(reinterpret_cast<const uint32_t*>(q0 + ks * 16 + 2 * t));
                qf[ks][1] = __ldg(reinterpret_cast<const uint32_t*>(q1 + ks * 16 + 2 * t));
                qf[ks][2] = __ldg(reinterpret_cast<const uint32_t*>(q0 + ks * 16 + 8 + 2 * t));
                qf[ks][3] = __ldg(reinterpret_cast<const uint32_t*>(q1 + ks * 16 + 8 + 2 * t));
            }
        }
        float o[8][4];
#pragma unroll
        for (int i = 0; i < 8; ++i) o[i][0] = o[i][1] = o[i][2] = o[i][3] = 0.f;
        float m0 = -INFINITY, m1 = -INFINITY, l0 = 0.f, l1 = 0.f;

        for (int kb0 = 0; kb0 < Lp; kb0 += 64) {
            const int nt = min(8, (Lp - kb0) >> 3);  // 8-key n-tiles in this block (even)
            float s[8][4];
#pragma unroll
            for (int j = 0; j < 8; ++j) {
                s[j][0] = s[j][1] = s[j][2] = s[j][3] = 0.f;
                if (j < nt) {
                    const int key = kb0 + j * 8 + (lane & 7);
                    const int cs = lane >> 3;  // which 8x8 matrix this lane addresses (dim chunk)
                    uint32_t b0, b1, b2, b3;
                    ldmatrix_x4(sK_u + att_swz(key, cs), b0, b1, b2, b3);  // dims 0..31
                    mma_bf16_16816(s[j], qf[0], b0, b1);
                    mma_bf16_16816(s[j], qf[1], b2, b3);
                    ldmatrix_x4(sK_u + att_swz(key, cs + 4), b0, b1, b2, b3);  // dims 32..63
                    mma_bf16_16816(s[j], qf[2], b0, b1);
                    mma_bf16_16816(s[j], qf[3], b2, b3);
                }
            }
            // mask padded keys, block row max
            float bm0 = -INFINITY, bm1 = -INFINITY;
#pragma unroll
            for (int j = 0; j < 8; ++j) {
                if (j < nt) {
                    const int key = kb0 + j * 8 + 2 * t;
                    if (key >= L) s[j][0] = s[j][2] = -INFINITY;
                    if (key + 1 >= L) s[j][1] = s[j][3] = -INFINITY;
                    bm0 = fmaxf(bm0, fmaxf(s[j][0], s[j][1]));
                    bm1 = fmaxf(bm1, fmaxf(s[j][2], s[j][3]));
                }
            }
            bm0 = fmaxf(bm0, __shfl_xor_sync(0xffffffffu, bm0, 1));
            bm0 = fmaxf(bm0, __shfl_xor_sync(0xffffffffu, bm0, 2));
            bm1 = fmaxf(bm1, __shfl_xor_sync(0xffffffffu, bm1, 1));
            bm1 = fmaxf(bm1, __shfl_xor_sync(0xffffffffu, bm1, 2));
            const float nm0 = fmaxf(m0, bm0), nm1 = fmaxf(m1, bm1);  // finite: every block has a valid key
            const float corr0 = exp2f((m0 - nm0) * scale_log2e), corr1 = exp2f((m1 - nm1) * scale_log2e);
            m0 = nm0, m1 = nm1;
            const float ms0 = nm0 * scale_log2e, ms1 = nm1 * scale_log2e;
            float ps0 = 0.f, ps1 = 0.f;
            uint32_t pf[8][2];
#pragma unroll
            for (int j = 0; j < 8; ++j) {
                if (j < nt) {
                    const float p0 = exp2f(fmaf(s[j][0], scale_log2e, -ms0));
                    const float p1 = exp2f(fmaf(s[j][1], scale_log2e, -ms0));
                    const float p2 = exp2f(fmaf(s[j][2], scale_log2e, -ms1));
                    const float p3 = exp2f(fmaf(s[j][3], scale_log2e, -ms1));
                    ps0 += p0 + p1;
                    ps1 += p2 + p3;
                    pf[j][0] = pack_bf16(p0, p1);
                    pf[j][1] = pack_bf16(p2, p3);
                } else {
                    pf[j][0] = pf[j][1] = 0u;
                }
            }
            l0 = l0 * corr0 + ps0;
            l1 = l1 * corr1 + ps1;
#pragma unroll
            for (int i = 0; i < 8; ++i) {
                o[i][0] *= corr0, o[i][1] *= corr0;
                o[i][2] *= corr1, o[i][3] *= corr1;
            }
            // O += P V  (k-steps of 16 keys)
#pragma unroll
            for (int kk = 0; kk < 4; ++kk) {
                if (2 * kk < nt) {
                    const uint32_t pa[4] = {pf[2 * kk][0], pf[2 * kk][1], pf[2 * kk + 1][0], pf[2 * kk + 1][1]};
                    // ldmatrix.trans: matrices (keys 0-7, dims d..d+7), (keys 8-15, same dims), then dims +8
                    const int key = kb0 + kk * 16 + (lane & 7) + ((lane >> 3) & 1) * 8;
                    const int dsel = lane >> 4;  // 0/1 -> dim chunk offset
#pragma unroll
                    for (int dn = 0; dn < 4; ++dn) {
                        uint32_t v0, v1, v2, v3;
                        ldmatrix_x4_trans(sV_u + att_swz(key, dn * 2 + dsel), v0, v1, v2, v3);
                        mma_bf16_16816(o[dn * 2], pa, v0, v1);
                        mma_bf16_16816(o[dn * 2 + 1], pa, v2, v3);
                    }
                }
            }
        }
        l0 += __shfl_xor_sync(0xffffffffu, l0, 1);
        l0 += __shfl_xor_sync(0xffffffffu, l0, 2);
        l1 += __shfl_xor_sync(0xffffffffu, l1, 1);
        l1 += __shfl_xor_sync(0xffffffffu, l1, 2);
        const float inv0 = 1.f / l0, inv1 = 1.f / l1;
        const int row0 = qb * 16 + g, row1 = row0 + 8;
        __nv_bfloat16* o0 = out + ((size_t)b * L + row0) * D + h * 64 + 2 * t;
        __nv_bfloat16* o1 = out + ((size_t)b * L + row1) * D + h * 64 + 2 * t;
#pragma unroll
        for (int i = 0; i < 8; ++i) {
            if (row0 < L) *reinterpret_cast<uint32_t*>(o0 + i * 8) = pack_bf16(o[i][0] * inv0, o[i][1] * inv0);
            if (row1 < L) *reinterpret_cast<uint32_t*>(o1 + i * 8) = pack_bf16(o[i][2] * inv1, o[i][3] * inv1);
        }
    }
}


// =====================================================================================================
// v2: tcgen05 / TMEM attention for the 256-patch-token layout of every reference config (L = 256 + extras).
//
// One CTA per (sample, head, 128-query tile); 2 CTAs co-reside per SM (80 KB smem, 256 TMEM columns each) so the
// softmax of one overlaps the MMAs / TMA loads of the other.
//   keys   [extras, L)  (the 256 patch tokens): S = Q K^T on the tensor core, M=128 x N=256 x K=64, fp32 in TMEM
//   keys   [0, extras)  (time / label tokens):  1-2 dot products per query row on the CUDA cores
//   softmax: one thread per query row reads its S row from TMEM (no shuffles), writes P (bf16, packed) back over
//            the first 128 columns of S; O = P V accumulates in columns [128,192) of the same allocation with
//            A = P from TMEM and B = V (MN-major, 128B swizzle) from shared memory.
// The query rows [0, extras) are handled by attention_extras_kernel below.
// =====================================================================================================
__device__ __forceinline__ float ex2_approx(float x) {
    float y;
    asm("ex2.approx.ftz.f32 %0, %1;" : "=f"(y) : "f"(x));
    return y;
}

struct AttnArgs {
    CUtensorMap tmQKV;  // [B, L, 3D] bf16, box {64, 128, 1}
    CUtensorMap tmKV;   // [B, L, 3D] bf16, box {64, 256, 1}
    CUtensorMap tmOut;  // [B, L, D]  bf16, box {64, 128, 1}
    const __nv_bfloat16* qkv;
    int L, H, extras;
    float scale_log2e;
    const int* b_dev;  // optional live batch size (early-exit compaction)
};

constexpr int ATT2_THREADS = 192;
constexpr int ATT2_SMEM = 16384 + 32768 + 32768 + 1024 + 128;

__global__ void __launch_bounds__(ATT2_THREADS, 2) attention_tcgen05_kernel(const __grid_constant__ AttnArgs a) {
    extern __shared__ uint8_t att2_raw[];
    uint8_t* smem = reinterpret_cast<uint8_t*>((reinterpret_cast<uintptr_t>(att2_raw) + 1023) & ~uintptr_t(1023));
    uint8_t* sQ = smem;           // 128 x 128 B (later reused as the output staging tile)
    uint8_t* sK = smem + 16384;   // 256 x 128 B
    uint8_t* sV = smem + 49152;   // 256 x 128 B
    uint64_t* bars = reinterpret_cast<uint64_t*>(smem + 81920);
    uint64_t* qk_full = bars + 0;
    uint64_t* v_full = bars + 1;
    uint64_t* s_full = bars + 2;
    uint64_t* p_full = bars + 3;
    uint64_t* o_full = bars + 4;
    uint32_t* tmem_holder = reinterpret_cast<uint32_t*>(bars + 5);

    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    const int tile = blockIdx.x & 1;
    const int bh = blockIdx.x >> 1;
    const int b = bh / a.H, h = bh % a.H;
    if (a.b_dev && b >= *a.b_dev) return;
    const int D = a.H * 64;
    const int q0 = a.extras + tile * 128;

    if (warp == 0 && lane == 0) {
        tma_prefetch_desc(&a.tmQKV);
        tma_prefetch_desc(&a.tmKV);
        tma_prefetch_desc(&a.tmOut);
        mbar_init(qk_full, 1);
        mbar_init(v_full, 1);
        mbar_init(s_full, 1);
        mbar_init(p_full, 4);
        mbar_init(o_full, 1);
        fence_mbar_init();
    }
    if (warp == 1) tmem_alloc<256>(tmem_holder);
    tc_fence_before();
    __syncthreads();
    tc_fence_after();
    const uint32_t tmem = *tmem_holder;

    if (warp == 0) {
        if (lane == 0) {
            mbar_expect_tx(qk_full, 16384 + 32768);
            tma_load_3d(sQ, &a.tmQKV, qk_full, h * 64, q0, b);
            tma_load_3d(sK, &a.tmKV, qk_full, D + h * 64, a.extras, b);
            mbar_expect_tx(v_full, 32768);
            tma_load_3d(sV, &a.tmKV, v_full, 2 * D + h * 64, a.extras, b);
        }
    } else if (warp == 1) {
        if (lane == 0) {
            mbar_wait(qk_full, 0);
            tc_fence_after();
            const uint64_t dq = umma_desc_kmajor_sw128(smem_u32(sQ));
            const uint64_t dk = umma_desc_kmajor_sw128(smem_u32(sK));
            constexpr uint32_t idesc_s = umma_idesc_bf16(128, 256);
#pragma unroll
            for (int k = 0; k < 4; ++k) umma_f16_ss(tmem, dq + 2 * k, dk + 2 * k, idesc_s, k != 0);
            umma_commit(s_full);
            mbar_wait(p_full, 0);
            mbar_wait(v_full, 0);
            tc_fence_after();
            const uint64_t dv = umma_desc_mnmajor_sw128(smem_u32(sV));
            constexpr uint32_t idesc_o = umma_idesc_bf16(128, 64, 0, 1);
#pragma unroll
            for (int k = 0; k < 16; ++k)  // 16 keys per step: P columns +8, V rows +16 (2048 B)
                umma_f16_ts(tmem + 128, tmem + 8 * k, dv + (uint64_t)(k * (2048 >> 4)), idesc_o, k != 0);
            umma_commit(o_full);
        }
    } else {
        // ================================================================= softmax + epilogue: one thread per query row
        const int quarter = warp & 3;
        const int r = quarter * 32 + lane;
        const uint32_t t_row = tmem + (uint32_t(quarter * 32) << 16);
        const float c = a.scale_log2e;
        const __nv_bfloat16* xrow = a.qkv + (size_t)b * a.L * 3 * D;  // extras tokens of this sample

        // scores against the extras keys on the CUDA cores (overlaps the Q K^T MMA)
        mbar_wait(qk_full, 0);
        float se[2] = {-INFINITY, -INFINITY};
        {
            float acc0 = 0.f, acc1 = 0.f;
#pragma unroll
            for (int j = 0; j < 8; ++j) {
                const uint4 q = *reinterpret_cast<const uint4*>(sQ + r * 128 + ((j ^ (r & 7)) << 4));
                const float qf[8] = {bf16_lo(q.x), bf16_hi(q.x), bf16_lo(q.y), bf16_hi(q.y),
                                     bf16_lo(q.z), bf16_hi(q.z), bf16_lo(q.w), bf16_hi(q.w)};
                const uint4 k0 = __ldg(reinterpret_cast<const uint4*>(xrow + D + h * 64) + j);
                const float kf[8] = {bf16_lo(k0.x), bf16_hi(k0.x), bf16_lo(k0.y), bf16_hi(k0.y),
                                     bf16_lo(k0.z), bf16_hi(k0.z), bf16_lo(k0.w), bf16_hi(k0.w)};
#pragma unroll
                for (int e = 0; e < 8; ++e) acc0 = fmaf(qf[e], kf[e], acc0);
                if (a.extras == 2) {
                    const uint4 k1 = __ldg(reinterpret_cast<const uint4*>(xrow + 3 * D + D + h * 64) + j);
                    const float kg[8] = {bf16_lo(k1.x), bf16_hi(k1.x), bf16_lo(k1.y), bf16_hi(k1.y),
                                         bf16_lo(k1.z), bf16_hi(k1.z), bf16_lo(k1.w), bf16_hi(k1.w)};
#pragma unroll
                    for (int e = 0; e < 8; ++e) acc1 = fmaf(qf[e], kg[e], acc1);
                }
            }
            se[0] = acc0;
            if (a.extras == 2) se[1] = acc1;
        }

        mbar_wait(s_full, 0);
        tc_fence_after();
        // pass 1: row max
        float m = fmaxf(se[0], se[1]);
#pragma unroll 1
        for (int j = 0; j < 8; j += 2) {
            uint32_t v0[32], v1[32];
            tmem_ld_32x32b_x32(t_row + j * 32, v0);
            tmem_ld_32x32b_x32(t_row + j * 32 + 32, v1);
            tmem_ld_wait();
#pragma unroll
            for (int e = 0; e < 32; ++e) m = fmaxf(m, fmaxf(__uint_as_float(v0[e]), __uint_as_float(v1[e])));
        }
        const float mc = m * c;
        float sum = ex2_approx(fmaf(se[0], c, -mc));
        float pe[2] = {sum, 0.f};
        if (a.extras == 2) {
            pe[1] = ex2_approx(fmaf(se[1], c, -mc));
            sum += pe[1];
        }
        // pass 2: P = exp2(s*c - m*c) -> bf16, written over the already-consumed S columns
#pragma unroll 1
        for (int j = 0; j < 8; ++j) {
            uint32_t v[32];
            tmem_ld_32x32b_x32(t_row + j * 32, v);
            tmem_ld_wait();
            uint32_t pk[16];
#pragma unroll
            for (int e = 0; e < 16; ++e) {
                const float p0 = ex2_approx(fmaf(__uint_as_float(v[2 * e]), c, -mc));
                const float p1 = ex2_approx(fmaf(__uint_as_float(v[2 * e + 1]), c, -mc));
                sum += p0 + p1;
                pk[e] = pack_bf16(p0, p1);
            }
            tmem_st_32x32b_x16(t_row + j * 16, pk);
        }
        tmem_st_wait();
        tc_fence_before();
        __syncwarp();
        if (lane == 0) mbar_arrive(p_full);

        // epilogue: O row (fp32) + extras-key contributions, normalise, bf16 -> smem (Q tile is dead) -> TMA store
        const float inv = 1.f / sum;
        mbar_wait(o_full, 0);
        tc_fence_after();
        uint32_t o0[32], o1[32];
        tmem_ld_32x32b_x32(t_row + 128, o0);
        tmem_ld_32x32b_x32(t_row + 160, o1);
        tmem_ld_wait();
        uint8_t* srow = sQ + r * 128;
#pragma unroll
        for (int j = 0; j < 8; ++j) {
            float o[8];
#pragma unroll
            for (int e = 0; e < 8; ++e) o[e] = __uint_as_float(j < 4 ? o0[j * 8 + e] : o1[(j - 4) * 8 + e]);
            const uint4 va = __ldg(reinterpret_cast<const uint4*>(xrow + 2 * D + h * 64) + j);
            const float vf[8] = {bf16_lo(va.x), bf16_hi(va.x), bf16_lo(va.y), bf16_hi(va.y),
                                 bf16_lo(va.z), bf16_hi(va.z), bf16_lo(va.w), bf16_hi(va.w)};
#pragma unroll
            for (int e = 0; e < 8; ++e) o[e] = fmaf(pe[0], vf[e], o[e]);
            if (a.extras == 2) {
                const uint4 vb = __ldg(reinterpret_cast<const uint4*>(xrow + 3 * D + 2 * D + h * 64) + j);
                const float vg[8] = {bf16_lo(vb.x), bf16_hi(vb.x), bf16_lo(vb.y), bf16_hi(vb.y),
                                     bf16_lo(vb.z), bf16_hi(vb.z), bf16_lo(vb.w), bf16_hi(vb.w)};
#pragma unroll
                for (int e = 0; e < 8; ++e) o[e] = fmaf(pe[1], vg[e], o[e]);
            }
            uint4 w;
            w.x = pack_bf16(o[0] * inv, o[1] * inv);
            w.y = pack_bf16(o[2] * inv, o[3] * inv);
            w.z = pack_bf16(o[4] * inv, o[5] * inv);
            w.w = pack_bf16(o[6] * inv, o[7] * inv);
            *reinterpret_cast<uint4*>(srow + ((j ^ (r & 7)) << 4)) = w;
        }
        fence_proxy_async_smem();
        named_bar_sync(1, 128);
        if (warp == 2 && lane == 0) {
            tma_store_3d(&a.tmOut, sQ, h * 64, q0, b);
            tma_store_commit();
            tma_store_wait_all<0>();
        }
    }
    tc_fence_before();
    __syncthreads();
    if (warp == 1) {
        tc_fence_after();
        tmem_dealloc<256>(tmem);
    }
}

// Query rows [0, extras) (time / label tokens) of every (sample, head): one warp per (sample, head, row).
// Lane l scores keys l, l+32, ... (16-byte loads of the key rows), warp-shuffle softmax, then each lane owns two
// output dims and streams V with coalesced 128-byte warp loads.  grid = ceil(B*H*extras / 4), 128 threads.
__global__ void __launch_bounds__(128) attention_extras_kernel(const __nv_bfloat16* __restrict__ qkv,
                                                               __nv_bfloat16* __restrict__ out, int L, int H,
                                                               int extras, float scale_log2e, int B,
                                                               const int* __restrict__ b_dev) {
    const int lane = threadIdx.x & 31;
    const int wid = blockIdx.x * 4 + (threadIdx.x >> 5);
    const int Bl = b_dev ? *b_dev : B;
    if (wid >= Bl * H * extras) return;
    const int e = wid % extras, bh = wid / extras;
    const int b = bh / H, h = bh % H;
    const int D = H * 64;
    const size_t rs = (size_t)3 * D;
    const __nv_bfloat16* base = qkv + (size_t)b * L * rs + h * 64;
    // q row (64 dims) in registers, identical in every lane
    float q[64];
    {
        const uint4* qr = reinterpret_cast<const uint4*>(base + (size_t)e * rs);
#pragma unroll
        for (int j = 0; j < 8; ++j) {
            const uint4 u = __ldg(qr + j);
            q[j * 8 + 0] = bf16_lo(u.x), q[j * 8 + 1] = bf16_hi(u.x), q[j * 8 + 2] = bf16_lo(u.y);
            q[j * 8 + 3] = bf16_hi(u.y), q[j * 8 + 4] = bf16_lo(u.z), q[j * 8 + 5] = bf16_hi(u.z);
            q[j * 8 + 6] = bf16_lo(u.w), q[j * 8 + 7] = bf16_hi(u.w);
        }
    }
    constexpr int KPL = 9;  // keys per lane: covers L <= 288
    float sc[KPL];
    float mx = -INFINITY;
#pragma unroll
    for (int i = 0; i < KPL; ++i) {
        const int k = i * 32 + lane;
        float s = -INFINITY;
        if (k < L) {
            const uint4* kr = reinterpret_cast<const uint4*>(base + (size_t)k * rs + D);
            s = 0.f;
#pragma unroll
            for (int j = 0; j < 8; ++j) {
                const uint4 u = __ldg(kr + j);
                s = fmaf(q[j * 8 + 0], bf16_lo(u.x), s), s = fmaf(q[j * 8 + 1], bf16_hi(u.x), s);
                s = fmaf(q[j * 8 + 2], bf16_lo(u.y), s), s = fmaf(q[j * 8 + 3], bf16_hi(u.y), s);
                s = fmaf(q[j * 8 + 4], bf16_lo(u.z), s), s = fmaf(q[j * 8 + 5], bf16_hi(u.z), s);
                s = fmaf(q[j * 8 + 6], bf16_lo(u.w), s), s = fmaf(q[j * 8 + 7], bf16_hi(u.w), s);
            }
        }
        sc[i] = s;
        mx = fmaxf(mx, s);
    }
    for (int o = 16; o > 0; o >>= 1) mx = fmaxf(mx, __shfl_xor_sync(0xffffffffu, mx, o));
    float sum = 0.f;
#pragma unroll
    for (int i = 0; i < KPL; ++i) {
        sc[i] = exp2f((sc[i] - mx) * scale_log2e);  // exp2(-inf) = 0 for k >= L
        sum += sc[i];
    }
    for (int o = 16; o > 0; o >>= 1) sum += __shfl_xor_sync(0xffffffffu, sum, o);
    // O[2*lane, 2*lane+1] = sum_k p_k V[k][.]
    float o0 = 0.f, o1 = 0.f;
    const __nv_bfloat16* vbase = base + 2 * D + 2 * lane;
#pragma unroll
    for (int i = 0; i < KPL; ++i) {
        const int kmax = min(32, L - i * 32);
#pragma unroll 8
        for (int src = 0; src < 32; ++src) {
            const float p = __shfl_sync(0xffffffffu, sc[i], src);
            if (src < kmax) {
                const uint32_t v = __ldg(reinterpret_cast<const uint32_t*>(vbase + (size_t)(i * 32 + src) * rs));
                o0 = fmaf(p, bf16_lo(v), o0);
                o1 = fmaf(p, bf16_hi(v), o1);
            }
        }
    }
    const float inv = 1.f / sum;
    *reinterpret_cast<uint32_t*>(out + ((size_t)b * L + e) * D + h * 64 + 2 * lane) = pack_bf16(o0 * inv, o1 * inv);
}

}  // namespace ddb
