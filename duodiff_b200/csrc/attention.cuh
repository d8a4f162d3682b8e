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

}  // namespace ddb
