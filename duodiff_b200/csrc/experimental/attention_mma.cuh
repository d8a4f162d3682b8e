// EXPERIMENTAL (compiled only with -DDDB_EXPERIMENTAL): generic-L attention on mma.sync.m16n8k16 with an online
// softmax.  No reference config reaches it (every U-ViT config has 256 patch tokens -> attention_tcgen05_kernel); it is
// kept as an independent second implementation for the operator tests.
#pragma once
#include "../attention.cuh"

namespace ddb {

// Generic-L kernel: one CTA per (sample, head), 8 warps, each warp owns 16-query-row blocks.
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
        AttRowState st;
        st.init();
        for (int kb0 = 0; kb0 < Lp; kb0 += 64)
            att_mma_block(sK_u, sV_u, kb0, min(8, (Lp - kb0) >> 3), L, qf, st, scale_log2e, lane);
        float l0 = st.l0, l1 = st.l1;
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
            if (row0 < L) *reinterpret_cast<uint32_t*>(o0 + i * 8) = pack_bf16(st.o[i][0] * inv0, st.o[i][1] * inv0);
            if (row1 < L) *reinterpret_cast<uint32_t*>(o1 + i * 8) = pack_bf16(st.o[i][2] * inv1, st.o[i][3] * inv1);
        }
    }
}

}  // namespace ddb
