// Fused non-causal attention for U-ViT (models/uvit.py:155-166): softmax(q k^T / sqrt(64)) v per (sample, head).
// Sequence length is 257/258 in every config and head_dim is 64, so K and V of one (sample, head) live in
// shared memory for the whole work item and the score matrix never touches HBM.
//
// Input  qkv [B*L, 3*D] bf16, feature index = k*(H*64) + h*64 + d  (k in {q,k,v}; models/uvit.py:159-161)
// Output o   [B*L, D]   bf16, feature index = h*64 + d             (models/uvit.py:164)
//
// attention_tcgen05_kernel is the model path (L = 256 + extras): persistent, warp-specialised, tcgen05 / TMEM.  The
// mma.sync block below (att_mma_block) serves its extras query rows; a generic-L kernel built from the same block lives
// in experimental/attention_mma.cuh (DDB_EXPERIMENTAL builds only: no reference config reaches it).
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

__device__ __forceinline__ float ex2_approx(float x) {
    float y;
    asm("ex2.approx.ftz.f32 %0, %1;" : "=f"(y) : "f"(x));
    return y;
}
__device__ __forceinline__ float fmax3(float a, float b, float c) {
    float d;
    asm("max.f32 %0, %1, %2, %3;" : "=f"(d) : "f"(a), "f"(b), "f"(c));
    return d;
}

constexpr int ATT_THREADS = 256;

// smem row = one key (64 bf16 = 128 B), 16-byte chunks XOR-swizzled by (row & 7): the TMA SWIZZLE_128B pattern
__device__ __forceinline__ uint32_t att_swz(int row, int chunk) { return row * 128 + ((chunk ^ (row & 7)) << 4); }

// Online-softmax state of one warp's 16-query-row block (mma.sync fragment layout: lane = 4*g + t owns rows g, g+8).
struct AttRowState {
    float o[8][4];
    float m0, m1, l0, l1;
    __device__ __forceinline__ void init() {
#pragma unroll
        for (int i = 0; i < 8; ++i) o[i][0] = o[i][1] = o[i][2] = o[i][3] = 0.f;
        m0 = m1 = -INFINITY;
        l0 = l1 = 0.f;
    }
};

// One block of up to 64 keys: S = Q K^T, online softmax, O += P V.  sK/sV: swizzled [key][64] bf16 tiles (shared
// addresses) indexed by absolute key; keys >= kend are masked.  nt = number of 8-key n-tiles in the block (V must be
// readable for an even number of them).
// TOP_ONLY: rows g+8 of the block are padding (all-zero queries): their softmax math is skipped.
template <bool TOP_ONLY = false>
__device__ __forceinline__ void att_mma_block(uint32_t sK_u, uint32_t sV_u, int kb0, int nt, int kend,
                                              const uint32_t (&qf)[4][4], AttRowState& st, float scale_log2e,
                                              int lane) {
    const int t = lane & 3;
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
            if (key >= kend) s[j][0] = s[j][2] = -INFINITY;
            if (key + 1 >= kend) s[j][1] = s[j][3] = -INFINITY;
            bm0 = fmaxf(bm0, fmaxf(s[j][0], s[j][1]));
            if (!TOP_ONLY) bm1 = fmaxf(bm1, fmaxf(s[j][2], s[j][3]));
        }
    }
    bm0 = fmaxf(bm0, __shfl_xor_sync(0xffffffffu, bm0, 1));
    bm0 = fmaxf(bm0, __shfl_xor_sync(0xffffffffu, bm0, 2));
    if (!TOP_ONLY) {
        bm1 = fmaxf(bm1, __shfl_xor_sync(0xffffffffu, bm1, 1));
        bm1 = fmaxf(bm1, __shfl_xor_sync(0xffffffffu, bm1, 2));
    }
    const float nm0 = fmaxf(st.m0, bm0), nm1 = TOP_ONLY ? 0.f : fmaxf(st.m1, bm1);  // finite: a valid key per block
    const float corr0 = exp2f((st.m0 - nm0) * scale_log2e);
    const float corr1 = TOP_ONLY ? 1.f : exp2f((st.m1 - nm1) * scale_log2e);
    st.m0 = nm0, st.m1 = nm1;
    const float ms0 = nm0 * scale_log2e, ms1 = nm1 * scale_log2e;
    float ps0 = 0.f, ps1 = 0.f;
    uint32_t pf[8][2];
#pragma unroll
    for (int j = 0; j < 8; ++j) {
        if (j < nt) {
            const float p0 = exp2f(fmaf(s[j][0], scale_log2e, -ms0));
            const float p1 = exp2f(fmaf(s[j][1], scale_log2e, -ms0));
            ps0 += p0 + p1;
            pf[j][0] = pack_bf16(p0, p1);
            if (!TOP_ONLY) {
                const float p2 = exp2f(fmaf(s[j][2], scale_log2e, -ms1));
                const float p3 = exp2f(fmaf(s[j][3], scale_log2e, -ms1));
                ps1 += p2 + p3;
                pf[j][1] = pack_bf16(p2, p3);
            } else {
                pf[j][1] = 0u;
            }
        } else {
            pf[j][0] = pf[j][1] = 0u;
        }
    }
    st.l0 = st.l0 * corr0 + ps0;
    if (!TOP_ONLY) st.l1 = st.l1 * corr1 + ps1;
#pragma unroll
    for (int i = 0; i < 8; ++i) {
        st.o[i][0] *= corr0, st.o[i][1] *= corr0;
        if (!TOP_ONLY) st.o[i][2] *= corr1, st.o[i][3] *= corr1;
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
                mma_bf16_16816(st.o[dn * 2], pa, v0, v1);
                mma_bf16_16816(st.o[dn * 2 + 1], pa, v2, v3);
            }
        }
    }
}

// =====================================================================================================
// Persistent, warp-specialised tcgen05 / TMEM attention for the 256-patch-token layout of every reference
// config (L = 256 + extras, extras = 1 time token (+ 1 label token)).
//
// One CTA per SM; a work item is one (sample, head): K and V are TMA-loaded ONCE per item and shared by the two
// 128-query tiles and the extras query rows; operands are double-buffered across items, and the tensor-core work
// of one tile overlaps the softmax of the other:
//   warps 0-3  softmax warpgroup of query tile 0 (one thread per query row, no shuffles)
//   warps 4-7  softmax warpgroup of query tile 1
//   warp  8    TMA producer: {Q0, Q1, K, V, X} of item i+1 while item i is being processed (2 x 102 KB stages);
//              X = tokens 0..15 of the sample (q, k and v slices): the extras tokens plus don't-care patch tokens
//   warp  9    MMA issuer of tile 0 + TMEM owner (512 columns = 2 tiles x 256)
//   warp  10   extras QUERY rows (tokens [0, extras)) on mma.sync from the same shared-memory K / V tiles
//   warp  11   MMA issuer of tile 1 (starts half a period late: the two softmax warpgroups then run out of phase
//              and keep the MUFU pipe -- the binding unit, 256 ex2 per query row -- busy)
// Per tile t (TMEM columns relative to 256 t):
//   S_t = Q_t K^T          M=128, N=256 (patch keys), K=64; fp32 in [0, 256)
//   extras KEY scores      1-2 dot products per query row on the CUDA cores (k rows read from the X tile)
//   softmax                pass 1 row max (FMNMX3), pass 2 P = exp2(s*c - m*c) -> bf16 written back over consumed S
//                          columns: keys [0,128) -> [0, 64), keys [128,256) -> [128, 192) (each TMEM load is issued
//                          one chunk ahead of the math on the previous chunk); the extras keys' probabilities go to
//                          [192, 200) as a 17th k-step of 16 keys (14-15 of them zero)
//   O_t = P_t [V; V_x]     A = P from TMEM, B = V (MN-major) from smem, 17 k-steps, fp32 in [64, 128); the first 8
//                          k-steps are issued as soon as keys [0,128) are done and run under the rest of pass 2
//   epilogue               O row / sum -> bf16 -> the tile's own (dead) Q buffer -> TMA store
// =====================================================================================================

// The extras QUERY rows (A rows g < 2 of the Qx fragments; rows g+8 are zero) against 32 consecutive, all-valid patch
// keys starting at kb0: ONE softmax block with the fragment traffic of att_mma_block<true>, without masks or running
// state.  Returns the block's row max (raw score units), row sum (reduced over the quad) and un-normalised output
// (lane (g, t): dims 8i + 2t, 8i + 2t + 1 of row g).
__device__ __forceinline__ void att_extras_slice32(uint32_t sK_u, uint32_t sV_u, int kb0, const uint32_t (&qf)[4][4],
                                                   float scale_log2e, int lane, float& m_out, float& l_out,
                                                   float (&o)[8][2]) {
    const int t = lane & 3;
    (void)t;
    float s[4][4];
#pragma unroll
    for (int j = 0; j < 4; ++j) {
        s[j][0] = s[j][1] = s[j][2] = s[j][3] = 0.f;
        const int key = kb0 + j * 8 + (lane & 7);
        const int cs = lane >> 3;
        uint32_t b0, b1, b2, b3;
        ldmatrix_x4(sK_u + att_swz(key, cs), b0, b1, b2, b3);  // dims 0..31
        mma_bf16_16816(s[j], qf[0], b0, b1);
        mma_bf16_16816(s[j], qf[1], b2, b3);
        ldmatrix_x4(sK_u + att_swz(key, cs + 4), b0, b1, b2, b3);  // dims 32..63
        mma_bf16_16816(s[j], qf[2], b0, b1);
        mma_bf16_16816(s[j], qf[3], b2, b3);
    }
    float m = fmaxf(fmax3(s[0][0], s[0][1], s[1][0]), fmax3(s[1][1], s[2][0], s[2][1]));
    m = fmax3(m, s[3][0], s[3][1]);
    m = fmaxf(m, __shfl_xor_sync(0xffffffffu, m, 1));
    m = fmaxf(m, __shfl_xor_sync(0xffffffffu, m, 2));
    const float ms = m * scale_log2e;
    float l = 0.f;
    uint32_t pf[4];
#pragma unroll
    for (int j = 0; j < 4; ++j) {
        const float p0 = ex2_approx(fmaf(s[j][0], scale_log2e, -ms));
        const float p1 = ex2_approx(fmaf(s[j][1], scale_log2e, -ms));
        l += p0 + p1;
        pf[j] = pack_bf16(p0, p1);
    }
    float oacc[8][4];
#pragma unroll
    for (int i = 0; i < 8; ++i) oacc[i][0] = oacc[i][1] = oacc[i][2] = oacc[i][3] = 0.f;
#pragma unroll
    for (int kk = 0; kk < 2; ++kk) {
        const uint32_t pa[4] = {pf[2 * kk], 0u, pf[2 * kk + 1], 0u};
        const int key = kb0 + kk * 16 + (lane & 7) + ((lane >> 3) & 1) * 8;
        const int dsel = lane >> 4;
#pragma unroll
        for (int dn = 0; dn < 4; ++dn) {
            uint32_t v0, v1, v2, v3;
            ldmatrix_x4_trans(sV_u + att_swz(key, dn * 2 + dsel), v0, v1, v2, v3);
            mma_bf16_16816(oacc[dn * 2], pa, v0, v1);
            mma_bf16_16816(oacc[dn * 2 + 1], pa, v2, v3);
        }
    }
    l += __shfl_xor_sync(0xffffffffu, l, 1);
    l += __shfl_xor_sync(0xffffffffu, l, 2);
    m_out = m, l_out = l;
#pragma unroll
    for (int i = 0; i < 8; ++i) o[i][0] = oacc[i][0], o[i][1] = oacc[i][1];
}

__device__ __forceinline__ uint32_t movmatrix_trans(uint32_t a) {
    uint32_t d;
    asm volatile("movmatrix.sync.aligned.m8n8.trans.b16 %0, %1;" : "=r"(d) : "r"(a));
    return d;
}
// The same slice with the KEYS as the M dimension of the mma (legacy HMMA costs ~27 clk per instruction and scheduler on
// sm_100, so the count matters): S^T[32 keys x 8 queries] = K_slice Qx^T (2 m-tiles x 4 k-steps = 8 HMMA instead of 16),
// softmax down the key axis (reductions across the lanes that share t), O^T[64 dims x 8 queries] = V^T P^T (4 m-tiles x
// 2 k-steps = 8 HMMA instead of 16); P^T reaches its B-fragment layout through movmatrix.  Queries = tokens 0..7 of the
// sample, of which [0, extras) are real; the columns of the padding queries are computed and ignored.
// Lane (g, t) returns for queries 2t (index 0) and 2t+1 (index 1): row max, row sum, and o[dm][2h + q] =
// O^T[dim 16 dm + 8 h + g][query 2t + q].  Only the t == 0 lanes hold the extras queries.
__device__ __forceinline__ void att_extras_slice32_t(uint32_t sK_u, uint32_t sV_u, uint32_t sQx_u, int kb0,
                                                     float scale_log2e, int lane, float (&m_out)[2], float (&l_out)[2],
                                                     float (&o)[4][4]) {
    uint32_t bq[4][2];  // B fragments of Qx^T: k-step ks -> (dims 16ks..+7, +8..+15) x queries 0..7
    ldmatrix_x4(sQx_u + att_swz(lane & 7, lane >> 3), bq[0][0], bq[0][1], bq[1][0], bq[1][1]);
    ldmatrix_x4(sQx_u + att_swz(lane & 7, (lane >> 3) + 4), bq[2][0], bq[2][1], bq[3][0], bq[3][1]);
    float s[2][4];
#pragma unroll
    for (int mt = 0; mt < 2; ++mt) {
        s[mt][0] = s[mt][1] = s[mt][2] = s[mt][3] = 0.f;
        const int row = kb0 + mt * 16 + (lane & 7) + ((lane >> 3) & 1) * 8;
#pragma unroll
        for (int ks = 0; ks < 4; ++ks) {
            uint32_t af[4];
            ldmatrix_x4(sK_u + att_swz(row, ks * 2 + (lane >> 4)), af[0], af[1], af[2], af[3]);
            mma_bf16_16816(s[mt], af, bq[ks][0], bq[ks][1]);
        }
    }
    // s[mt] = {(key g, q 2t), (key g, q 2t+1), (key g+8, q 2t), (key g+8, q 2t+1)} of key tile mt
    float ma = fmaxf(fmaxf(s[0][0], s[0][2]), fmaxf(s[1][0], s[1][2]));
    float mb = fmaxf(fmaxf(s[0][1], s[0][3]), fmaxf(s[1][1], s[1][3]));
#pragma unroll
    for (int sh = 4; sh < 32; sh <<= 1) {
        ma = fmaxf(ma, __shfl_xor_sync(0xffffffffu, ma, sh));
        mb = fmaxf(mb, __shfl_xor_sync(0xffffffffu, mb, sh));
    }
    const float msa = ma * scale_log2e, msb = mb * scale_log2e;
    float la = 0.f, lb = 0.f;
    uint32_t pb[2][2];  // B fragments of P^T per key tile: keys 0-7 / 8-15
#pragma unroll
    for (int mt = 0; mt < 2; ++mt) {
        const float p0 = ex2_approx(fmaf(s[mt][0], scale_log2e, -msa)), p1 = ex2_approx(fmaf(s[mt][1], scale_log2e, -msb));
        const float p2 = ex2_approx(fmaf(s[mt][2], scale_log2e, -msa)), p3 = ex2_approx(fmaf(s[mt][3], scale_log2e, -msb));
        la += p0 + p2;
        lb += p1 + p3;
        pb[mt][0] = movmatrix_trans(pack_bf16(p0, p1));
        pb[mt][1] = movmatrix_trans(pack_bf16(p2, p3));
    }
#pragma unroll
    for (int sh = 4; sh < 32; sh <<= 1) {
        la += __shfl_xor_sync(0xffffffffu, la, sh);
        lb += __shfl_xor_sync(0xffffffffu, lb, sh);
    }
#pragma unroll
    for (int dm = 0; dm < 4; ++dm) o[dm][0] = o[dm][1] = o[dm][2] = o[dm][3] = 0.f;
#pragma unroll
    for (int kk = 0; kk < 2; ++kk) {
        const int key = kb0 + kk * 16 + (lane & 7) + ((lane >> 4) & 1) * 8;
#pragma unroll
        for (int dm = 0; dm < 4; ++dm) {
            uint32_t av[4];  // V^T fragments: (dims 0-7 | 8-15) x (keys 0-7 | 8-15) of dim tile dm
            ldmatrix_x4_trans(sV_u + att_swz(key, dm * 2 + ((lane >> 3) & 1)), av[0], av[1], av[2], av[3]);
            mma_bf16_16816(o[dm], av, pb[kk][0], pb[kk][1]);
        }
    }
    m_out[0] = ma, m_out[1] = mb, l_out[0] = la, l_out[1] = lb;
}

struct AttnArgs {
    CUtensorMap tmQKV;    // [B, L, 3D] bf16, box {64, 128, 1}
    CUtensorMap tmKV;     // [B, L, 3D] bf16, box {64, 256, 1}
    CUtensorMap tmX;      // [B, L, 3D] bf16, box {64, 16, 1}
    CUtensorMap tmOut;    // [B, L, D]  bf16, box {64, 128, 1}
    CUtensorMap tmOut32;  // [B, L, D]  bf16, box {64, 32, 1}: one epilogue warp's rows
    const __nv_bfloat16* qkv;
    __nv_bfloat16* out;
    int L, H, extras, B;
    float scale_log2e;
    const int* b_dev;  // optional live batch size (early-exit compaction)
    int reverse;       // walk the (sample, head) items from the last to the first (see GemmArgs::reverse)
    int discard;       // drop the consumed q|k|v lines from L2 instead of letting them be written back (model path only)
    int token;         // the two query tiles take turns in the exp pass (one MUFU pipe per SM: see the kernel header)
    int direct_store;  // epilogue writes the output rows straight from registers instead of staging + TMA store
    long long* trace;  // bench-only: CTA 0 records clock64() at the phase boundaries of every item ([it][tile][8])
};

constexpr int ATT3_THREADS = 384;
// stage: Q0 16K | Q1 16K | K 32K | V 32K | Vx 2K (must follow V: 17th k-step of the PV MMA) | Kx 2K | Qx 2K
constexpr int ATT3_OFF_K = 32768, ATT3_OFF_V = 65536, ATT3_OFF_VX = 98304, ATT3_OFF_KX = 100352, ATT3_OFF_QX = 102400;
constexpr int ATT3_STAGE = 104448;
constexpr int ATT3_OFF_BAR = 2 * ATT3_STAGE;
// extras-QUERY partials: per stage, one record per softmax warp = 2 query rows x (64 output dims + row max + row sum)
constexpr int ATT3_PX_ROW = 66, ATT3_PX_REC = 2 * ATT3_PX_ROW;  // floats
constexpr int ATT3_OFF_PX = ATT3_OFF_BAR + 512;
constexpr int ATT3_SMEM = ATT3_OFF_PX + 2 * 8 * ATT3_PX_REC * 4;
constexpr uint32_t ATT3_Q_BYTES = 16384, ATT3_K_BYTES = 32768 + 2048 + 2048, ATT3_V_BYTES = 32768 + 2048;

// Round-2 pipeline notes (profiles/r02_attention_trace_before.txt -> _after.txt).  With ONE "stage free" barrier per
// operand stage (round 1) the two query tiles ran in lock-step: a tile that finished an item early still had to wait
// for the other tile (and the extras warp) before the NEXT-BUT-ONE item's operands could even be requested, and then
// for the ~3 000 clk the 102 KB TMA refill takes; both tiles then entered the exp pass together and shared the SM's
// single MUFU pipe (pass 2: ~3 900 clk each instead of ~2 100), while the pipe idled during everything else.  Now
//   * every operand has its own full / empty barrier pair (K+Kx+Qx, V+Vx, Q0, Q1): K is re-requested as soon as both
//     S MMAs have retired (early in the item), each tile's Q slot as soon as that tile's own output stores have read it;
//     the producer polls the four "empty" barriers and issues whichever refill is possible;
//   * each epilogue warp stores its own 32 rows with its own TMA store (no warpgroup barrier in the epilogue);
//   * optionally (AttnArgs::token) the tiles alternate in the exp pass through a pair of token barriers, so that one
//     tile's MMA waits / row max / epilogue always run under the other tile's exponentials;
//   * the extras QUERY rows (1-2 rows x 257 keys) used to be one warp's job (five serial mma.sync blocks per item,
//     ~5 000 clk on a scheduler it shares with two softmax warps) and every operand slot waited for it: with its key loop
//     switched off the launch was 13 % faster.  Now each of the eight softmax warps computes the partial attention of
//     the extras rows over ITS OWN 32 patch keys (one mma.sync block, in the slot where it used to compute the
//     extras-KEY scores with fp32 FMAs -- those now come from 8 mma.sync as well), and the extras warp only handles the
//     extras keys and merges the nine partials (max-rescaled sums).
__global__ void __launch_bounds__(ATT3_THREADS, 1) attention_tcgen05_kernel(const __grid_constant__ AttnArgs a) {
    extern __shared__ __align__(1024) uint8_t smem[];
    uint64_t* bars = reinterpret_cast<uint64_t*>(smem + ATT3_OFF_BAR);
    uint64_t* k_full = bars + 0;      // [2] per operand stage: K, Kx, Qx landed
    uint64_t* v_full = bars + 2;      // [2] V, Vx landed
    uint64_t* q_full = bars + 4;      // [2 tiles][2 stages] Q_t landed
    uint64_t* k_empty = bars + 8;     // [2] 11 arrivals: both tiles' S MMAs retired, extras warp done, and the eight
                                      //     softmax warps have read Kx (extras_scores)
    uint64_t* v_empty = bars + 10;    // [2] 11 arrivals: both tiles' PV MMAs retired, extras warp done, eight softmax warps
    uint64_t* q_empty = bars + 12;    // [2][2] 4 arrivals: the tile's four epilogue warps (their TMA stores have read it)
    uint64_t* s_full = bars + 16;     // [2] per query tile
    uint64_t* p_full = bars + 18;     // [2]
    uint64_t* o_full = bars + 20;     // [2]
    uint64_t* tmem_free = bars + 22;  // [2]
    uint64_t* p_half = bars + 24;     // [2] P of keys [0, 128) written
    uint64_t* tok = bars + 26;        // [2] tok[t]: tile t may enter its exp pass (4 arrivals from the other tile)
    uint64_t* px_full = bars + 28;    // [2] 8 arrivals: the softmax warps' extras-query partials of the item are written
    uint64_t* px_empty = bars + 30;   // [2] 1 arrival: the extras warp has merged them
    uint32_t* tmem_holder = reinterpret_cast<uint32_t*>(bars + 32);
    float* px_buf = reinterpret_cast<float*>(smem + ATT3_OFF_PX);

    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    const int D = a.H * 64;
    pdl_launch_dependents();   // PDL: the set-up below overlaps the predecessor's tail

    if (threadIdx.x == 0 && (smem_u32(smem) & 1023u)) __trap();  // swizzled tiles need 1024-byte alignment
    if (warp == 8 && lane == 0) {
        tma_prefetch_desc(&a.tmQKV);
        tma_prefetch_desc(&a.tmKV);
        tma_prefetch_desc(&a.tmX);
        tma_prefetch_desc(&a.tmOut32);
        for (int i = 0; i < 2; ++i) {
            mbar_init(&k_full[i], 1);
            mbar_init(&v_full[i], 1);
            mbar_init(&k_empty[i], 11);
            mbar_init(&v_empty[i], 11);
            mbar_init(&px_full[i], 8);
            mbar_init(&px_empty[i], 1);
            mbar_init(&s_full[i], 1);
            mbar_init(&p_full[i], 4);
            mbar_init(&o_full[i], 1);
            mbar_init(&tmem_free[i], 4);
            mbar_init(&p_half[i], 4);
            mbar_init(&tok[i], 4);
        }
        for (int i = 0; i < 4; ++i) {
            mbar_init(&q_full[i], 1);
            mbar_init(&q_empty[i], 4);
        }
        fence_mbar_init();
    }
    if (warp == 9) tmem_alloc<512>(tmem_holder);
    tc_fence_before();
    __syncthreads();
    tc_fence_after();
    const uint32_t tmem = *tmem_holder;
    pdl_wait();  // qkv is the predecessor's output -- and so is the live batch size of the early-exit compaction
    const int n_items = (a.b_dev ? ld_state(a.b_dev) : a.B) * a.H;
    const int my_items =
        (n_items > (int)blockIdx.x) ? (n_items - (int)blockIdx.x + (int)gridDim.x - 1) / (int)gridDim.x : 0;

    auto wait = [&](uint64_t* bar, uint32_t parity) { mbar_wait(bar, parity); };
    auto item_of = [&](int it) {
        return a.reverse ? n_items - 1 - (int)(blockIdx.x + it * gridDim.x) : (int)(blockIdx.x + it * gridDim.x);
    };

    if (warp == 8) {
        // ================================================================= TMA producer (+ L2 discard of consumed q|k|v)
        // a.discard: the q|k|v slice of a finished item is dead (the next block's qkv GEMM rewrites the whole buffer),
        // but its lines sit dirty in L2 and would be written back to HBM on eviction -- 101 MB per launch.  Once every
        // operand slot of item it-2 has been released its 3 x L lines of 128 bytes (one head's slice of one token:
        // exclusively this item's) are dropped with discard.global.L2.
        auto discard_item = [&](int it) {
            const int item = item_of(it);
            const int b = item / a.H, h = item % a.H;
            const __nv_bfloat16* base = a.qkv + (size_t)b * a.L * 3 * D + h * 64;
            for (int l = lane; l < a.L; l += 32) {
#pragma unroll
                for (int k = 0; k < 3; ++k)
                    asm volatile("discard.global.L2 [%0], 128;" ::"l"(base + (size_t)l * 3 * D + k * D) : "memory");
            }
        };
        for (int it = 0; it < my_items + (a.discard ? 2 : 0); ++it) {
            const int s = it & 1;
            if (lane == 0) {
                // the four operand slots of stage s are released at different times (K early, the Q slots when their
                // tile finishes, V last): poll, and refill whichever is free
                const uint32_t par = ((it >> 1) & 1) ^ 1;
                const bool load = it < my_items;
                const int item = load ? item_of(it) : 0;
                const int b = item / a.H, h = item % a.H;
                uint8_t* st = smem + s * ATT3_STAGE;
                uint32_t pend = 0xF, spins = 0;
                while (pend) {
                    if ((pend & 1) && mbar_test(&k_empty[s], par)) {
                        pend &= ~1u;
                        if (load) {
                            mbar_expect_tx(&k_full[s], ATT3_K_BYTES);
                            tma_load_3d(st + ATT3_OFF_K, &a.tmKV, &k_full[s], D + h * 64, a.extras, b);
                            tma_load_3d(st + ATT3_OFF_KX, &a.tmX, &k_full[s], D + h * 64, 0, b);
                            tma_load_3d(st + ATT3_OFF_QX, &a.tmX, &k_full[s], h * 64, 0, b);
                        }
                    }
                    if ((pend & 2) && mbar_test(&q_empty[0 * 2 + s], par)) {
                        pend &= ~2u;
                        if (load) {
                            mbar_expect_tx(&q_full[0 * 2 + s], ATT3_Q_BYTES);
                            tma_load_3d(st, &a.tmQKV, &q_full[0 * 2 + s], h * 64, a.extras, b);
                        }
                    }
                    if ((pend & 4) && mbar_test(&q_empty[1 * 2 + s], par)) {
                        pend &= ~4u;
                        if (load) {
                            mbar_expect_tx(&q_full[1 * 2 + s], ATT3_Q_BYTES);
                            tma_load_3d(st + 16384, &a.tmQKV, &q_full[1 * 2 + s], h * 64, a.extras + 128, b);
                        }
                    }
                    if ((pend & 8) && mbar_test(&v_empty[s], par)) {
                        pend &= ~8u;
                        if (load) {
                            mbar_expect_tx(&v_full[s], ATT3_V_BYTES);
                            tma_load_3d(st + ATT3_OFF_V, &a.tmKV, &v_full[s], 2 * D + h * 64, a.extras, b);
                            tma_load_3d(st + ATT3_OFF_VX, &a.tmX, &v_full[s], 2 * D + h * 64, 0, b);
                        }
                    }
                    if (++spins > DDB_SPIN_LIMIT) __trap();
                    if (pend) __nanosleep(64);  // nothing else to do: leave the issue slots to the softmax warps
                }
            }
            __syncwarp();
            if (a.discard && it >= 2) discard_item(it - 2);  // every consumer of item it-2 has released its slot
            __syncwarp();
        }
    } else if (warp == 9 || warp == 11) {
        // ================================================================= MMA issuers: one thread per query tile
        // (two independent in-order streams, so neither tile ever waits behind the other tile's barrier)
        if (lane == 0) {
            const int t = (warp == 9) ? 0 : 1;
            constexpr uint32_t idesc_s = umma_idesc_bf16(128, 256);
            constexpr uint32_t idesc_o = umma_idesc_bf16(128, 64, 0, 1);
            for (int it = 0; it < my_items; ++it) {
                const int s = it & 1;
                const uint32_t sph = (it >> 1) & 1;
                const uint8_t* st = smem + s * ATT3_STAGE;
                // S_t(it): needs K and Q_t of the item and the tile's TMEM columns (O_t(it-1) drained)
                wait(&k_full[s], sph);
                wait(&q_full[t * 2 + s], sph);
                wait(&tmem_free[t], (it & 1) ^ 1);
                // start tile 1 half a period late so the two softmax warpgroups do not fight over the MUFU pipe
                if (t == 1 && it == 0 && !(a.token & 1)) wait(&p_full[0], 0);
                tc_fence_after();
                const uint64_t dq = umma_desc_kmajor_sw128(smem_u32(st + t * 16384));
                const uint64_t dk = umma_desc_kmajor_sw128(smem_u32(st + ATT3_OFF_K));
#pragma unroll
                for (int k = 0; k < 4; ++k) umma_f16_ss(tmem + t * 256, dq + 2 * k, dk + 2 * k, idesc_s, k != 0);
                umma_commit(&s_full[t]);
                umma_commit(&k_empty[s]);  // this tile's S MMAs no longer read K once retired
                // O_t(it) = P_t [V; V_x]: 16 keys per k-step (P columns +8, V rows +16 = 2048 B); the first 8
                // k-steps start as soon as the first half of P is written
                long long* tr = (a.trace && blockIdx.x == 0) ? a.trace + (it * 2 + t) * 16 + 8 : nullptr;
                if (tr) tr[0] = clock64();
                wait(&v_full[s], sph);
                const uint64_t dv = umma_desc_mnmajor_sw128(smem_u32(st + ATT3_OFF_V));
                wait(&p_half[t], it & 1);
                tc_fence_after();
                if (tr) tr[1] = clock64();
#pragma unroll
                for (int k = 0; k < 8; ++k)
                    umma_f16_ts(tmem + t * 256 + 64, tmem + t * 256 + 8 * k, dv + (uint64_t)(k * (2048 >> 4)),
                                idesc_o, k != 0);
                if (tr) tr[2] = clock64();
                wait(&p_full[t], it & 1);
                tc_fence_after();
                if (tr) tr[3] = clock64();
#pragma unroll
                for (int k = 8; k < 17; ++k)
                    umma_f16_ts(tmem + t * 256 + 64, tmem + t * 256 + 64 + 8 * k, dv + (uint64_t)(k * (2048 >> 4)),
                                idesc_o, 1u);
                umma_commit(&o_full[t]);
                umma_commit(&v_empty[s]);  // this tile's PV MMAs no longer read V once retired
                if (tr) {
                    tr[4] = clock64();
                    wait(&o_full[t], it & 1);
                    tr[5] = clock64();
                }
            }
        }
    } else if (warp == 10) {
        // ================================================================= extras query rows: extras keys + merge
        const int g = lane >> 2, t = lane & 3;
        for (int it = 0; it < my_items; ++it) {
            const int item = item_of(it);
            const int b = item / a.H, h = item % a.H;
            const int s = it & 1;
            const uint8_t* st = smem + s * ATT3_STAGE;
            wait(&k_full[s], (it >> 1) & 1);
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
            wait(&v_full[s], (it >> 1) & 1);
            // extras keys: the first 8 tokens of the X tile, of which [0, extras) are valid
            att_mma_block<true>(smem_u32(st + ATT3_OFF_KX), smem_u32(st + ATT3_OFF_VX), 0, 1, a.extras, qf, rs,
                                a.scale_log2e, lane);
            __syncwarp();
            if (lane == 0) {  // this warp no longer reads the stage
                mbar_arrive(&k_empty[s]);
                mbar_arrive(&v_empty[s]);
            }
            float l0 = rs.l0;
            l0 += __shfl_xor_sync(0xffffffffu, l0, 1);
            l0 += __shfl_xor_sync(0xffffffffu, l0, 2);
            // merge with the eight partials over the patch keys (fixed order: deterministic)
            wait(&px_full[s], (it >> 1) & 1);
            if (g < a.extras) {
                const float* rec = px_buf + (size_t)s * 8 * ATT3_PX_REC + g * ATT3_PX_ROW;
                float M = rs.m0;
#pragma unroll
                for (int p = 0; p < 8; ++p) M = fmaxf(M, rec[p * ATT3_PX_REC + 64]);
                float sc = exp2f((rs.m0 - M) * a.scale_log2e);
                float Lsum = l0 * sc;
                float o[8][2];
#pragma unroll
                for (int i = 0; i < 8; ++i) o[i][0] = rs.o[i][0] * sc, o[i][1] = rs.o[i][1] * sc;
#pragma unroll
                for (int p = 0; p < 8; ++p) {
                    const float* r = rec + p * ATT3_PX_REC;
                    sc = exp2f((r[64] - M) * a.scale_log2e);
                    Lsum = fmaf(r[65], sc, Lsum);
#pragma unroll
                    for (int i = 0; i < 8; ++i) {
                        const float2 v = *reinterpret_cast<const float2*>(r + i * 8 + 2 * t);
                        o[i][0] = fmaf(v.x, sc, o[i][0]);
                        o[i][1] = fmaf(v.y, sc, o[i][1]);
                    }
                }
                const float inv0 = 1.f / Lsum;
                __nv_bfloat16* o0 = a.out + ((size_t)b * a.L + g) * D + h * 64 + 2 * t;
#pragma unroll
                for (int i = 0; i < 8; ++i)
                    *reinterpret_cast<uint32_t*>(o0 + i * 8) = pack_bf16(o[i][0] * inv0, o[i][1] * inv0);
            }
            __syncwarp();
            if (lane == 0) mbar_arrive(&px_empty[s]);
        }
    } else {
        // ================================================================= softmax + epilogue: one thread per query row
        const int t = warp >> 2;  // query tile == warpgroup
        const int quarter = warp & 3;
        const int r = quarter * 32 + lane;
        const uint32_t t_row = tmem + (uint32_t(quarter * 32) << 16) + t * 256;
        const float c = a.scale_log2e;
        const int q0 = a.extras + t * 128;

        // Per-item side work of a softmax warp on mma.sync (placements measured in profiles/r02_attention_experiments.txt):
        //  extras_scores(it)  scores of this warp's 32 query rows against the extras KEYS (tokens [0, extras)):
        //                     S = Q_rows Kx^T, 8 HMMA;
        //  extras_slice(it)   the extras QUERY rows' partial attention over this warp's own 32 patch keys (one softmax
        //                     block); the record (row max, row sum, 64 output dims per extras row) goes to shared memory
        //                     for the extras warp.  Last reader of K / V of the item in this warp.
        const int w8 = t * 4 + quarter;
        auto extras_scores = [&](int it, float& se0, float& se1) {
            const int s = it & 1;
            const uint32_t sph = (it >> 1) & 1;
            const uint8_t* st = smem + s * ATT3_STAGE;
            wait(&k_full[s], sph);
            wait(&q_full[t * 2 + s], sph);
#ifdef ATT_EXP_NOSCORES
            se0 = se1 = -INFINITY;
            return;
#endif
            const uint32_t sQ_u = smem_u32(st + t * 16384), sKx_u = smem_u32(st + ATT3_OFF_KX);
            uint32_t bk[4][2];  // B fragments of Kx: k-step ks -> (dims 16ks..+7, +8..+15) x keys 0..7
            ldmatrix_x4(sKx_u + att_swz(lane & 7, lane >> 3), bk[0][0], bk[0][1], bk[1][0], bk[1][1]);
            ldmatrix_x4(sKx_u + att_swz(lane & 7, (lane >> 3) + 4), bk[2][0], bk[2][1], bk[3][0], bk[3][1]);
            float cacc[2][4];
#pragma unroll
            for (int mt = 0; mt < 2; ++mt) {
                cacc[mt][0] = cacc[mt][1] = cacc[mt][2] = cacc[mt][3] = 0.f;
                const int row = quarter * 32 + mt * 16 + (lane & 7) + ((lane >> 3) & 1) * 8;
#pragma unroll
                for (int ks = 0; ks < 4; ++ks) {
                    uint32_t af[4];
                    ldmatrix_x4(sQ_u + att_swz(row, ks * 2 + (lane >> 4)), af[0], af[1], af[2], af[3]);
                    mma_bf16_16816(cacc[mt], af, bk[ks][0], bk[ks][1]);
                }
            }
            // row (lane) of this warp = m-tile lane/16, fragment row (lane & 15): held by lane 4 * (lane & 7)
            const int src = 4 * (lane & 7);
            float v[2][4];
#pragma unroll
            for (int mt = 0; mt < 2; ++mt)
#pragma unroll
                for (int k = 0; k < 4; ++k) v[mt][k] = __shfl_sync(0xffffffffu, cacc[mt][k], src);
            const int mt = lane >> 4, up = (lane >> 3) & 1;
            se0 = mt ? (up ? v[1][2] : v[1][0]) : (up ? v[0][2] : v[0][0]);
            se1 = mt ? (up ? v[1][3] : v[1][1]) : (up ? v[0][3] : v[0][1]);
            if (a.extras != 2) se1 = -INFINITY;
        };
        auto extras_slice = [&](int it) {
            const int s = it & 1;
            const uint32_t sph = (it >> 1) & 1;
            const uint8_t* st = smem + s * ATT3_STAGE;
            const int g = lane >> 2, tt = lane & 3;
            wait(&v_full[s], sph);
            float mq[2], lq[2], o[4][4];
#ifdef ATT_EXP_NOSLICE
            mq[0] = mq[1] = 0.f, lq[0] = lq[1] = 1.f;
            for (int i = 0; i < 4; ++i) o[i][0] = o[i][1] = o[i][2] = o[i][3] = 0.f;
#else
            att_extras_slice32_t(smem_u32(st + ATT3_OFF_K), smem_u32(st + ATT3_OFF_V), smem_u32(st + ATT3_OFF_QX), w8 * 32,
                                 a.scale_log2e, lane, mq, lq, o);
#endif
            wait(&px_empty[s], sph ^ 1);  // the extras warp has merged the record of item it-2
            if (tt == 0) {  // these lanes hold queries 0 and 1 = the extras rows: dims 16 dm + 8 h + g
                float* rec = px_buf + ((size_t)s * 8 + w8) * ATT3_PX_REC;
#pragma unroll
                for (int dm = 0; dm < 4; ++dm)
#pragma unroll
                    for (int h2 = 0; h2 < 2; ++h2) {
                        rec[dm * 16 + h2 * 8 + g] = o[dm][2 * h2];
                        rec[ATT3_PX_ROW + dm * 16 + h2 * 8 + g] = o[dm][2 * h2 + 1];
                    }
                if (g == 0) {
                    rec[64] = mq[0], rec[65] = lq[0];
                    rec[ATT3_PX_ROW + 64] = mq[1], rec[ATT3_PX_ROW + 65] = lq[1];
                }
            }
            __syncwarp();
            if (lane == 0) {  // this warp has read K / Kx / Qx / V of the item; its record is written
                mbar_arrive(&px_full[s]);
                mbar_arrive(&k_empty[s]);
                mbar_arrive(&v_empty[s]);
            }
        };

        float se0 = 0.f, se1 = 0.f;
        for (int it = 0; it < my_items; ++it) {
            const int item = item_of(it);
            const int b = item / a.H, h = item % a.H;
            const int s = it & 1;
            const uint32_t ph = it & 1;
            uint8_t* sQ = smem + s * ATT3_STAGE + t * 16384;  // this tile's Q; later its output staging buffer

            long long* tr = (a.trace && blockIdx.x == 0 && r == 0) ? a.trace + (it * 2 + t) * 16 : nullptr;
            if (tr) tr[0] = clock64();
            extras_scores(it, se0, se1);  // under the item's S MMA
            wait(&s_full[t], ph);
            tc_fence_after();
            if (tr) tr[1] = clock64();
            // this warp's store of the previous item was issued ~1000 clk ago: once it has read its 32 staging rows
            // (in the Q_t buffer of the other stage) that Q slot may be refilled
            if (a.direct_store) {
                // nothing is staged in the Q slot: it is free as soon as the S MMA has retired (s_full) and this warp has
                // read its rows for the extras-key scores (above)
                if (lane == 0) mbar_arrive(&q_empty[t * 2 + s]);
            } else if (lane == 0 && it > 0) {
                tma_store_wait_read<0>();
                mbar_arrive(&q_empty[t * 2 + (s ^ 1)]);
            }
            if (tr) tr[2] = clock64();
            uint32_t va[32], vb[32];
            // ---- pass 1: row max; the load of chunk j+1 is in flight while chunk j is reduced
            float m0 = se0, m1 = se1, m2 = -INFINITY, m3 = -INFINITY;
            tmem_ld_32x32b_x32(t_row, va);
            tmem_ld_wait();
#pragma unroll 1
            for (int jj = 0; jj < 8; jj += 2) {
#pragma unroll
                for (int u = 0; u < 2; ++u) {
                    uint32_t(&cur)[32] = u ? vb : va;
                    uint32_t(&nxt)[32] = u ? va : vb;
                    const int j = jj + u;
                    // (unconditional: a predicate here becomes a branch around the .sync.aligned load and cuts the loop
                    // body into basic blocks; after the last chunk the load wraps around to chunk 0, which pass 2 starts on)
                    tmem_ld_32x32b_x32(t_row + ((j + 1) & 7) * 32, nxt);
#pragma unroll
                    for (int e = 0; e < 32; e += 8) {
                        m0 = fmax3(m0, __uint_as_float(cur[e + 0]), __uint_as_float(cur[e + 1]));
                        m1 = fmax3(m1, __uint_as_float(cur[e + 2]), __uint_as_float(cur[e + 3]));
                        m2 = fmax3(m2, __uint_as_float(cur[e + 4]), __uint_as_float(cur[e + 5]));
                        m3 = fmax3(m3, __uint_as_float(cur[e + 6]), __uint_as_float(cur[e + 7]));
                    }
                    tmem_ld_wait();
                }
            }
            const float mc = fmaxf(fmaxf(m0, m1), fmaxf(m2, m3)) * c;
            if (tr) tr[3] = clock64();
            // the tiles take turns in the exp pass: tile t waits for its token (the other tile's previous exp pass has
            // ended), so that tile's MMA waits, row max and epilogue always run under this tile's exponentials
            if ((a.token & 1) && (t == 1 || it > 0)) wait(&tok[t], (t == 1 ? it : it - 1) & 1);
            if (tr) tr[14] = clock64();
            const float pe0 = ex2_approx(fmaf(se0, c, -mc));
            const float pe1 = (a.extras == 2) ? ex2_approx(fmaf(se1, c, -mc)) : 0.f;
            // ---- pass 2: P = exp2(s*c - m*c) -> bf16, written over the already-consumed S columns
            const f32x2 c2 = f2_splat(c), nmc2 = f2_splat(-mc);
            f32x2 sum2 = f2_pack(pe0, pe1), sum2b = f2_splat(0.f);
            // (va already holds chunk 0: pass 1's last prefetch wrapped around to it -- no TMEM round trip in front of
            // the first exponentials)
            // (fully unrolled, one basic block: the scheduler can run chunk j's exponentials on the MUFU pipe under the
            // conversions / row sums of chunk j-1 and the scaling of chunk j+1)
#pragma unroll 1
            for (int jj = 0; jj < 8; jj += 2) {
#pragma unroll
                for (int u = 0; u < 2; ++u) {
                    uint32_t(&cur)[32] = u ? vb : va;
                    uint32_t(&nxt)[32] = u ? va : vb;
                    const int j = jj + u;
                    // (unconditional, see pass 1: the branch around this load kept ptxas from interleaving the chunk's 32
                    // MUFU with its FADD2 / F2FP -- 438 vs 255 issue cycles per chunk; the extra load after the last
                    // chunk reads already-overwritten columns and is discarded)
                    tmem_ld_32x32b_x32(t_row + ((j + 1) & 7) * 32, nxt);
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
                    tmem_ld_wait();
                    // P chunk j < 4 -> columns [16j, 16j+16) (inside S chunk j/2), j >= 4 -> [128 + 16(j-4), ..)
                    // (inside S chunks 4, 5): always columns whose scores are already in registers
                    tmem_st_32x32b_x16(t_row + (j < 4 ? j * 16 : 64 + j * 16), pk);
                }
                if (jj == 2) {
                    // keys [0, 128) are done: the first 8 k-steps of O = P V run under the rest of pass 2
                    tmem_st_wait();
                    tc_fence_before();
                    __syncwarp();
                    if (lane == 0) mbar_arrive(&p_half[t]);
                }
            }
            if ((a.token & 1) && lane == 0) mbar_arrive(&tok[t ^ 1]);  // exp pass over: the other tile's turn
            {
                // 17th k-step: keys = tokens 0..15 of the sample, non-zero weight only for the extras tokens
                const uint32_t px[8] = {pack_bf16(pe0, pe1), 0u, 0u, 0u, 0u, 0u, 0u, 0u};
                tmem_st_32x32b_x8(t_row + 192, px);
            }
            tmem_st_wait();
            tc_fence_before();
            __syncwarp();
            if (lane == 0) mbar_arrive(&p_full[t]);
            if (tr) tr[4] = clock64();
            float sa, sb, sc, sd;
            f2_unpack(sum2, sa, sb);
            f2_unpack(sum2b, sc, sd);
            const float inv = 1.f / ((sa + sb) + (sc + sd));
            // while the tensor core finishes O: the extras rows' slice of THIS item (K and V are in shared memory: no wait
            // on a refill; the V slot of the next item is only requested when the OTHER tile's PV has retired, ~4 000 clk
            // before it lands -- working on item it+1 here stalled for ~2 000 clk per item)
            extras_slice(it);
            if (tr) tr[5] = clock64();

            // ---- epilogue: O row (fp32) out of TMEM, then release the tile's columns for S_t of the next item
            wait(&o_full[t], ph);
            tc_fence_after();
            if (tr) tr[6] = clock64();
            tmem_ld_32x32b_x32(t_row + 64, va);
            tmem_ld_32x32b_x32(t_row + 96, vb);
            tmem_ld_wait();
            tc_fence_before();
            __syncwarp();
            if (lane == 0) mbar_arrive(&tmem_free[t]);
            const f32x2 inv2 = f2_splat(inv);
            if (a.direct_store) {
                // the thread's output row is one full 128-byte line: eight 16-byte stores straight from registers -- no
                // staging tile, no generic->async proxy fence (MEMBAR.ALL.CTA, ~300 clk with the stores in flight), no TMA
                uint4* orow = reinterpret_cast<uint4*>(a.out + ((size_t)b * a.L + q0 + r) * D + h * 64);
#pragma unroll
                for (int j = 0; j < 8; ++j) {
                    const uint32_t* o = (j < 4) ? &va[j * 8] : &vb[(j - 4) * 8];
                    uint4 w;
                    w.x = f2_to_bf16x2(f2_mul(f2_pack_u(o[0], o[1]), inv2));
                    w.y = f2_to_bf16x2(f2_mul(f2_pack_u(o[2], o[3]), inv2));
                    w.z = f2_to_bf16x2(f2_mul(f2_pack_u(o[4], o[5]), inv2));
                    w.w = f2_to_bf16x2(f2_mul(f2_pack_u(o[6], o[7]), inv2));
                    orow[j] = w;
                }
            } else {
                // row r of the Q tile is only ever touched by this thread and by the (retired) S MMA: reuse it as staging
                uint8_t* srow = sQ + r * 128;
#pragma unroll
                for (int j = 0; j < 8; ++j) {
                    const uint32_t* o = (j < 4) ? &va[j * 8] : &vb[(j - 4) * 8];
                    uint4 w;
                    w.x = f2_to_bf16x2(f2_mul(f2_pack_u(o[0], o[1]), inv2));
                    w.y = f2_to_bf16x2(f2_mul(f2_pack_u(o[2], o[3]), inv2));
                    w.z = f2_to_bf16x2(f2_mul(f2_pack_u(o[4], o[5]), inv2));
                    w.w = f2_to_bf16x2(f2_mul(f2_pack_u(o[6], o[7]), inv2));
                    *reinterpret_cast<uint4*>(srow + ((j ^ (r & 7)) << 4)) = w;
                }
                // each warp stores its own 32 rows (4 KB, a whole number of 1 KB swizzle atoms): no warpgroup barrier
                fence_proxy_async_smem();
                __syncwarp();
                if (lane == 0) {
                    tma_store_3d(&a.tmOut32, sQ + quarter * 4096, h * 64, q0 + quarter * 32, b);
                    tma_store_commit();
                }
            }
            if (tr) tr[7] = clock64();
        }
        if (lane == 0 && my_items > 0 && !a.direct_store) {
            tma_store_wait_read<0>();
            mbar_arrive(&q_empty[t * 2 + ((my_items - 1) & 1)]);
            tma_store_wait_all<0>();
        }
    }
    __syncwarp();
    tc_fence_before();
    __syncthreads();
    if (warp == 9) {
        tc_fence_after();
        tmem_dealloc<512>(tmem);
    }
}

}  // namespace ddb
