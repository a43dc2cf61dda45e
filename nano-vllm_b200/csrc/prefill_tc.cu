// Causal varlen prefill attention on Blackwell's 5th-generation tensor cores (tcgen05 + TMEM + TMA).
//
// Replaces the prefill branch of the reference's Attention.forward (nanovllm/layers/attention.py:64-70,
// flash_attn_varlen_func with an optional block_table): packed or paged K/V, bottom-right aligned
// causal mask (query i sees keys j <= i + len_k - len_q).
//
// One CTA = 128 queries of one head (UMMA M = 128, cta_group::1), key blocks of 64, two CTAs per SM:
//   * Q, K and V tiles are staged by TMA (cp.async.bulk.tensor, 128-byte swizzle) into shared memory;
//     a paged K/V tile is fetched page by page through the block table with the same tensor map;
//   * S = Q K^T : tcgen05.mma  M128 N64 K16 x 8, both operands K-major, fp32 accumulator in TMEM
//     (two S buffers, so the next block's QK^T runs under this block's softmax);
//   * softmax: one thread per query row reads its row with tcgen05.ld (32x32b), so row max / row sum need no
//     shuffles; P is rounded to bf16 (as the reference's kernel does) and written to shared memory in the
//     K-major 128B-swizzled layout the MMA expects;
//   * O += P V : tcgen05.mma M128 N128 K16 x 4, V consumed as an MN-major operand straight from its
//     [key][d] tile, accumulator stays in TMEM across key blocks; the running reference max only moves
//     (and O is only rescaled through tcgen05.ld/st) when a row's max grows by more than 2^8;
//   * completion of MMAs is tracked with tcgen05.commit -> mbarrier; one elected thread issues TMA and MMA.
#include "tc_common.cuh"

namespace {

using namespace b200tc;

constexpr int BM = 128;
constexpr int BN = 64;
constexpr int D = B200_HEAD_DIM;
constexpr int TC_THREADS = 128;
constexpr uint32_t Q_BYTES = BM * D * 2;          // 32 KB: two [128][64] halves
constexpr uint32_t KV_BYTES = BN * D * 2;         // 16 KB: two [64][64] halves
constexpr uint32_t Q_HALF = BM * 128;             // 16 KB
constexpr uint32_t KV_HALF = BN * 128;            // 8 KB
constexpr uint32_t P_BYTES = BM * BN * 2;         // 16 KB: [128][64]
constexpr uint32_t OFF_Q = 0;
constexpr uint32_t OFF_K = OFF_Q + Q_BYTES;       // 2 buffers
constexpr uint32_t OFF_V = OFF_K + 2 * KV_BYTES;  // 2 buffers
constexpr uint32_t OFF_P = OFF_V + 2 * KV_BYTES;
constexpr uint32_t OFF_BAR = OFF_P + P_BYTES;     // 8 mbarriers + tmem slot
constexpr uint32_t TC_SMEM = OFF_BAR + 128;       // 112.1 KB: two CTAs per SM
constexpr uint32_t TMEM_COLS = 256;               // S0 [0,64) S1 [64,128) O [128,256)
constexpr float RESCALE_THRESHOLD = 8.0f;         // log2 units

struct TcParams {
    const int32_t* cu_q;
    const int32_t* cu_k;
    const int32_t* block_tables;   // null => packed k/v
    int bt_stride;
    __nv_bfloat16* out;
    int64_t out_stride;
    int hq, hkv, block_shift;
    int box_rows;                  // rows per TMA box of a paged tile: min(block_size, 64)
    int64_t layer_row0;            // first row of this layer in the [rows, 128] view of the cache
    float scale_log2;
};

constexpr uint32_t IDESC_QK = make_idesc(BM, BN, false);
constexpr uint32_t IDESC_PV = make_idesc(BM, D, true);

__global__ void __launch_bounds__(TC_THREADS, 2)
prefill_tc_kernel(const __grid_constant__ CUtensorMap tm_q, const __grid_constant__ CUtensorMap tm_k,
                  const __grid_constant__ CUtensorMap tm_v, const TcParams p) {
    B200_PDL_SYNC();
    extern __shared__ __align__(1024) uint8_t smem_raw[];   // 128-byte swizzle atoms need 1024-byte alignment
    const uint32_t base = smem_u32(smem_raw);
    if (base & 1023u) __trap();
    uint8_t* base_ptr = smem_raw;
    const uint32_t sQ = base + OFF_Q, sK = base + OFF_K, sV = base + OFF_V, sP = base + OFF_P;
    const uint32_t bars = base + OFF_BAR;
    const uint32_t bar_q = bars, bar_k0 = bars + 8, bar_v0 = bars + 24, bar_s0 = bars + 40, bar_o = bars + 56;
    volatile uint32_t* tmem_slot = reinterpret_cast<volatile uint32_t*>(base_ptr + OFF_BAR + 64);

    const int seq = blockIdx.z;
    const int head = blockIdx.y;
    const int mb = gridDim.x - 1 - blockIdx.x;      // longest tiles first
    const int q0 = p.cu_q[seq];
    const int len_q = p.cu_q[seq + 1] - q0;
    const int k0 = p.cu_k[seq];
    const int len_k = p.cu_k[seq + 1] - k0;
    if (mb * BM >= len_q) return;
    const int off = len_k - len_q;
    const int kvh = head / (p.hq / p.hkv);
    int j_end = mb * BM + BM + off;
    j_end = j_end > len_k ? len_k : j_end;
    const int nblocks = (j_end + BN - 1) / BN;

    const int tid = threadIdx.x;
    const int warp = tid >> 5;
    const bool paged = p.block_tables != nullptr;

    if (tid == 0) {
        for (int i = 0; i < 8; ++i) mbar_init(bars + i * 8, 1);
        mbar_fence_init();
    }
    if (warp == 0) {
        asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(bars + 64), "r"(TMEM_COLS) : "memory");
        asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;" ::: "memory");
    }
    tc_fence_before();
    __syncthreads();
    tc_fence_after();
    const uint32_t tmem = *tmem_slot;
    const uint32_t tmem_o = tmem + 128;

    auto load_kv = [&](const CUtensorMap* map, uint32_t dst, uint32_t bar, int jb) {   // one 64-key tile, both halves
        mbar_expect_tx(bar, KV_BYTES);
        if (!paged) {
#pragma unroll
            for (int h = 0; h < 2; ++h) tma_load_2d(dst + h * KV_HALF, map, bar, kvh * D + h * 64, k0 + jb * BN);
        } else {
            const int R = p.box_rows;
            for (int s = 0; s < BN / R; ++s) {
                const int key = jb * BN + s * R;
                int page = key < len_k ? p.block_tables[(int64_t)seq * p.bt_stride + (key >> p.block_shift)] : 0;
                page = page < 0 ? 0 : page;
                const int64_t row = p.layer_row0 + ((((int64_t)page * p.hkv + kvh) << p.block_shift) + (key & ((1 << p.block_shift) - 1)));
#pragma unroll
                for (int h = 0; h < 2; ++h) tma_load_2d(dst + h * KV_HALF + s * R * 128, map, bar, h * 64, (int)row);
            }
        }
    };
    auto issue_qk = [&](int jb) {            // S[jb & 1] = Q K(jb)^T
        const uint32_t kb = sK + (jb & 1) * KV_BYTES;
#pragma unroll
        for (int ks = 0; ks < D / 16; ++ks) {
            const uint32_t hoff_q = (ks >> 2) * Q_HALF + (ks & 3) * 32;
            const uint32_t hoff_k = (ks >> 2) * KV_HALF + (ks & 3) * 32;
            tc_mma(tmem + (jb & 1) * BN, make_desc(sQ + hoff_q, 16, 1024), make_desc(kb + hoff_k, 16, 1024), IDESC_QK, ks > 0);
        }
        tc_commit(bar_s0 + (jb & 1) * 8);
    };

    if (tid == 0) {
        mbar_expect_tx(bar_q, Q_BYTES);
#pragma unroll
        for (int h = 0; h < 2; ++h) tma_load_2d(sQ + h * Q_HALF, &tm_q, bar_q, head * D + h * 64, q0 + mb * BM);
        load_kv(&tm_k, sK, bar_k0, 0);
        load_kv(&tm_v, sV, bar_v0, 0);
        if (nblocks > 1) {
            load_kv(&tm_k, sK + KV_BYTES, bar_k0 + 8, 1);
            load_kv(&tm_v, sV + KV_BYTES, bar_v0 + 8, 1);
        }
        mbar_wait(bar_q, 0);
        mbar_wait(bar_k0, 0);
        tc_fence_after();
        issue_qk(0);
    }

    const int row = mb * BM + tid;                        // in-sequence query index of this thread's row
    const uint32_t lane_base = (uint32_t)(warp * 32) << 16;
    float m_ref = -INFINITY, l_sum = 0.f;

    for (int j = 0; j < nblocks; ++j) {
        const int b = j & 1;
        if (tid == 0 && j + 1 < nblocks) {                // next block's QK^T runs under this block's softmax
            mbar_wait(bar_k0 + ((j + 1) & 1) * 8, ((j + 1) >> 1) & 1);
            tc_fence_after();
            issue_qk(j + 1);
        }
        mbar_wait(bar_s0 + b * 8, (j >> 1) & 1);          // S(j) is in TMEM, K buffer b is free again
        tc_fence_after();
        if (tid == 0 && j + 2 < nblocks) load_kv(&tm_k, sK + b * KV_BYTES, bar_k0 + b * 8, j + 2);

        // ---- this thread's row of S -------------------------------------------------------------
        float s[BN];
        {
            float t0[32], t1[32];
            tmem_ld32(tmem + lane_base + b * BN, t0);
            tmem_ld32(tmem + lane_base + b * BN + 32, t1);
#pragma unroll
            for (int c = 0; c < 32; ++c) { s[c] = t0[c] * p.scale_log2; s[32 + c] = t1[c] * p.scale_log2; }
        }
        const bool need_mask = (j * BN + BN > len_k) || (j * BN + BN - 1 > mb * BM + warp * 32 + off);
        if (need_mask) {
#pragma unroll
            for (int c = 0; c < BN; ++c) {
                const int jj = j * BN + c;
                if (jj >= len_k || jj > row + off) s[c] = -INFINITY;
            }
        }
        float m_new = s[0];
#pragma unroll
        for (int c = 1; c < BN; ++c) m_new = fmaxf(m_new, s[c]);

        // ---- P V of the previous block must be finished before P is overwritten or O is rescaled ---------
        if (j > 0) {
            mbar_wait(bar_o, (j - 1) & 1);
            tc_fence_after();
            if (tid == 0 && j + 1 < nblocks) load_kv(&tm_v, sV + ((j + 1) & 1) * KV_BYTES, bar_v0 + ((j + 1) & 1) * 8, j + 1);
        }
        if (j == 0) {
            m_ref = (m_new == -INFINITY) ? 0.f : m_new;
        } else if (__any_sync(0xffffffffu, m_new > m_ref + RESCALE_THRESHOLD)) {
            const float m_upd = fmaxf(m_ref, m_new);
            const float f = fast_exp2(m_ref - m_upd);
            m_ref = m_upd;
            l_sum *= f;
#pragma unroll
            for (int q4 = 0; q4 < 4; ++q4) {
                float o[32];
                tmem_ld32(tmem_o + lane_base + q4 * 32, o);
#pragma unroll
                for (int c = 0; c < 32; ++c) o[c] *= f;
                tmem_st32(tmem_o + lane_base + q4 * 32, o);
            }
        }

        // ---- P = exp2(S - m_ref) -> bf16 -> shared memory (K-major, 128-byte swizzle) -----------------------
        uint8_t* prow = base_ptr + OFF_P + tid * 128;
#pragma unroll
        for (int c8 = 0; c8 < 8; ++c8) {
            float pv[8];
#pragma unroll
            for (int e = 0; e < 8; ++e) {
                pv[e] = fast_exp2(s[c8 * 8 + e] - m_ref);
                l_sum += pv[e];
            }
            *reinterpret_cast<uint4*>(prow + ((c8 ^ (tid & 7)) << 4)) = pack8(pv);
        }
        if (j * BN + BN > len_k) {
            // Tail block: V rows at keys >= len_k are not this sequence's (stale page rows, the next sequence of a packed
            // batch, anything -- possibly NaN/Inf).  Their P is exactly 0, but 0 * NaN would poison the accumulator, so
            // zero those rows of the staged tile.  A whole 128-byte row is zeroed, which is swizzle-agnostic.
            mbar_wait(bar_v0 + b * 8, (j >> 1) & 1);
            const int first = len_k - j * BN;                               // 1..63
            uint8_t* vtile = base_ptr + OFF_V + b * KV_BYTES;
            for (int i = tid; i < (BN - first) * 16; i += TC_THREADS) {     // 16 chunks of 16 B per row: 2 halves x 8
                const int r = first + (i >> 4), c = i & 15;
                *reinterpret_cast<uint4*>(vtile + (c >> 3) * KV_HALF + r * 128 + ((c & 7) << 4)) = make_uint4(0, 0, 0, 0);
            }
        }
        asm volatile("fence.proxy.async.shared::cta;" ::: "memory");     // generic-proxy writes -> visible to the MMA
        tc_fence_before();
        __syncthreads();

        // ---- O (+)= P V(j) ---------------------------------------------------------------------------
        if (tid == 0) {
            mbar_wait(bar_v0 + b * 8, (j >> 1) & 1);
            tc_fence_after();
            const uint32_t vb = sV + b * KV_BYTES;
#pragma unroll
            for (int ks = 0; ks < BN / 16; ++ks)
                tc_mma(tmem_o, make_desc(sP + ks * 32, 16, 1024), make_desc(vb + ks * 2048, KV_HALF, 1024), IDESC_PV,
                       (j > 0 || ks > 0) ? 1u : 0u);
            tc_commit(bar_o);
        }
    }

    // ---- epilogue: O / l -> bf16 -> global ------------------------------------------------------------
    mbar_wait(bar_o, (nblocks - 1) & 1);
    tc_fence_after();
    const float inv = l_sum > 0.f ? 1.f / l_sum : 0.f;
    const bool valid = row < len_q;
    __nv_bfloat16* dst = p.out + (int64_t)(q0 + (valid ? row : 0)) * p.out_stride + head * D;
#pragma unroll
    for (int q4 = 0; q4 < 4; ++q4) {
        float o[32];
        tmem_ld32(tmem_o + lane_base + q4 * 32, o);
        if (valid) {
#pragma unroll
            for (int c8 = 0; c8 < 4; ++c8) {
                float r[8];
#pragma unroll
                for (int e = 0; e < 8; ++e) r[e] = o[c8 * 8 + e] * inv;
                *reinterpret_cast<uint4*>(dst + q4 * 32 + c8 * 8) = pack8(r);
            }
        }
    }
    tc_fence_before();
    __syncthreads();
    if (warp == 0) asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" ::"r"(tmem), "r"(TMEM_COLS) : "memory");
}

}  // namespace

// Called from b200_paged_prefill (prefill_attn.cu).  Returns B200_EUNSUPPORTED if TMA descriptors cannot be built.
int b200_prefill_tc_launch(b200_ctx* ctx, int layer, const void* q, int64_t q_stride0, const void* k, int64_t k_stride0,
                           const void* v, int64_t v_stride0, const int32_t* cu_q, const int32_t* cu_k,
                           const int32_t* block_tables, int bt_stride, void* out, int64_t out_stride0, int total_q,
                           int num_seqs, int max_seqlen_q, int num_q_heads, int num_kv_heads, float scale,
                           cudaStream_t stream) {
    const int total_k = total_q;   // packed mode: K/V rows are the step's own tokens (attention.py:67-70 without block_table)
    if ((out_stride0 % 8) || ((uintptr_t)out & 15)) return B200_EINVAL;
    CUtensorMap tq, tk, tv;
    if (!make_map(&tq, q, (uint64_t)num_q_heads * D, (uint64_t)total_q, (uint64_t)q_stride0, BM)) return B200_EUNSUPPORTED;
    TcParams prm;
    prm.cu_q = cu_q;
    prm.cu_k = cu_k;
    prm.block_tables = block_tables;
    prm.bt_stride = bt_stride;
    prm.out = static_cast<__nv_bfloat16*>(out);
    prm.out_stride = out_stride0;
    prm.hq = num_q_heads;
    prm.hkv = num_kv_heads;
    prm.scale_log2 = scale * 1.4426950408889634f;
    prm.block_shift = 0;
    prm.box_rows = BN;
    prm.layer_row0 = 0;
    if (block_tables) {
        const uint32_t R = ctx->block_size < BN ? ctx->block_size : BN;
        const uint64_t rows = (uint64_t)ctx->layers * ctx->num_blocks * ctx->num_kv_heads * ctx->block_size;
        if (rows >= (1ull << 31)) return B200_EUNSUPPORTED;
        if (!make_map(&tk, ctx->k_base, D, rows, D, R) || !make_map(&tv, ctx->v_base, D, rows, D, R)) return B200_EUNSUPPORTED;
        prm.block_shift = ctx->block_shift;
        prm.box_rows = (int)R;
        prm.layer_row0 = (int64_t)layer * ctx->num_blocks * ctx->num_kv_heads * ctx->block_size;
    } else {
        if (!make_map(&tk, k, (uint64_t)num_kv_heads * D, (uint64_t)total_k, (uint64_t)k_stride0, BN) ||
            !make_map(&tv, v, (uint64_t)num_kv_heads * D, (uint64_t)total_k, (uint64_t)v_stride0, BN))
            return B200_EUNSUPPORTED;
    }
    static B200SmemOptIn optin;
    B200_CUDA_CHECK(ctx, optin.ensure(prefill_tc_kernel, TC_SMEM));
    dim3 grid((max_seqlen_q + BM - 1) / BM, num_q_heads, num_seqs);
    if (grid.z > 65535 || grid.y > 65535) return B200_EUNSUPPORTED;
    B200_LAUNCH((prefill_tc_kernel), grid, TC_THREADS, TC_SMEM, stream, tq, tk, tv, prm);
    return b200_launch_status(ctx);
}
