// Warp-specialised causal varlen prefill attention (tcgen05 + TMEM + TMA), second generation of prefill_tc.cu.
//
// Same operator (reference nanovllm/layers/attention.py:64-70: flash_attn_varlen_func, packed or paged K/V, bottom-right
// aligned causal mask), different machine mapping.  prefill_tc.cu's CTA does MMA issue and softmax with the same 128
// threads, so its tensor pipe idles while the rows are exponentiated (ncu: 21.6 % tensor-active, top stalls on the S / PV
// mbarriers).  Here one CTA owns 128 queries of the TWO q-heads of a GQA group that share a kv head (every K / V tile is
// staged once for both), and the roles are split:
//
//   warps 0-3  ("WG0")  softmax of head 0: thread r owns query row r; reads its row of S from TMEM (tcgen05.ld 32x32b),
//                       scales / masks / exponentiates it, writes P (bf16, K-major 128-byte swizzle) to shared memory,
//                       keeps the running reference max and row sum, rescales O in TMEM when the max moves by > 2^8
//   warps 4-7  ("WG1")  the same for head 1
//   warp 8              TMA producer: Q (both heads) once, then a 3-deep ring of K and V tiles (64 keys), paged tiles
//                       page by page through the block table
//   warp 9              tcgen05.mma issuer: S_t = Q_t K^T (M128 N64 K16 x8) into a double-buffered S per head, O_t += P_t V
//                       (M128 N128 K16 x4, V as an MN-major operand) -- QK^T of block j+1 is issued before the PV of block j,
//                       so the tensor pipe works on the next scores of both heads while both warpgroups exponentiate
//
// TMEM (512 columns): S[head][buf] 4 x 64, O[head] 2 x 128.  Shared memory: Q 2 x 32 KB, K 3 x 16 KB, V 3 x 16 KB,
// P 2 x 16 KB = 192 KB, one CTA per SM.  Arithmetic and rounding points are those of prefill_tc.cu (P rounded to bf16 as
// the reference's kernel does, fp32 accumulation, one rounding of the output).
#include <cstdlib>

#include "tc_common.cuh"

namespace {

using namespace b200tc;

constexpr int BM = 128, BN = 64, D = B200_HEAD_DIM;
constexpr int WS_THREADS = 320;                    // 2 softmax warpgroups + producer warp + MMA warp
constexpr int KV_STAGES = 3;
constexpr uint32_t Q_BYTES = BM * D * 2;           // 32 KB per head: two [128][64] halves
constexpr uint32_t Q_HALF = BM * 128;
constexpr uint32_t KV_BYTES = BN * D * 2;          // 16 KB: two [64][64] halves
constexpr uint32_t KV_HALF = BN * 128;
constexpr uint32_t P_BYTES = BM * BN * 2;          // 16 KB
constexpr uint32_t OFF_Q = 0;                                        // 2 heads
constexpr uint32_t OFF_K = OFF_Q + 2 * Q_BYTES;
constexpr uint32_t OFF_V = OFF_K + KV_STAGES * KV_BYTES;
constexpr uint32_t OFF_P = OFF_V + KV_STAGES * KV_BYTES;             // 2 heads
constexpr uint32_t OFF_BAR = OFF_P + 2 * P_BYTES;
constexpr uint32_t WS_SMEM = OFF_BAR + 256;
constexpr uint32_t TMEM_COLS = 512;                // S[t][b] at (t * 2 + b) * 64, O[t] at 256 + t * 128
constexpr float RESCALE_THRESHOLD = 8.0f;          // log2 units
constexpr uint32_t IDESC_QK = make_idesc(BM, BN, false);
constexpr uint32_t IDESC_PV = make_idesc(BM, D, true);

// barrier slots (8 bytes each) from OFF_BAR
constexpr uint32_t B_Q = 0;                        // Q tiles landed
constexpr uint32_t B_KFULL = 1, B_KEMPTY = 4, B_VFULL = 7, B_VEMPTY = 10;   // 3 each
constexpr uint32_t B_SFULL = 13;                   // [t][b]: 13 + t * 2 + b
constexpr uint32_t B_PREADY = 17;                  // [t]: P_t written (and S_t read, O_t rescaled): 128 arrivals
constexpr uint32_t B_ODONE = 19;                   // [t]: the PV MMA of the block has completed
constexpr uint32_t B_COUNT = 21;
constexpr uint32_t TMEM_SLOT = B_COUNT * 8;        // byte offset of the tcgen05.alloc result

struct WsParams {
    const int32_t* cu_q;
    const int32_t* cu_k;
    const int32_t* block_tables;   // null => packed k/v
    int bt_stride;
    __nv_bfloat16* out;
    int64_t out_stride;
    int hq, hkv, block_shift;
    int box_rows;                  // rows per TMA box of a paged tile: min(block_size, 64)
    int64_t layer_row0;
    float scale_log2;
};

// 2^x for x <= ~8 on the FMA/ALU pipes instead of the 16-per-clock MUFU unit (the softmax of this kernel is MUFU-bound: one
// exp2 per score).  Cody-Waite split x = n + f, f in [0, 1), and a degree-3 minimax polynomial for 2^f (relative error
// < 1e-4, far inside the bf16 rounding P gets anyway); the exponent is added into the float's exponent field.
__device__ __forceinline__ float exp2_poly(float x) {
    const float xc = fmaxf(x, -120.0f);
    const float n = floorf(xc);
    const float f = xc - n;
    float pf = fmaf(f, 0.07790716f, 0.22623319f);      // 2^f ~ c0 + f (c1 + f (c2 + f c3)): relative error < 7.8e-5 on [0, 1)
    pf = fmaf(f, pf, 0.6957771f);
    pf = fmaf(f, pf, 0.99992783f);
    const float y = __int_as_float(__float_as_int(pf) + ((int)n << 23));
    return x < -120.0f ? 0.f : y;                      // masked scores (-inf) must give exactly 0
}

__device__ __forceinline__ void mbar_arrive_ws(uint32_t bar) {
    asm volatile("mbarrier.arrive.shared::cta.b64 _, [%0];" ::"r"(bar) : "memory");
}

template <bool POLY>
__global__ void __launch_bounds__(WS_THREADS, 1)
prefill_ws_kernel(const __grid_constant__ CUtensorMap tm_q, const __grid_constant__ CUtensorMap tm_k,
                  const __grid_constant__ CUtensorMap tm_v, const WsParams p) {
    B200_PDL_SYNC();
    extern __shared__ __align__(1024) uint8_t smem_raw[];
    const uint32_t base = smem_u32(smem_raw);
    if (base & 1023u) __trap();
    const uint32_t bars = base + OFF_BAR;
    auto bar = [&](uint32_t i) { return bars + i * 8; };
    volatile uint32_t* tmem_slot = reinterpret_cast<volatile uint32_t*>(smem_raw + OFF_BAR + TMEM_SLOT);

    const int seq = blockIdx.z;
    const int g_heads = p.hq / p.hkv;                      // q heads per kv head
    const int pairs_per_kv = (g_heads + 1) / 2;
    const int kvh = blockIdx.y / pairs_per_kv;
    const int head0 = kvh * g_heads + (blockIdx.y % pairs_per_kv) * 2;
    const bool two = (head0 + 1) < (kvh + 1) * g_heads;    // an odd group leaves the second tile of its last pair idle
    const int mb = gridDim.x - 1 - blockIdx.x;             // longest tiles first
    const int q0 = p.cu_q[seq];
    const int len_q = p.cu_q[seq + 1] - q0;
    const int k0 = p.cu_k[seq];
    const int len_k = p.cu_k[seq + 1] - k0;
    if (mb * BM >= len_q) return;
    const int off = len_k - len_q;
    int j_end = mb * BM + BM + off;
    j_end = j_end > len_k ? len_k : j_end;
    const int nblocks = (j_end + BN - 1) / BN;

    const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
    const bool paged = p.block_tables != nullptr;

    if (tid == 0) {
        for (uint32_t i = 0; i < B_COUNT; ++i) mbar_init(bar(i), (i == B_PREADY || i == B_PREADY + 1) ? 128 : 1);
        mbar_fence_init();
    }
    if (warp == 0) {
        asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(bars + TMEM_SLOT), "r"(TMEM_COLS) : "memory");
        asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;" ::: "memory");
    }
    tc_fence_before();
    __syncthreads();
    tc_fence_after();
    const uint32_t tmem = *tmem_slot;

    if (warp == 8) {
        // ---- TMA producer -------------------------------------------------------------------------------------------
        if (lane == 0) {
            mbar_expect_tx(bar(B_Q), two ? 2 * Q_BYTES : Q_BYTES);
            for (int t = 0; t < (two ? 2 : 1); ++t)
#pragma unroll
                for (int h = 0; h < 2; ++h)
                    tma_load_2d(base + OFF_Q + t * Q_BYTES + h * Q_HALF, &tm_q, bar(B_Q), (head0 + t) * D + h * 64, q0 + mb * BM);
            auto load_kv = [&](const CUtensorMap* map, uint32_t dst, uint32_t full, int jb) {   // one 64-key tile, both halves
                mbar_expect_tx(full, KV_BYTES);
                if (!paged) {
#pragma unroll
                    for (int h = 0; h < 2; ++h) tma_load_2d(dst + h * KV_HALF, map, full, kvh * D + h * 64, k0 + jb * BN);
                } else {
                    const int R = p.box_rows;
                    for (int s = 0; s < BN / R; ++s) {
                        const int key = jb * BN + s * R;
                        int page = key < len_k ? p.block_tables[(int64_t)seq * p.bt_stride + (key >> p.block_shift)] : 0;
                        page = page < 0 ? 0 : page;
                        const int64_t row = p.layer_row0 + ((((int64_t)page * p.hkv + kvh) << p.block_shift) + (key & ((1 << p.block_shift) - 1)));
#pragma unroll
                        for (int h = 0; h < 2; ++h) tma_load_2d(dst + h * KV_HALF + s * R * 128, map, full, h * 64, (int)row);
                    }
                }
            };
            for (int j = 0; j < nblocks; ++j) {
                const int s = j % KV_STAGES;
                if (j >= KV_STAGES) {
                    const uint32_t par = ((j / KV_STAGES) - 1) & 1;
                    mbar_wait(bar(B_KEMPTY + s), par);
                    load_kv(&tm_k, base + OFF_K + s * KV_BYTES, bar(B_KFULL + s), j);
                    mbar_wait(bar(B_VEMPTY + s), par);
                    load_kv(&tm_v, base + OFF_V + s * KV_BYTES, bar(B_VFULL + s), j);
                } else {
                    load_kv(&tm_k, base + OFF_K + s * KV_BYTES, bar(B_KFULL + s), j);
                    load_kv(&tm_v, base + OFF_V + s * KV_BYTES, bar(B_VFULL + s), j);
                }
            }
        }
    } else if (warp == 9) {
        // ---- MMA issuer ---------------------------------------------------------------------------------------------------
        if (lane == 0) {
            const int nt = two ? 2 : 1;
            auto issue_qk = [&](int j) {                   // S[t][j & 1] = Q_t K(j)^T for both heads, then K slot free
                const int s = j % KV_STAGES;
                mbar_wait(bar(B_KFULL + s), (j / KV_STAGES) & 1);
                tc_fence_after();
                const uint32_t kb = base + OFF_K + s * KV_BYTES;
                for (int t = 0; t < nt; ++t) {
                    const uint32_t qb = base + OFF_Q + t * Q_BYTES;
#pragma unroll
                    for (int ks = 0; ks < D / 16; ++ks) {
                        const uint32_t hq_off = (ks >> 2) * Q_HALF + (ks & 3) * 32;
                        const uint32_t hk_off = (ks >> 2) * KV_HALF + (ks & 3) * 32;
                        tc_mma(tmem + (t * 2 + (j & 1)) * BN, make_desc(qb + hq_off, 16, 1024), make_desc(kb + hk_off, 16, 1024), IDESC_QK, ks > 0);
                    }
                    tc_commit(bar(B_SFULL + t * 2 + (j & 1)));
                }
                tc_commit(bar(B_KEMPTY + s));
            };
            mbar_wait(bar(B_Q), 0);
            tc_fence_after();
            issue_qk(0);
            for (int j = 0; j < nblocks; ++j) {
                if (j + 1 < nblocks) issue_qk(j + 1);      // the next scores run under this block's softmax
                const int s = j % KV_STAGES;
                mbar_wait(bar(B_VFULL + s), (j / KV_STAGES) & 1);
                const uint32_t vb = base + OFF_V + s * KV_BYTES;
                for (int t = 0; t < nt; ++t) {
                    mbar_wait(bar(B_PREADY + t), j & 1);   // P_t(j) is in shared memory, S_t(j) has been read, O_t is rescaled
                    tc_fence_after();
                    const uint32_t pb = base + OFF_P + t * P_BYTES;
#pragma unroll
                    for (int ks = 0; ks < BN / 16; ++ks)
                        tc_mma(tmem + 256 + t * D, make_desc(pb + ks * 32, 16, 1024), make_desc(vb + ks * 2048, KV_HALF, 1024), IDESC_PV,
                               (j > 0 || ks > 0) ? 1u : 0u);
                    tc_commit(bar(B_ODONE + t));
                }
                tc_commit(bar(B_VEMPTY + s));
            }
        }
    } else {
        // ---- softmax warpgroups ---------------------------------------------------------------------------------------------
        const int t = warp >> 2;                           // head of the pair
        const int r = tid & 127;                           // query row within the tile = TMEM lane
        if (t == 0 || two) {
            const int wq = warp & 3;
            const uint32_t lane_base = (uint32_t)(wq * 32) << 16;
            const uint32_t tmem_o = tmem + 256 + t * D;
            const int row = mb * BM + r;                   // in-sequence query index
            uint8_t* prow = smem_raw + OFF_P + t * P_BYTES + r * 128;
            float m_ref = -INFINITY, l_sum = 0.f;
            for (int j = 0; j < nblocks; ++j) {
                const int b = j & 1;
                mbar_wait(bar(B_SFULL + t * 2 + b), (j >> 1) & 1);
                tc_fence_after();
                float s[BN];
                {
                    float t0[32], t1[32];
                    tmem_ld32(tmem + lane_base + (t * 2 + b) * BN, t0);
                    tmem_ld32(tmem + lane_base + (t * 2 + b) * BN + 32, t1);
#pragma unroll
                    for (int c = 0; c < 32; ++c) { s[c] = t0[c] * p.scale_log2; s[32 + c] = t1[c] * p.scale_log2; }
                }
                const bool need_mask = (j * BN + BN > len_k) || (j * BN + BN - 1 > mb * BM + wq * 32 + off);
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

                if (j > 0) {                               // PV_t(j-1) finished: P_t may be overwritten, O_t is stable
                    mbar_wait(bar(B_ODONE + t), (j - 1) & 1);
                    tc_fence_after();
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
#pragma unroll
                for (int c8 = 0; c8 < 8; ++c8) {
                    float pv[8];
#pragma unroll
                    for (int e = 0; e < 8; ++e) {
                        // every other score goes through the polynomial: halves the MUFU load, the FMA pipe has room
                        pv[e] = (POLY && (e & 1)) ? exp2_poly(s[c8 * 8 + e] - m_ref) : fast_exp2(s[c8 * 8 + e] - m_ref);
                        l_sum += pv[e];
                    }
                    *reinterpret_cast<uint4*>(prow + ((c8 ^ (r & 7)) << 4)) = pack8(pv);
                }
                if (j * BN + BN > len_k) {
                    // tail block: V rows at keys >= len_k are not this sequence's (stale page rows / the next packed sequence):
                    // P is exactly 0 there but 0 * NaN would poison the accumulator -> zero those rows of the staged tile.  Both
                    // warpgroups do it (same values) so that each one's arrival below covers the tile its own PV MMA reads.
                    const int sv = j % KV_STAGES;
                    mbar_wait(bar(B_VFULL + sv), (j / KV_STAGES) & 1);
                    const int first = len_k - j * BN;                               // 1..63
                    uint8_t* vtile = smem_raw + OFF_V + sv * KV_BYTES;
                    for (int i = r; i < (BN - first) * 16; i += 128) {
                        const int rr = first + (i >> 4), c = i & 15;
                        *reinterpret_cast<uint4*>(vtile + (c >> 3) * KV_HALF + rr * 128 + ((c & 7) << 4)) = make_uint4(0, 0, 0, 0);
                    }
                }
                asm volatile("fence.proxy.async.shared::cta;" ::: "memory");        // P (and V) writes -> visible to the MMA
                tc_fence_before();
                mbar_arrive_ws(bar(B_PREADY + t));
            }
            // ---- epilogue: O / l -> bf16 -> global --------------------------------------------------------------------------
            mbar_wait(bar(B_ODONE + t), (nblocks - 1) & 1);
            tc_fence_after();
            const float inv = l_sum > 0.f ? 1.f / l_sum : 0.f;
            const bool valid = row < len_q;
            __nv_bfloat16* dst = p.out + (int64_t)(q0 + (valid ? row : 0)) * p.out_stride + (head0 + t) * D;
#pragma unroll
            for (int q4 = 0; q4 < 4; ++q4) {
                float o[32];
                tmem_ld32(tmem_o + lane_base + q4 * 32, o);
                if (valid) {
#pragma unroll
                    for (int c8 = 0; c8 < 4; ++c8) {
                        float rr[8];
#pragma unroll
                        for (int e = 0; e < 8; ++e) rr[e] = o[c8 * 8 + e] * inv;
                        *reinterpret_cast<uint4*>(dst + q4 * 32 + c8 * 8) = pack8(rr);
                    }
                }
            }
        }
    }
    tc_fence_before();
    __syncthreads();
    if (warp == 0) asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" ::"r"(tmem), "r"(TMEM_COLS) : "memory");
}

}  // namespace

// Called from b200_paged_prefill (prefill_attn.cu).  Returns B200_EUNSUPPORTED if TMA descriptors cannot be built.
int b200_prefill_ws_launch(b200_ctx* ctx, int layer, const void* q, int64_t q_stride0, const void* k, int64_t k_stride0,
                           const void* v, int64_t v_stride0, const int32_t* cu_q, const int32_t* cu_k,
                           const int32_t* block_tables, int bt_stride, void* out, int64_t out_stride0, int total_q,
                           int num_seqs, int max_seqlen_q, int num_q_heads, int num_kv_heads, float scale,
                           cudaStream_t stream) {
    const int total_k = total_q;
    if ((out_stride0 % 8) || ((uintptr_t)out & 15)) return B200_EINVAL;
    CUtensorMap tq, tk, tv;
    if (!make_map(&tq, q, (uint64_t)num_q_heads * D, (uint64_t)total_q, (uint64_t)q_stride0, BM)) return B200_EUNSUPPORTED;
    WsParams prm;
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
    static const bool poly = [] { const char* e = getenv("B200_PREFILL_POLY"); return e ? atoi(e) != 0 : true; }();
    static B200SmemOptIn optin_poly, optin_mufu;
    const int g_heads = num_q_heads / num_kv_heads;
    dim3 grid((max_seqlen_q + BM - 1) / BM, num_kv_heads * ((g_heads + 1) / 2), num_seqs);
    if (grid.z > 65535 || grid.y > 65535) return B200_EUNSUPPORTED;
    if (poly) {
        B200_CUDA_CHECK(ctx, optin_poly.ensure(prefill_ws_kernel<true>, WS_SMEM));
        B200_LAUNCH((prefill_ws_kernel<true>), grid, WS_THREADS, WS_SMEM, stream, tq, tk, tv, prm);
    } else {
        B200_CUDA_CHECK(ctx, optin_mufu.ensure(prefill_ws_kernel<false>, WS_SMEM));
        B200_LAUNCH((prefill_ws_kernel<false>), grid, WS_THREADS, WS_SMEM, stream, tq, tk, tv, prm);
    }
    return b200_launch_status(ctx);
}
