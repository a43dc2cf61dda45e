// Causal varlen prefill attention over packed or paged K/V for sm_100a.
//
// Replaces the prefill branch of the reference's Attention.forward (nanovllm/layers/attention.py:64-70,
// flash_attn_varlen_func with an optional block_table).  Mask is bottom-right aligned: query i of a
// sequence sees keys j <= i + (len_k - len_q)  (prefix-cache hits and chunked prefill have
// len_q < len_k, engine/model_runner.py:139-146).
//
// Round-1 kernel: flash-attention tiling (64 queries x 64 keys, head_dim 128) on the warp-level
// tensor-core path (mma.sync m16n8k16, SASS HMMA), cp.async double-buffered K/V tiles with an
// XOR-swizzled shared layout, fp32 online softmax with warp-shuffle row reductions, P rounded to
// bf16 for the PV product exactly like the reference's kernel.  The tcgen05/TMEM version of this
// kernel is the next milestone (DESIGN.md "prefill kernel"); prefill attention is < 2 % of the
// benchmark's ideal time, the decode kernel is > 90 %.
#include "common.cuh"

namespace {

constexpr int BM = 64;
constexpr int BN = 64;
constexpr int D = B200_HEAD_DIM;
constexpr int PREFILL_THREADS = 128;
constexpr int TILE_BYTES = BN * D * 2;   // 16 KB

struct PrefillParams {
    const __nv_bfloat16* q;
    int64_t q_stride;
    const __nv_bfloat16* k;
    int64_t k_stride;
    const __nv_bfloat16* v;
    int64_t v_stride;
    const __nv_bfloat16* k_cache;
    const __nv_bfloat16* v_cache;
    const int32_t* cu_q;
    const int32_t* cu_k;
    const int32_t* block_tables;   // null => packed k/v
    int bt_stride;
    __nv_bfloat16* out;
    int64_t out_stride;
    int hq, hkv, block_shift;
    float scale_log2;
};

__device__ __forceinline__ void cp_async16(uint32_t dst, const void* src, bool valid) {
    const int sz = valid ? 16 : 0;   // src-size 0 => zero fill
    asm volatile("cp.async.cg.shared.global [%0], [%1], 16, %2;" ::"r"(dst), "l"(src), "r"(sz) : "memory");
}
__device__ __forceinline__ void cp_async_commit() { asm volatile("cp.async.commit_group;" ::: "memory"); }
template <int N>
__device__ __forceinline__ void cp_async_wait() { asm volatile("cp.async.wait_group %0;" ::"n"(N) : "memory"); }

__device__ __forceinline__ void ldsm_x4(uint32_t addr, uint32_t& r0, uint32_t& r1, uint32_t& r2, uint32_t& r3) {
    asm volatile("ldmatrix.sync.aligned.m8n8.x4.shared.b16 {%0,%1,%2,%3}, [%4];"
                 : "=r"(r0), "=r"(r1), "=r"(r2), "=r"(r3) : "r"(addr));
}
__device__ __forceinline__ void ldsm_x4_t(uint32_t addr, uint32_t& r0, uint32_t& r1, uint32_t& r2, uint32_t& r3) {
    asm volatile("ldmatrix.sync.aligned.m8n8.x4.trans.shared.b16 {%0,%1,%2,%3}, [%4];"
                 : "=r"(r0), "=r"(r1), "=r"(r2), "=r"(r3) : "r"(addr));
}
__device__ __forceinline__ void mma_bf16(float (&c)[4], const uint32_t (&a)[4], uint32_t b0, uint32_t b1) {
    asm volatile(
        "mma.sync.aligned.m16n8k16.row.col.f32.bf16.bf16.f32 {%0,%1,%2,%3}, {%4,%5,%6,%7}, {%8,%9}, {%0,%1,%2,%3};"
        : "+f"(c[0]), "+f"(c[1]), "+f"(c[2]), "+f"(c[3])
        : "r"(a[0]), "r"(a[1]), "r"(a[2]), "r"(a[3]), "r"(b0), "r"(b1));
}

// byte offset of 16-byte chunk `c` of row `r` inside a [rows][128] bf16 tile (256 B rows)
__device__ __forceinline__ uint32_t swz(int r, int c) { return (uint32_t)(r * 256 + ((c ^ (r & 7)) << 4)); }

__global__ void __launch_bounds__(PREFILL_THREADS, 2) prefill_kernel(const PrefillParams p) {
    extern __shared__ __align__(128) uint8_t smem[];
    const uint32_t sQ = smem_u32(smem);
    const uint32_t sK = sQ + TILE_BYTES;            // two buffers
    const uint32_t sV = sK + 2 * TILE_BYTES;        // two buffers

    const int seq = blockIdx.z;
    const int head = blockIdx.y;
    const int mb = gridDim.x - 1 - blockIdx.x;      // longest tiles first
    const int q0 = p.cu_q[seq];
    const int len_q = p.cu_q[seq + 1] - q0;
    const int k0 = p.cu_k[seq];
    const int len_k = p.cu_k[seq + 1] - k0;
    if (mb * BM >= len_q) return;
    const int off = len_k - len_q;                  // causal offset (>= 0)
    const int kvh = head / (p.hq / p.hkv);
    int j_end = mb * BM + BM + off;
    j_end = j_end > len_k ? len_k : j_end;
    const int nblocks = (j_end + BN - 1) / BN;

    const int tid = threadIdx.x;
    const int warp = tid >> 5;
    const int lane = tid & 31;
    const bool paged = p.block_tables != nullptr;
    const int bs_mask = (1 << p.block_shift) - 1;

    auto load_kv = [&](int nb, int buf) {
#pragma unroll
        for (int it = 0; it < 8; ++it) {
            const int id = tid + it * PREFILL_THREADS;   // 1024 chunks of 16 B per tile
            const int r = id >> 4, c = id & 15;
            const int jkey = nb * BN + r;
            const bool ok = jkey < len_k;
            const __nv_bfloat16 *ksrc, *vsrc;
            if (paged) {
                int64_t row = 0;
                if (ok) {
                    const int page = p.block_tables[(int64_t)seq * p.bt_stride + (jkey >> p.block_shift)];
                    row = (((int64_t)page * p.hkv + kvh) << p.block_shift) + (jkey & bs_mask);
                }
                ksrc = p.k_cache + row * D + c * 8;
                vsrc = p.v_cache + row * D + c * 8;
            } else {
                const int64_t t = ok ? (int64_t)(k0 + jkey) : (int64_t)k0;
                ksrc = p.k + t * p.k_stride + kvh * D + c * 8;
                vsrc = p.v + t * p.v_stride + kvh * D + c * 8;
            }
            cp_async16(sK + buf * TILE_BYTES + swz(r, c), ksrc, ok);
            cp_async16(sV + buf * TILE_BYTES + swz(r, c), vsrc, ok);
        }
    };

    // Q tile + first K/V tile
#pragma unroll
    for (int it = 0; it < 8; ++it) {
        const int id = tid + it * PREFILL_THREADS;
        const int r = id >> 4, c = id & 15;
        const int qi = mb * BM + r;
        const bool ok = qi < len_q;
        const __nv_bfloat16* src = p.q + (int64_t)(q0 + (ok ? qi : 0)) * p.q_stride + head * D + c * 8;
        cp_async16(sQ + swz(r, c), src, ok);
    }
    load_kv(0, 0);
    cp_async_commit();

    uint32_t qa[8][4];
    float o[16][4];
#pragma unroll
    for (int n = 0; n < 16; ++n) { o[n][0] = o[n][1] = o[n][2] = o[n][3] = 0.f; }
    float m_row[2] = {-INFINITY, -INFINITY};
    float l_row[2] = {0.f, 0.f};
    const int g = lane >> 2, t4 = lane & 3;
    const int row_base = mb * BM + warp * 16 + g;   // in-sequence query index of c0/c1 (c2/c3: +8)

    for (int nb = 0; nb < nblocks; ++nb) {
        const int buf = nb & 1;
        if (nb + 1 < nblocks) {
            load_kv(nb + 1, buf ^ 1);
            cp_async_commit();
            cp_async_wait<1>();
        } else {
            cp_async_wait<0>();
        }
        __syncthreads();

        if (nb == 0) {
#pragma unroll
            for (int ks = 0; ks < 8; ++ks) {
                const int r = warp * 16 + (lane & 7) + ((lane >> 3) & 1) * 8;
                const int c = ks * 2 + (lane >> 4);
                ldsm_x4(sQ + swz(r, c), qa[ks][0], qa[ks][1], qa[ks][2], qa[ks][3]);
            }
        }

        // ---- S = Q K^T -------------------------------------------------------------------------
        float s[8][4];
#pragma unroll
        for (int n = 0; n < 8; ++n) { s[n][0] = s[n][1] = s[n][2] = s[n][3] = 0.f; }
        const uint32_t kb = sK + buf * TILE_BYTES;
#pragma unroll
        for (int ks = 0; ks < 8; ++ks) {
#pragma unroll
            for (int np = 0; np < 4; ++np) {
                const int mi = lane >> 3;
                const int r = np * 16 + (lane & 7) + (mi >> 1) * 8;
                const int c = ks * 2 + (mi & 1);
                uint32_t b0, b1, b2, b3;
                ldsm_x4(kb + swz(r, c), b0, b1, b2, b3);
                mma_bf16(s[2 * np], qa[ks], b0, b1);
                mma_bf16(s[2 * np + 1], qa[ks], b2, b3);
            }
        }

        // ---- mask + online softmax -------------------------------------------------------------
        const bool need_mask = (nb * BN + BN > len_k) || (nb * BN + BN - 1 > mb * BM + warp * 16 + off);
        float mx[2] = {-INFINITY, -INFINITY};
#pragma unroll
        for (int n = 0; n < 8; ++n) {
#pragma unroll
            for (int e = 0; e < 4; ++e) {
                float v = s[n][e] * p.scale_log2;
                if (need_mask) {
                    const int jj = nb * BN + n * 8 + t4 * 2 + (e & 1);
                    const int ii = row_base + (e >> 1) * 8;
                    if (jj >= len_k || jj > ii + off) v = -INFINITY;
                }
                s[n][e] = v;
                mx[e >> 1] = fmaxf(mx[e >> 1], v);
            }
        }
        float alpha[2], msafe[2];
#pragma unroll
        for (int r = 0; r < 2; ++r) {
            mx[r] = fmaxf(mx[r], __shfl_xor_sync(0xffffffffu, mx[r], 1));
            mx[r] = fmaxf(mx[r], __shfl_xor_sync(0xffffffffu, mx[r], 2));
            const float m_new = fmaxf(m_row[r], mx[r]);
            msafe[r] = (m_new == -INFINITY) ? 0.f : m_new;
            alpha[r] = fast_exp2(m_row[r] - msafe[r]);
            m_row[r] = m_new;
            l_row[r] *= alpha[r];
        }
        uint32_t pa[4][4];
#pragma unroll
        for (int n = 0; n < 8; ++n) {
            const float p0 = fast_exp2(s[n][0] - msafe[0]);
            const float p1 = fast_exp2(s[n][1] - msafe[0]);
            const float p2 = fast_exp2(s[n][2] - msafe[1]);
            const float p3 = fast_exp2(s[n][3] - msafe[1]);
            l_row[0] += p0 + p1;
            l_row[1] += p2 + p3;
            pa[n >> 1][(n & 1) * 2 + 0] = pack_bf16x2(p0, p1);
            pa[n >> 1][(n & 1) * 2 + 1] = pack_bf16x2(p2, p3);
        }
#pragma unroll
        for (int n = 0; n < 16; ++n) {
            o[n][0] *= alpha[0]; o[n][1] *= alpha[0];
            o[n][2] *= alpha[1]; o[n][3] *= alpha[1];
        }

        // ---- O += P V --------------------------------------------------------------------------
        const uint32_t vb = sV + buf * TILE_BYTES;
#pragma unroll
        for (int kk = 0; kk < 4; ++kk) {
#pragma unroll
            for (int np = 0; np < 8; ++np) {
                const int mi = lane >> 3;
                const int r = kk * 16 + (lane & 7) + (mi & 1) * 8;
                const int c = np * 2 + (mi >> 1);
                uint32_t b0, b1, b2, b3;
                ldsm_x4_t(vb + swz(r, c), b0, b1, b2, b3);
                mma_bf16(o[2 * np], pa[kk], b0, b1);
                mma_bf16(o[2 * np + 1], pa[kk], b2, b3);
            }
        }
        __syncthreads();   // both buffers may be refilled from the next iteration on
    }

    // ---- epilogue ------------------------------------------------------------------------------
#pragma unroll
    for (int r = 0; r < 2; ++r) {
        l_row[r] += __shfl_xor_sync(0xffffffffu, l_row[r], 1);
        l_row[r] += __shfl_xor_sync(0xffffffffu, l_row[r], 2);
    }
#pragma unroll
    for (int r = 0; r < 2; ++r) {
        const int qi = row_base + r * 8;
        if (qi < len_q) {
            const float inv = l_row[r] > 0.f ? 1.f / l_row[r] : 0.f;
            __nv_bfloat16* dst = p.out + (int64_t)(q0 + qi) * p.out_stride + head * D + t4 * 2;
#pragma unroll
            for (int n = 0; n < 16; ++n)
                *reinterpret_cast<uint32_t*>(dst + n * 8) = pack_bf16x2(o[n][r * 2] * inv, o[n][r * 2 + 1] * inv);
        }
    }
}

constexpr int PREFILL_SMEM = 5 * TILE_BYTES;   // Q + 2 K + 2 V = 80 KB

}  // namespace

extern "C" int b200_paged_prefill(b200_ctx* ctx, int layer, const void* q, int64_t q_stride0,
                                  const void* k, int64_t k_stride0, const void* v,
                                  int64_t v_stride0, const int32_t* cu_seqlens_q,
                                  const int32_t* cu_seqlens_k, const int32_t* block_tables,
                                  int bt_stride, void* out, int64_t out_stride0, int total_q,
                                  int num_seqs, int max_seqlen_q, int max_seqlen_k,
                                  int num_q_heads, int num_kv_heads, float scale, void* stream) {
    (void)max_seqlen_k;
    if (!ctx || !q || !out || !cu_seqlens_q || !cu_seqlens_k || num_seqs < 0 || total_q < 0) return B200_EINVAL;
    if (num_kv_heads <= 0 || num_q_heads % num_kv_heads) return B200_EINVAL;
    if (total_q == 0 || num_seqs == 0 || max_seqlen_q <= 0) return B200_OK;
    if ((q_stride0 % 8) || (out_stride0 % 2) || ((uintptr_t)q & 15) || ((uintptr_t)out & 3)) return B200_EINVAL;
    PrefillParams prm;
    prm.q = static_cast<const __nv_bfloat16*>(q);
    prm.q_stride = q_stride0;
    prm.k = static_cast<const __nv_bfloat16*>(k);
    prm.k_stride = k_stride0;
    prm.v = static_cast<const __nv_bfloat16*>(v);
    prm.v_stride = v_stride0;
    prm.k_cache = nullptr;
    prm.v_cache = nullptr;
    prm.block_shift = 0;
    if (block_tables) {
        if (!ctx->k_base) return B200_ENOTBOUND;
        if (layer < 0 || layer >= ctx->layers || num_kv_heads != ctx->num_kv_heads) return B200_EINVAL;
        prm.k_cache = ctx->k_layer(layer);
        prm.v_cache = ctx->v_layer(layer);
        prm.block_shift = ctx->block_shift;
    } else {
        if (!k || !v || (k_stride0 % 8) || (v_stride0 % 8) || ((uintptr_t)k & 15) || ((uintptr_t)v & 15)) return B200_EINVAL;
    }
    prm.cu_q = cu_seqlens_q;
    prm.cu_k = cu_seqlens_k;
    prm.block_tables = block_tables;
    prm.bt_stride = bt_stride;
    prm.out = static_cast<__nv_bfloat16*>(out);
    prm.out_stride = out_stride0;
    prm.hq = num_q_heads;
    prm.hkv = num_kv_heads;
    prm.scale_log2 = scale * 1.4426950408889634f;

    static bool configured = false;
    if (!configured) {
        B200_CUDA_CHECK(ctx, cudaFuncSetAttribute(prefill_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, PREFILL_SMEM));
        configured = true;
    }
    dim3 grid((max_seqlen_q + BM - 1) / BM, num_q_heads, num_seqs);
    if (grid.z > 65535 || grid.y > 65535) return B200_EUNSUPPORTED;
    prefill_kernel<<<grid, PREFILL_THREADS, PREFILL_SMEM, static_cast<cudaStream_t>(stream)>>>(prm);
    return b200_launch_status(ctx);
}
