// The small bandwidth-bound ops around the attention operator, hand-written for sm_100a.
// They replace the reference's Triton KV scatter and its five @torch.compile sites (SURVEY.md 2b,
// K1 and K5-K8): fp32 math, one rounding at every stored output (what Inductor generates on a
// GPU), 128-bit coalesced global accesses, warp-shuffle reductions.
#include "common.cuh"

namespace {

constexpr int NORM_THREADS = 128;
constexpr int NORM_MAXV = 8;   // uint4 per thread -> cols <= 128 * 8 * 8 = 8192

__device__ __forceinline__ float block_sum_128(float v, float* red) {
    v = warp_sum(v);
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
    if (lane == 0) red[warp] = v;
    __syncthreads();
    float t = red[0] + red[1] + red[2] + red[3];
    __syncthreads();
    return t;
}

// RMSNorm.rms_forward (layers/layernorm.py:16-26) when HAS_RES == false,
// RMSNorm.add_rms_forward (layers/layernorm.py:28-40) when true.
template <bool HAS_RES>
__global__ void __launch_bounds__(NORM_THREADS) rmsnorm_kernel(const __nv_bfloat16* __restrict__ x, int64_t x_stride,
                                                               __nv_bfloat16* residual,
                                                               const __nv_bfloat16* __restrict__ w,
                                                               __nv_bfloat16* out, int64_t out_stride, int cols,
                                                               float eps) {
    __shared__ float red[4];
    const int row = blockIdx.x;
    const int nvec = cols >> 3;
    const uint4* x4 = reinterpret_cast<const uint4*>(x + (int64_t)row * x_stride);
    uint4* r4 = HAS_RES ? reinterpret_cast<uint4*>(residual + (int64_t)row * cols) : nullptr;
    const uint4* w4 = reinterpret_cast<const uint4*>(w);
    uint4* o4 = reinterpret_cast<uint4*>(out + (int64_t)row * out_stride);

    float v[NORM_MAXV][8];
    float ss = 0.f;
#pragma unroll
    for (int k = 0; k < NORM_MAXV; ++k) {
        const int idx = threadIdx.x + k * NORM_THREADS;
        if (idx < nvec) {
            unpack8(x4[idx], v[k]);
            if (HAS_RES) {
                float r[8];
                unpack8(r4[idx], r);
#pragma unroll
                for (int e = 0; e < 8; ++e) v[k][e] += r[e];
                r4[idx] = pack8(v[k]);            // residual <- bf16(x + residual)
            }
#pragma unroll
            for (int e = 0; e < 8; ++e) ss = fmaf(v[k][e], v[k][e], ss);
        }
    }
    const float tot = block_sum_128(ss, red);
    const float rstd = 1.0f / sqrtf(tot / (float)cols + eps);
#pragma unroll
    for (int k = 0; k < NORM_MAXV; ++k) {
        const int idx = threadIdx.x + k * NORM_THREADS;
        if (idx < nvec) {
            float wf[8], y[8];
            unpack8(w4[idx], wf);
#pragma unroll
            for (int e = 0; e < 8; ++e) y[e] = __fmul_rn(__fmul_rn(v[k][e], rstd), wf[e]);
            o4[idx] = pack8(y);
        }
    }
}

// q_norm / k_norm (models/qwen3.py:82-84) + rotary_emb (layers/rotary_embedding.py:37-48) +
// store_kvcache (layers/attention.py:10-40) over the fused qkv GEMM output.  One warp per
// (token, head); lane l owns elements {2l, 2l+1} of each rotation half.
__global__ void __launch_bounds__(128) qknorm_rope_store_kernel(
    __nv_bfloat16* qkv, int64_t stride, int hq, int hkv, const int64_t* __restrict__ positions,
    const __nv_bfloat16* __restrict__ qw, const __nv_bfloat16* __restrict__ kw,
    const float* __restrict__ cos_sin, float eps, const int32_t* __restrict__ slot_mapping,
    __nv_bfloat16* k_cache, __nv_bfloat16* v_cache, int block_shift, int n) {
    const int heads = hq + 2 * hkv;
    const int64_t gwarp = (int64_t)blockIdx.x * 4 + (threadIdx.x >> 5);
    if (gwarp >= (int64_t)n * heads) return;
    const int tok = (int)(gwarp / heads);
    const int head = (int)(gwarp - (int64_t)tok * heads);
    const int lane = threadIdx.x & 31;
    __nv_bfloat16* src = qkv + (int64_t)tok * stride + head * B200_HEAD_DIM;

    int slot = -1;
    if (slot_mapping != nullptr && k_cache != nullptr) slot = slot_mapping[tok];

    if (head >= hq + hkv) {                       // a value head: scatter only
        if (slot < 0) return;
        const int kvh = head - hq - hkv;
        const int64_t row = ((((int64_t)(slot >> block_shift)) * hkv + kvh) << block_shift) + (slot & ((1 << block_shift) - 1));
        reinterpret_cast<uint2*>(v_cache + row * B200_HEAD_DIM)[lane] = reinterpret_cast<const uint2*>(src)[lane];
        return;
    }

    const bool is_q = head < hq;
    const uint32_t a = reinterpret_cast<const uint32_t*>(src)[lane];         // x1: elements 2l, 2l+1
    const uint32_t b = reinterpret_cast<const uint32_t*>(src + 64)[lane];    // x2: 64+2l, 64+2l+1
    float x1[2] = {bf16lo(a), bf16hi(a)};
    float x2[2] = {bf16lo(b), bf16hi(b)};
    float ss = x1[0] * x1[0];
    ss = fmaf(x1[1], x1[1], ss);
    ss = fmaf(x2[0], x2[0], ss);
    ss = fmaf(x2[1], x2[1], ss);
    ss = warp_sum(ss);
    const float rstd = 1.0f / sqrtf(ss / (float)B200_HEAD_DIM + eps);
    const __nv_bfloat16* w = is_q ? qw : kw;
    const uint32_t wa = reinterpret_cast<const uint32_t*>(w)[lane];
    const uint32_t wb = reinterpret_cast<const uint32_t*>(w + 64)[lane];
    // the reference's norm is its own kernel: its output is a stored bf16 tensor
    x1[0] = round_bf16(__fmul_rn(__fmul_rn(x1[0], rstd), bf16lo(wa)));
    x1[1] = round_bf16(__fmul_rn(__fmul_rn(x1[1], rstd), bf16hi(wa)));
    x2[0] = round_bf16(__fmul_rn(__fmul_rn(x2[0], rstd), bf16lo(wb)));
    x2[1] = round_bf16(__fmul_rn(__fmul_rn(x2[1], rstd), bf16hi(wb)));

    const float* cs = cos_sin + positions[tok] * B200_HEAD_DIM;
    const float2 c = reinterpret_cast<const float2*>(cs)[lane];
    const float2 s = reinterpret_cast<const float2*>(cs + 64)[lane];
    const float y1a = __fsub_rn(__fmul_rn(x1[0], c.x), __fmul_rn(x2[0], s.x));
    const float y1b = __fsub_rn(__fmul_rn(x1[1], c.y), __fmul_rn(x2[1], s.y));
    const float y2a = __fadd_rn(__fmul_rn(x2[0], c.x), __fmul_rn(x1[0], s.x));
    const float y2b = __fadd_rn(__fmul_rn(x2[1], c.y), __fmul_rn(x1[1], s.y));
    const uint32_t o1 = pack_bf16x2(y1a, y1b);
    const uint32_t o2 = pack_bf16x2(y2a, y2b);
    reinterpret_cast<uint32_t*>(src)[lane] = o1;
    reinterpret_cast<uint32_t*>(src + 64)[lane] = o2;
    if (!is_q && slot >= 0) {
        const int kvh = head - hq;
        const int64_t row = ((((int64_t)(slot >> block_shift)) * hkv + kvh) << block_shift) + (slot & ((1 << block_shift) - 1));
        __nv_bfloat16* dst = k_cache + row * B200_HEAD_DIM;
        reinterpret_cast<uint32_t*>(dst)[lane] = o1;
        reinterpret_cast<uint32_t*>(dst + 64)[lane] = o2;
    }
}

// store_kvcache_kernel (layers/attention.py:10-30): one warp per (token, kv head).
__global__ void __launch_bounds__(128) store_kv_kernel(const __nv_bfloat16* __restrict__ k, int64_t k_stride,
                                                       const __nv_bfloat16* __restrict__ v, int64_t v_stride,
                                                       const int32_t* __restrict__ slot_mapping,
                                                       __nv_bfloat16* k_cache, __nv_bfloat16* v_cache, int hkv,
                                                       int block_shift, int n) {
    const int64_t gwarp = (int64_t)blockIdx.x * 4 + (threadIdx.x >> 5);
    if (gwarp >= (int64_t)n * hkv) return;
    const int tok = (int)(gwarp / hkv);
    const int kvh = (int)(gwarp - (int64_t)tok * hkv);
    const int lane = threadIdx.x & 31;
    const int slot = slot_mapping[tok];
    if (slot < 0) return;
    const int64_t row = ((((int64_t)(slot >> block_shift)) * hkv + kvh) << block_shift) + (slot & ((1 << block_shift) - 1));
    reinterpret_cast<uint2*>(k_cache + row * B200_HEAD_DIM)[lane] =
        reinterpret_cast<const uint2*>(k + (int64_t)tok * k_stride + kvh * B200_HEAD_DIM)[lane];
    reinterpret_cast<uint2*>(v_cache + row * B200_HEAD_DIM)[lane] =
        reinterpret_cast<const uint2*>(v + (int64_t)tok * v_stride + kvh * B200_HEAD_DIM)[lane];
}

// SiluAndMul.forward (layers/activation.py:8-11)
__global__ void __launch_bounds__(256) silu_mul_kernel(const __nv_bfloat16* __restrict__ x, __nv_bfloat16* out,
                                                       int64_t total_vec, int inter_vec) {
    for (int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; i < total_vec; i += (int64_t)gridDim.x * blockDim.x) {
        const int64_t row = i / inter_vec;
        const int col = (int)(i - row * inter_vec);
        const uint4* xr = reinterpret_cast<const uint4*>(x) + row * 2 * inter_vec;
        float g[8], u[8], y[8];
        unpack8(xr[col], g);
        unpack8(xr[inter_vec + col], u);
#pragma unroll
        for (int e = 0; e < 8; ++e) y[e] = __fmul_rn(g[e] / (1.0f + expf(-g[e])), u[e]);
        reinterpret_cast<uint4*>(out)[i] = pack8(y);
    }
}

// F.embedding (layers/embed_head.py:38)
__global__ void __launch_bounds__(128) embedding_kernel(const int64_t* __restrict__ ids,
                                                        const __nv_bfloat16* __restrict__ table,
                                                        __nv_bfloat16* out, int hidden_vec) {
    const int tok = blockIdx.x;
    const uint4* src = reinterpret_cast<const uint4*>(table) + ids[tok] * hidden_vec;
    uint4* dst = reinterpret_cast<uint4*>(out) + (int64_t)tok * hidden_vec;
    for (int i = threadIdx.x; i < hidden_vec; i += 128) dst[i] = src[i];
}

__global__ void gather_tokens_kernel(int64_t* ids, const int32_t* __restrict__ src, const int64_t* __restrict__ prev, int n) {
    const int i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i < n) {
        const int s = src[i];
        if (s >= 0) ids[i] = prev[s];
    }
}

inline bool aligned16(const void* p) { return ((uintptr_t)p & 15) == 0; }

}  // namespace

extern "C" int b200_rmsnorm(const void* x, int64_t x_stride0, const void* weight, void* out,
                            int64_t out_stride0, int rows, int cols, float eps, void* stream) {
    if (!x || !weight || !out || rows < 0) return B200_EINVAL;
    if (cols <= 0 || cols % 8 || cols > NORM_THREADS * NORM_MAXV * 8) return B200_EUNSUPPORTED;
    if (x_stride0 % 8 || out_stride0 % 8 || !aligned16(x) || !aligned16(out) || !aligned16(weight)) return B200_EINVAL;
    if (rows == 0) return B200_OK;
    rmsnorm_kernel<false><<<rows, NORM_THREADS, 0, static_cast<cudaStream_t>(stream)>>>(
        static_cast<const __nv_bfloat16*>(x), x_stride0, nullptr, static_cast<const __nv_bfloat16*>(weight),
        static_cast<__nv_bfloat16*>(out), out_stride0, cols, eps);
    return b200_launch_status(nullptr);
}

extern "C" int b200_add_rmsnorm(const void* x, void* residual, const void* weight, void* out,
                                int rows, int cols, float eps, void* stream) {
    if (!x || !residual || !weight || !out || rows < 0) return B200_EINVAL;
    if (cols <= 0 || cols % 8 || cols > NORM_THREADS * NORM_MAXV * 8) return B200_EUNSUPPORTED;
    if (!aligned16(x) || !aligned16(out) || !aligned16(weight) || !aligned16(residual)) return B200_EINVAL;
    if (rows == 0) return B200_OK;
    rmsnorm_kernel<true><<<rows, NORM_THREADS, 0, static_cast<cudaStream_t>(stream)>>>(
        static_cast<const __nv_bfloat16*>(x), cols, static_cast<__nv_bfloat16*>(residual),
        static_cast<const __nv_bfloat16*>(weight), static_cast<__nv_bfloat16*>(out), cols, cols, eps);
    return b200_launch_status(nullptr);
}

extern "C" int b200_qknorm_rope_store(b200_ctx* ctx, int layer, void* qkv, int64_t qkv_stride0,
                                      int num_q_heads, int num_kv_heads, const int64_t* positions,
                                      const void* q_norm_weight, const void* k_norm_weight,
                                      const float* cos_sin, float eps, const int32_t* slot_mapping,
                                      int n, void* stream) {
    if (!ctx || !qkv || !positions || !q_norm_weight || !k_norm_weight || !cos_sin || n < 0) return B200_EINVAL;
    if (qkv_stride0 % 8 || !aligned16(qkv)) return B200_EINVAL;
    if (n == 0) return B200_OK;
    __nv_bfloat16 *kc = nullptr, *vc = nullptr;
    int shift = 0;
    if (slot_mapping != nullptr && ctx->k_base != nullptr) {
        if (layer < 0 || layer >= ctx->layers || num_kv_heads != ctx->num_kv_heads) return B200_EINVAL;
        kc = ctx->k_layer(layer);
        vc = ctx->v_layer(layer);
        shift = ctx->block_shift;
    }
    const int64_t warps = (int64_t)n * (num_q_heads + 2 * num_kv_heads);
    const unsigned blocks = (unsigned)((warps + 3) / 4);
    qknorm_rope_store_kernel<<<blocks, 128, 0, static_cast<cudaStream_t>(stream)>>>(
        static_cast<__nv_bfloat16*>(qkv), qkv_stride0, num_q_heads, num_kv_heads, positions,
        static_cast<const __nv_bfloat16*>(q_norm_weight), static_cast<const __nv_bfloat16*>(k_norm_weight),
        cos_sin, eps, slot_mapping, kc, vc, shift, n);
    return b200_launch_status(ctx);
}

extern "C" int b200_store_kv(b200_ctx* ctx, int layer, const void* k, int64_t k_stride0,
                             const void* v, int64_t v_stride0, const int32_t* slot_mapping, int n,
                             void* stream) {
    if (!ctx || !k || !v || !slot_mapping || n < 0) return B200_EINVAL;
    if (!ctx->k_base) return B200_ENOTBOUND;
    if (layer < 0 || layer >= ctx->layers) return B200_EINVAL;
    if (k_stride0 % 4 || v_stride0 % 4 || ((uintptr_t)k & 7) || ((uintptr_t)v & 7)) return B200_EINVAL;
    if (n == 0) return B200_OK;
    const int64_t warps = (int64_t)n * ctx->num_kv_heads;
    store_kv_kernel<<<(unsigned)((warps + 3) / 4), 128, 0, static_cast<cudaStream_t>(stream)>>>(
        static_cast<const __nv_bfloat16*>(k), k_stride0, static_cast<const __nv_bfloat16*>(v), v_stride0,
        slot_mapping, ctx->k_layer(layer), ctx->v_layer(layer), ctx->num_kv_heads, ctx->block_shift, n);
    return b200_launch_status(ctx);
}

extern "C" int b200_silu_mul(const void* x, void* out, int rows, int inter, void* stream) {
    if (!x || !out || rows < 0) return B200_EINVAL;
    if (inter <= 0 || inter % 8 || !aligned16(x) || !aligned16(out)) return B200_EINVAL;
    if (rows == 0) return B200_OK;
    const int64_t total = (int64_t)rows * (inter / 8);
    int64_t blocks = (total + 255) / 256;
    if (blocks > 148 * 16) blocks = 148 * 16;
    silu_mul_kernel<<<(unsigned)blocks, 256, 0, static_cast<cudaStream_t>(stream)>>>(
        static_cast<const __nv_bfloat16*>(x), static_cast<__nv_bfloat16*>(out), total, inter / 8);
    return b200_launch_status(nullptr);
}

extern "C" int b200_embedding(const int64_t* ids, const void* table, void* out, int n, int hidden,
                              void* stream) {
    if (!ids || !table || !out || n < 0) return B200_EINVAL;
    if (hidden <= 0 || hidden % 8 || !aligned16(table) || !aligned16(out)) return B200_EINVAL;
    if (n == 0) return B200_OK;
    embedding_kernel<<<n, 128, 0, static_cast<cudaStream_t>(stream)>>>(
        ids, static_cast<const __nv_bfloat16*>(table), static_cast<__nv_bfloat16*>(out), hidden / 8);
    return b200_launch_status(nullptr);
}

extern "C" int b200_gather_tokens(int64_t* ids, const int32_t* src, const int64_t* prev_tokens, int n, void* stream) {
    if (!ids || !src || !prev_tokens || n < 0) return B200_EINVAL;
    if (n == 0) return B200_OK;
    gather_tokens_kernel<<<(n + 255) / 256, 256, 0, static_cast<cudaStream_t>(stream)>>>(ids, src, prev_tokens, n);
    return b200_launch_status(nullptr);
}
