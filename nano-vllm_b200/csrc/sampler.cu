// Token sampling for sm_100a: replaces Sampler.forward (nanovllm/layers/sampler.py:7-12, an
// Inductor-compiled softmax + exponential race) and adds the greedy branch (temperature == 0)
// the north-star requires.  One CTA per row streams the logits once with 128-bit loads and
// keeps a running (value, index) maximum; nothing of size [rows, vocab] is materialised.
#include "common.cuh"

namespace {

constexpr int SAMPLE_THREADS = 256;

__device__ __forceinline__ uint64_t mix64(uint64_t z) {   // splitmix64 finaliser
    z = (z ^ (z >> 30)) * 0xbf58476d1ce4e5b9ull;
    z = (z ^ (z >> 27)) * 0x94d049bb133111ebull;
    return z ^ (z >> 31);
}

struct Best {
    float v;
    int i;
};
__device__ __forceinline__ void take(Best& b, float v, int i) {
    if (v > b.v || (v == b.v && i < b.i)) { b.v = v; b.i = i; }
}

template <bool FP32>
__global__ void __launch_bounds__(SAMPLE_THREADS) sample_kernel(const void* __restrict__ logits, int64_t stride,
                                                                const float* __restrict__ temperatures, int vocab,
                                                                int64_t index_offset, uint64_t seed, uint64_t step,
                                                                const int64_t* __restrict__ step_dev,
                                                                int64_t* out, int64_t* out_keys) {
    const int row = blockIdx.x;
    const float t = temperatures ? temperatures[row] : 0.f;
    const bool greedy = !(t > 0.f);
    const float inv_t = greedy ? 1.f : 1.f / t;
    if (step_dev) step += (uint64_t)*step_dev;
    const uint64_t key = mix64(seed ^ mix64(step * 0x9e3779b97f4a7c15ull + (uint64_t)row));
    Best best{-INFINITY, 0x7fffffff};

    auto score = [&](float x, int idx) -> float {
        if (greedy) return x;
        const uint64_t r = mix64(key + (uint64_t)(index_offset + idx) * 0xd1342543de82ef95ull);
        const float u = ((float)(r >> 40) + 0.5f) * (1.0f / 16777216.0f);      // (0, 1)
        const float e = fmaxf(-logf(u), 1e-10f);                               // Exp(1), clamped like sampler.py:11
        return x * inv_t - logf(e);
    };

    if (FP32) {
        const float* rowp = static_cast<const float*>(logits) + (int64_t)row * stride;
        for (int i = threadIdx.x; i < vocab; i += SAMPLE_THREADS) take(best, score(rowp[i], i), i);
    } else {
        const __nv_bfloat16* rowp = static_cast<const __nv_bfloat16*>(logits) + (int64_t)row * stride;
        const int nvec = vocab >> 3;
        const uint4* r4 = reinterpret_cast<const uint4*>(rowp);
        for (int vi = threadIdx.x; vi < nvec; vi += SAMPLE_THREADS) {
            float f[8];
            unpack8(r4[vi], f);
#pragma unroll
            for (int e = 0; e < 8; ++e) take(best, score(f[e], vi * 8 + e), vi * 8 + e);
        }
        for (int i = nvec * 8 + threadIdx.x; i < vocab; i += SAMPLE_THREADS)
            take(best, score(__bfloat162float(rowp[i]), i), i);
    }

#pragma unroll
    for (int o = 16; o > 0; o >>= 1) {
        const float ov = __shfl_xor_sync(0xffffffffu, best.v, o);
        const int oi = __shfl_xor_sync(0xffffffffu, best.i, o);
        take(best, ov, oi);
    }
    __shared__ float sv[SAMPLE_THREADS / 32];
    __shared__ int si[SAMPLE_THREADS / 32];
    if ((threadIdx.x & 31) == 0) { sv[threadIdx.x >> 5] = best.v; si[threadIdx.x >> 5] = best.i; }
    __syncthreads();
    if (threadIdx.x == 0) {
        Best b{sv[0], si[0]};
        for (int w = 1; w < SAMPLE_THREADS / 32; ++w) take(b, sv[w], si[w]);
        const int64_t tok = index_offset + (b.i == 0x7fffffff ? 0 : b.i);
        if (out) out[row] = tok;
        if (out_keys) {
            uint32_t bits = __float_as_uint(b.v);
            bits ^= (bits >> 31) ? 0xffffffffu : 0x80000000u;            // unsigned order == float order
            const uint64_t k = ((uint64_t)bits << 32) | (uint64_t)(0xffffffffu - (uint32_t)tok);
            out_keys[row] = (int64_t)(k ^ 0x8000000000000000ull);        // signed order == unsigned order
        }
    }
}

}  // namespace

extern "C" int b200_sample(const void* logits, int logits_is_fp32, int64_t logits_stride0,
                           const float* temperatures, int rows, int vocab, int64_t index_offset,
                           uint64_t seed, uint64_t step, const int64_t* step_dev, int64_t* out,
                           int64_t* out_keys, void* stream) {
    if (!logits || (!out && !out_keys) || rows < 0 || vocab <= 0 || index_offset < 0) return B200_EINVAL;
    if (rows == 0) return B200_OK;
    cudaStream_t st = static_cast<cudaStream_t>(stream);
    if (logits_is_fp32) {
        sample_kernel<true><<<rows, SAMPLE_THREADS, 0, st>>>(logits, logits_stride0, temperatures, vocab, index_offset, seed, step, step_dev, out, out_keys);
    } else {
        if (((uintptr_t)logits & 15) || (logits_stride0 % 8)) return B200_EINVAL;
        sample_kernel<false><<<rows, SAMPLE_THREADS, 0, st>>>(logits, logits_stride0, temperatures, vocab, index_offset, seed, step, step_dev, out, out_keys);
    }
    return b200_launch_status(nullptr);
}
