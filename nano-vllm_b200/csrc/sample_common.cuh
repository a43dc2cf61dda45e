// Scoring shared by the sampling kernels (sampler.cu, and the fused LM-head epilogue of linear_tc.cu): counter-based RNG,
// the exponential-race score of Sampler.forward (nanovllm/layers/sampler.py:7-12) and the order-preserving key packing.
#pragma once
#include "common.cuh"

namespace b200sample {

__device__ __forceinline__ uint64_t mix64(uint64_t z) {   // splitmix64 finaliser
    z = (z ^ (z >> 30)) * 0xbf58476d1ce4e5b9ull;
    z = (z ^ (z >> 27)) * 0x94d049bb133111ebull;
    return z ^ (z >> 31);
}
__device__ __forceinline__ uint32_t mix32(uint32_t x) {   // 32-bit avalanche (two multiplies)
    x ^= x >> 16; x *= 0x7feb352du;
    x ^= x >> 15; x *= 0x846ca68bu;
    x ^= x >> 16;
    return x;
}

struct Best {
    float v;
    int i;
};
__device__ __forceinline__ void take(Best& b, float v, int i) {
    if (v > b.v || (v == b.v && i < b.i)) { b.v = v; b.i = i; }
}
// order-preserving packing: larger key <=> larger score, then lower index
__device__ __forceinline__ unsigned long long pack_key(float v, uint32_t idx) {
    uint32_t bits = __float_as_uint(v);
    bits ^= (bits >> 31) ? 0xffffffffu : 0x80000000u;
    return ((unsigned long long)bits << 32) | (unsigned long long)(0xffffffffu - idx);
}

// Per-row sampling state: temperature handling and the RNG key of (seed, step, row).
struct RowSampler {
    bool greedy;
    float inv_t;
    uint32_t k0, k1;
    __device__ __forceinline__ RowSampler(float t, uint64_t seed, uint64_t step, int row) {
        greedy = !(t > 0.f);
        inv_t = greedy ? 1.f : 1.f / t;
        const uint64_t key64 = mix64(seed ^ mix64(step * 0x9e3779b97f4a7c15ull + (uint64_t)row));
        k0 = (uint32_t)key64;
        k1 = (uint32_t)(key64 >> 32);
    }
    // Exponential race (sampler.py:10-11): argmax softmax(x/t)_j / E_j == argmax x_j/t - log E_j, E_j = -log u_j.
    // `global_idx` is the vocabulary id (shard offset included), so that every tensor-parallel layout draws the same noise.
    __device__ __forceinline__ float score(float x, int64_t global_idx) const {
        if (greedy) return x;
        const uint32_t r = mix32(((uint32_t)global_idx ^ k0) * 0x9e3779b1u + k1);
        const float u = ((float)(r >> 8) + 0.5f) * (1.0f / 16777216.0f);       // (0, 1)
        const float e = fmaxf(-__logf(u), 1e-10f);                             // clamp as in sampler.py:11
        return fmaf(x, inv_t, -__logf(e));
    }
};

}  // namespace b200sample
