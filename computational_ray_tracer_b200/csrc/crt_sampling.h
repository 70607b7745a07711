// crt_sampling.h -- counter-based sampler stack, host + device.
//
// The reference's samplers (ThirdParty/pbrv4/samplers.h:25-136) are stateful objects, one per CPU thread.
// Here a sampler is a 32-byte value re-seeded per path from (pixel, sampleIndex, dimension, seed), which is
// what makes the multi-GPU partition invisible in the image: a sample's random numbers depend only on
// its own coordinates.  Integer code follows hash.h:18-63, HelperFunctions.h:137-203 and rng.h:24-162.
#pragma once
#include "crt_math.h"

namespace crt {

// ---- MurmurHash64A over 12 / 16 packed bytes (hash.h:18-63, :96-104) ------------------------------
CRT_HD uint64_t murmur_mix_block(uint64_t h, uint64_t k) {
    const uint64_t m = 0xc6a4a7935bd1e995ull;
    k *= m; k ^= k >> 47; k *= m;
    h ^= k; h *= m;
    return h;
}
CRT_HD uint64_t murmur_finish(uint64_t h) {
    const uint64_t m = 0xc6a4a7935bd1e995ull;
    h ^= h >> 47; h *= m; h ^= h >> 47;
    return h;
}
// generic byte-buffer version (host-side known-answer tests and odd lengths)
CRT_HD uint64_t murmur64a(const unsigned char* key, uint64_t len, uint64_t seed) {
    const uint64_t m = 0xc6a4a7935bd1e995ull;
    uint64_t h = seed ^ (len * m);
    uint64_t nblocks = len / 8;
    for (uint64_t b = 0; b < nblocks; ++b) {
        uint64_t k = 0;
        for (int i = 7; i >= 0; --i) k = (k << 8) | key[8 * b + i];
        h = murmur_mix_block(h, k);
    }
    const unsigned char* tail = key + 8 * nblocks;
    int rem = (int)(len & 7);
    for (int i = rem - 1; i >= 0; --i) h ^= (uint64_t)tail[i] << (8 * i);
    if (rem) h *= m;
    return murmur_finish(h);
}
// Hash(ivec2 p, int seed): 12 bytes = one 8-byte block (x | y<<32) + 4 tail bytes
CRT_HD uint64_t hash_pixel_seed(int x, int y, int seed) {
    const uint64_t m = 0xc6a4a7935bd1e995ull;
    uint64_t h = 0 ^ (12ull * m);
    h = murmur_mix_block(h, (uint64_t)(uint32_t)x | ((uint64_t)(uint32_t)y << 32));
    h ^= (uint64_t)(uint32_t)seed;       // tail bytes 3..0
    h *= m;
    return murmur_finish(h);
}
// Hash(ivec2 p, int dim, int seed): 16 bytes = two blocks
CRT_HD uint64_t hash_pixel_dim_seed(int x, int y, int dim, int seed) {
    const uint64_t m = 0xc6a4a7935bd1e995ull;
    uint64_t h = 0 ^ (16ull * m);
    h = murmur_mix_block(h, (uint64_t)(uint32_t)x | ((uint64_t)(uint32_t)y << 32));
    h = murmur_mix_block(h, (uint64_t)(uint32_t)dim | ((uint64_t)(uint32_t)seed << 32));
    return murmur_finish(h);
}
CRT_HD uint64_t mix_bits(uint64_t v) {                  // hash.h:67-74
    v ^= (v >> 31); v *= 0x7fb5d329728ea185ull;
    v ^= (v >> 27); v *= 0x81dadef4bc2dd44dull;
    v ^= (v >> 33);
    return v;
}
CRT_HD int permutation_element(uint32_t i, uint32_t l, uint32_t p) {     // HelperFunctions.h:175-203
    uint32_t w = l - 1;
    w |= w >> 1; w |= w >> 2; w |= w >> 4; w |= w >> 8; w |= w >> 16;
    do {
        i ^= p;             i *= 0xe170893d;
        i ^= p >> 16;       i ^= (i & w) >> 4;
        i ^= p >> 8;        i *= 0x0929eb3f;
        i ^= p >> 23;       i ^= (i & w) >> 1;
        i *= 1 | p >> 27;   i *= 0x6935fa69;
        i ^= (i & w) >> 11; i *= 0x74dcb303;
        i ^= (i & w) >> 2;  i *= 0x9e501cc3;
        i ^= (i & w) >> 2;  i *= 0xc860a3df;
        i &= w;
        i ^= i >> 5;
    } while (i >= l);
    return (int)((i + p) % l);
}

// ---- PCG32 (rng.h:24-162) ------------------------------------------------------------------------
struct Pcg32 {
    uint64_t state, inc;
};
#define CRT_PCG_MULT 0x5851f42d4c957f2dULL
CRT_HD uint32_t pcg_next_u32(Pcg32& r) {
    uint64_t old = r.state;
    r.state = old * CRT_PCG_MULT + r.inc;
    uint32_t xs = (uint32_t)(((old >> 18u) ^ old) >> 27u);
    uint32_t rot = (uint32_t)(old >> 59u);
    return (xs >> rot) | (xs << ((~rot + 1u) & 31));
}
// Uniform<float>: min(OneMinusEpsilon, u32 * 2^-32) with the reference's OneMinusEpsilon == 1.0f (pch.h:37)
CRT_HD float pcg_next_float(Pcg32& r) {
    float v = (float)pcg_next_u32(r) * 0x1p-32f;
    return min_std(1.0f, v);
}
CRT_HD void pcg_set_sequence(Pcg32& r, uint64_t seq, uint64_t seed) {
    r.state = 0u;
    r.inc = (seq << 1u) | 1u;
    pcg_next_u32(r);
    r.state += seed;
    pcg_next_u32(r);
}
CRT_HD void pcg_advance(Pcg32& r, int64_t idelta) {
    uint64_t curMult = CRT_PCG_MULT, curPlus = r.inc, accMult = 1u, accPlus = 0u, delta = (uint64_t)idelta;
    while (delta > 0) {
        if (delta & 1) { accMult *= curMult; accPlus = accPlus * curMult + curPlus; }
        curPlus = (curMult + 1) * curPlus;
        curMult *= curMult;
        delta /= 2;
    }
    r.state = accMult * r.state + accPlus;
}

// ---- Sampler value type (samplers.h:38-136) ------------------------------------------------------
struct SamplerCfg {
    int kind;       // 0 independent, 1 stratified
    int xs, ys;
    int jitter;
    int seed;
};
struct SamplerState {
    Pcg32 rng;
    int px, py;
    int sampleIndex, dimension;
};
CRT_HD int sampler_spp(const SamplerCfg& c) { return c.xs * c.ys; }
CRT_HD void sampler_start(const SamplerCfg& c, SamplerState& s, int px, int py, int index, int dim) {
    // StratifiedSampler with jitter off refuses index >= spp and keeps its previous state (samplers.h:83-87);
    // the stateless device version has no previous state, so the host rejects that configuration up front.
    s.px = px; s.py = py; s.sampleIndex = index; s.dimension = dim;
    uint64_t seq = hash_pixel_seed(px, py, c.seed);
    pcg_set_sequence(s.rng, seq, mix_bits(seq));
    pcg_advance(s.rng, (int64_t)((uint64_t)index * 65536ull + (uint64_t)dim));
}
CRT_HD float sampler_get1d(const SamplerCfg& c, SamplerState& s) {
    if (c.kind == 0) return pcg_next_float(s.rng);
    uint64_t h = hash_pixel_dim_seed(s.px, s.py, s.dimension, c.seed);
    int spp = sampler_spp(c);
    int stratum = permutation_element((uint32_t)s.sampleIndex, (uint32_t)spp, (uint32_t)h);
    ++s.dimension;
    float delta = c.jitter ? pcg_next_float(s.rng) : 0.5f;
    return ((float)stratum + delta) / (float)spp;
}
CRT_HD f2 sampler_get2d(const SamplerCfg& c, SamplerState& s) {
    f2 r;
    if (c.kind == 0) { r.x = pcg_next_float(s.rng); r.y = pcg_next_float(s.rng); return r; }
    int spp = sampler_spp(c);
    if (s.sampleIndex >= spp) { r.x = 0; r.y = 0; return r; }          // samplers.h:109-112
    uint64_t h = hash_pixel_dim_seed(s.px, s.py, s.dimension, c.seed);
    int stratum = permutation_element((uint32_t)s.sampleIndex, (uint32_t)spp, (uint32_t)h);
    s.dimension += 2;
    int x = stratum % c.xs, y = stratum / c.xs;
    float dx = c.jitter ? pcg_next_float(s.rng) : 0.5f;
    float dy = c.jitter ? pcg_next_float(s.rng) : 0.5f;
    r.x = ((float)x + dx) / (float)c.xs;
    r.y = ((float)y + dy) / (float)c.ys;
    return r;
}

// ---- sampling routines (RayTracer/Sampling.h) ------------------------------------------------------
CRT_HD float sample_linear(float u, float a, float b) {                  // Sampling.h:205-211
    if (u == 0 && a == 0) return 0;
    float x = (u * (a + b)) / (a + sqrtf(lerp_pbrt(u, a * a, b * b)));
    return min_std(x, 1.0f);
}
// Deterministic tent sampling: the coin is u < 0.5, u is remapped to the chosen half (pbrt-v4's rule).  The
// reference flips a non-reproducible global mt19937 coin (Sampling.h:228-235); see DESIGN.md "deviations".
CRT_HD float sample_tent(float u, float r) {
    float up = u;
    if (up == 1.0f) up = 0x1.fffffep-1f;
    if (up < 0.5f) {
        float ur = min_std(up / 0.5f, 1.0f);
        return -r + r * sample_linear(ur, 0, 1);
    }
    float ur = min_std((up - 0.5f) / 0.5f, 1.0f);
    return r * sample_linear(ur, 1, 0);
}
struct FilterSample { float px, py, weight; };
// GaussianFilter (filters.h:96-163) state: the two tabulated CDFs of Continuous_Inversion_Sampler (RayTracer/Sampling.h:781-806),
// built on the host by gaussian_filter_build (same libm as a CPU build of the reference) and read-only on the device
#define CRT_GAUSS_N 10000
struct GaussFilter { const float* cdf_x; const float* cdf_y; float sigma, exp_x, exp_y; };
CRT_HD float gaussian_pbrt(float x, float sigma) {                       // helpers.h:221-225 with mu = 0
    const float Pi = 3.14159265358979323846f;
    return 1.0f / sqrtf(2 * Pi * sigma * sigma) * expf(-powf(x, 2.0f) / (2 * sigma * sigma));
}
CRT_HD float inversion_sample(const float* cdf, float a, float b, float U) {   // Continuous_Inversion_Sampler::Sample, Sampling.h:809-848
    const int N = CRT_GAUSS_N;
    int index = -1, low = 0, high = N;
    while (low <= high) {
        int mid = (int)(low + (high - low) / 2.0f);
        if (mid < N && cdf[mid] < U && U <= cdf[mid + 1]) { index = mid; break; }
        if (cdf[mid] < U) low = mid + 1;
        else high = mid - 1;
    }
    if (index == -1) return 0;                                           // U == 0: "couldn't find index"
    float t = (U - cdf[index]) / (cdf[index + 1] - cdf[index]);
    t = t < 0.0f ? 0.0f : (1.f < t ? 1.f : t);                           // std::clamp
    float delta_x = (b - a) / (float)N;
    return (a + delta_x * index) + t * (delta_x * (index + 1) - delta_x * index);
}
CRT_HD FilterSample filter_sample(int kind, float rx, float ry, f2 u, const GaussFilter& g) {  // filters.h:83-87, :129-135, :285-290
    FilterSample fs;
    fs.weight = 1.0f;
    if (kind == 0) { fs.px = lerp_pbrt(u.x, -rx, rx); fs.py = lerp_pbrt(u.y, -ry, ry); }
    else if (kind == 2) {
        fs.px = inversion_sample(g.cdf_x, -rx, rx, u.x);
        fs.py = inversion_sample(g.cdf_y, -ry, ry, u.y);
        // Evaluate(p) / (PDF_x(p.x) * PDF_y(p.y)): 1 wherever the filter is positive, NaN on its zero boundary, as in the reference
        float ex = fmaxf(0.0f, gaussian_pbrt(fs.px, g.sigma) - g.exp_x), ey = fmaxf(0.0f, gaussian_pbrt(fs.py, g.sigma) - g.exp_y);
        fs.weight = (ex * ey) / (ex * ey);
    }
    else { fs.px = sample_tent(u.x, rx); fs.py = sample_tent(u.y, ry); }
    return fs;
}

}  // namespace crt
