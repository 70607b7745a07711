// ORACLE (test infrastructure, never shipped, never on the product path).
//
// CPU restatement of the edited pbrt-v4 subset the reference's render loop touches:
// bit/float helpers, MurmurHash64A, PCG32, Independent/Stratified samplers, wavelength
// sampling, spectra, colour space, XYZ pixel sensor and pixel filters.  Each block cites
// the reference file:line it follows (paths relative to /root/reference).
//
// PARITY STATUS: pinned.  Integer functions (hash, PCG32, PermutationElement) by known-answer
// vectors of the published algorithms (tests/test_cpu_kat.py); everything here, floating point
// included, by the outputs of the reference's own sources compiled unmodified into oracle/_ref
// (ref_harness.cpp; tests/test_cpu_ref_pin.py, tests/golden/ref_pin.npz): bit-identical.  Open
// only below the glm boundary (ovec.h) and in the platform libm.
#pragma once
#include <algorithm>
#include <array>
#include <cmath>
#include <cstdint>
#include <cstring>
#include <memory>
#include <string>
#include <vector>

#include "ovec.h"

namespace orc {

// ---- pch.h:37-45 -------------------------------------------------------------------
// NB: 1 - FLT_MIN rounds to exactly 1.0f, so Uniform<float>() may return 1.
constexpr float OneMinusEpsilon = 1 - std::numeric_limits<float>::min();
constexpr float Pi = 3.14159265358979323846;
constexpr float InvPi = 0.31830988618379067154;
constexpr float PiOver2 = 1.57079632679489661923;
constexpr float PiOver4 = 0.78539816339744830961;

// ---- ThirdParty/pbrv4/helpers.h:50-71,154-178 -----------------------------------------
constexpr float MachineEpsilon = std::numeric_limits<float>::epsilon() * 0.5;
inline constexpr float gamma_n(int n) { return (n * MachineEpsilon) / (1 - n * MachineEpsilon); }
inline float DifferenceOfProducts(float a, float b, float c, float d) {
    float cd = c * d;
    float dop = std::fma(a, b, -cd);
    float err = std::fma(-c, d, cd);
    return dop + err;
}
inline int MaxComponentIndex(vec3 t) { return (t.x > t.y) ? ((t.x > t.z) ? 0 : 2) : ((t.y > t.z) ? 1 : 2); }
inline float MaxComponentValue(vec3 t) { return std::max({t.x, t.y, t.z}); }
inline float Lerp(float x, float a, float b) { return (1 - x) * a + x * b; }
inline float SafeSqrt(float x) { return std::sqrt(std::max(0.f, x)); }
template <typename Pred>
inline size_t FindInterval(size_t sz, const Pred& pred) {
    using ssz = std::make_signed_t<size_t>;
    ssz size = (ssz)sz - 2, first = 1;
    while (size > 0) {
        size_t half = (size_t)size >> 1, middle = first + half;
        bool r = pred(middle);
        first = r ? middle + 1 : first;
        size = r ? size - (half + 1) : half;
    }
    ssz v = first - 1, hi = (ssz)sz - 2;
    return (size_t)(v < 0 ? 0 : (v > hi ? hi : v));
}

// ---- ThirdParty/pbrv4/hash.h:18-63,67-74,96-104; Util/HelperFunctions.h:137-203 -----------
inline uint64_t MurmurHash64A(const unsigned char* key, size_t len, uint64_t seed) {
    const uint64_t m = 0xc6a4a7935bd1e995ull;
    const int r = 47;
    uint64_t h = seed ^ (len * m);
    const unsigned char* end = key + 8 * (len / 8);
    while (key != end) {
        uint64_t k;
        std::memcpy(&k, key, 8);
        key += 8;
        k *= m; k ^= k >> r; k *= m;
        h ^= k; h *= m;
    }
    size_t tail = len & 7;
    for (size_t i = tail; i-- > 0;) h ^= uint64_t(key[i]) << (8 * i);
    if (tail) h *= m;
    h ^= h >> r; h *= m; h ^= h >> r;
    return h;
}
inline uint64_t MixBits(uint64_t v) {
    v ^= (v >> 31); v *= 0x7fb5d329728ea185ull;
    v ^= (v >> 27); v *= 0x81dadef4bc2dd44dull;
    v ^= (v >> 33);
    return v;
}
// Hash(ivec2 p, int seed): 12 packed bytes; Hash(ivec2 p, int dim, int seed): 16 bytes
inline uint64_t HashPixelSeed(ivec2 p, int seed) {
    int32_t buf[4] = {p.x, p.y, seed, 0};
    return MurmurHash64A((const unsigned char*)buf, 12, 0);
}
inline uint64_t HashPixelDimSeed(ivec2 p, int dim, int seed) {
    int32_t buf[4] = {p.x, p.y, dim, seed};
    return MurmurHash64A((const unsigned char*)buf, 16, 0);
}
inline int PermutationElement(uint32_t i, uint32_t l, uint32_t p) {
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
    return (i + p) % l;
}

// ---- ThirdParty/pbrv4/rng.h:24-162 --------------------------------------------------------
struct RNG {
    static constexpr uint64_t kMult = 0x5851f42d4c957f2dULL;
    uint64_t state = 0x853c49e6748fea9bULL, inc = 0xda3e39cb94b95bdbULL;
    RNG() = default;
    explicit RNG(uint64_t seq) { SetSequence(seq); }
    void SetSequence(uint64_t seq, uint64_t seed) {
        state = 0u;
        inc = (seq << 1u) | 1u;
        UniformU32();
        state += seed;
        UniformU32();
    }
    void SetSequence(uint64_t seq) { SetSequence(seq, MixBits(seq)); }
    uint32_t UniformU32() {
        uint64_t old = state;
        state = old * kMult + inc;
        uint32_t xs = (uint32_t)(((old >> 18u) ^ old) >> 27u);
        uint32_t rot = (uint32_t)(old >> 59u);
        return (xs >> rot) | (xs << ((~rot + 1u) & 31));
    }
    float UniformFloat() { return std::min<float>(OneMinusEpsilon, UniformU32() * 0x1p-32f); }
    void Advance(int64_t idelta) {
        uint64_t curMult = kMult, curPlus = inc, accMult = 1u, accPlus = 0u, delta = (uint64_t)idelta;
        while (delta > 0) {
            if (delta & 1) { accMult *= curMult; accPlus = accPlus * curMult + curPlus; }
            curPlus = (curMult + 1) * curPlus;
            curMult *= curMult;
            delta /= 2;
        }
        state = accMult * state + accPlus;
    }
};

// ---- ThirdParty/pbrv4/samplers.h:25-136 ------------------------------------------------------
struct Sampler {
    virtual ~Sampler() = default;
    virtual int SamplesPerPixel() const = 0;
    virtual void StartPixelSample(ivec2 p, int sampleIndex, int dimension = 0) = 0;
    virtual float Get1D() = 0;
    virtual vec2 Get2D() = 0;
    virtual vec2 GetPixel2D() = 0;
    virtual std::unique_ptr<Sampler> Clone() const = 0;
};
struct IndependentSampler : Sampler {
    int spp, seed; RNG rng;
    IndependentSampler(int spp_, int seed_ = 0) : spp(spp_), seed(seed_) {}
    int SamplesPerPixel() const override { return spp; }
    void StartPixelSample(ivec2 p, int idx, int dim) override {
        rng.SetSequence(HashPixelSeed(p, seed));
        rng.Advance(idx * 65536ull + dim);
    }
    float Get1D() override { return rng.UniformFloat(); }
    vec2 Get2D() override { float a = rng.UniformFloat(); float b = rng.UniformFloat(); return {a, b}; }
    vec2 GetPixel2D() override { return Get2D(); }
    std::unique_ptr<Sampler> Clone() const override { return std::make_unique<IndependentSampler>(spp, seed); }
};
struct StratifiedSampler : Sampler {
    int xs, ys, seed; bool jitter; RNG rng; ivec2 pixel; int sampleIndex = 0, dimension = 0;
    StratifiedSampler(int x, int y, bool jit, int seed_ = 0) : xs(x), ys(y), seed(seed_), jitter(jit) {}
    int SamplesPerPixel() const override { return xs * ys; }
    void StartPixelSample(ivec2 p, int index, int dim) override {
        if (jitter == false && index >= SamplesPerPixel()) return;  // samplers.h:83-87 (keeps stale state)
        pixel = p; sampleIndex = index; dimension = dim;
        rng.SetSequence(HashPixelSeed(p, seed));
        rng.Advance(sampleIndex * 65536ull + dimension);
    }
    float Get1D() override {
        uint64_t h = HashPixelDimSeed(pixel, dimension, seed);
        int stratum = PermutationElement(sampleIndex, SamplesPerPixel(), (uint32_t)h);
        ++dimension;
        float delta = jitter ? rng.UniformFloat() : 0.5f;
        return (stratum + delta) / SamplesPerPixel();
    }
    vec2 Get2D() override {
        if (sampleIndex >= SamplesPerPixel()) return vec2(0, 0);
        uint64_t h = HashPixelDimSeed(pixel, dimension, seed);
        int stratum = PermutationElement(sampleIndex, SamplesPerPixel(), (uint32_t)h);
        dimension += 2;
        int x = stratum % xs, y = stratum / xs;
        float dx = jitter ? rng.UniformFloat() : 0.5f;
        float dy = jitter ? rng.UniformFloat() : 0.5f;
        return {(x + dx) / xs, (y + dy) / ys};
    }
    vec2 GetPixel2D() override { return Get2D(); }
    std::unique_ptr<Sampler> Clone() const override { return std::make_unique<StratifiedSampler>(xs, ys, jitter, seed); }
};

// ---- RayTracer/Sampling.h:63-71,205-211,228-235,383-403,449-459 ------------------------------
inline float VisibleWavelengthsPDF(float lambda) {
    if (lambda < 360 || lambda > 830) return 0;
    // std::pow(float, int) promotes to double (C++11 [c.math]); the quotient is double, narrowed on return
    return 0.0039398042f / std::pow(std::cosh(0.0072f * (lambda - 538)), 2);
}
inline float SampleVisibleWavelengths(float u) {
    return 538 - 138.888889f * std::atanh(0.85691062f - 1.82750197f * u);
}
inline float SampleLinear(float u, float a, float b) {
    if (u == 0 && a == 0) return 0;
    float x = (u * (a + b)) / (a + std::sqrt(Lerp(u, a * a, b * b)));
    return std::min(x, OneMinusEpsilon);
}
// DOCUMENTED DEVIATION: the reference flips the left/right coin with a global,
// random_device-seeded mt19937 (Sampling.h:228-235 -> :100-102), so its TriangleFilter is not
// reproducible.  The oracle uses upstream pbrt-v4's deterministic rule: the coin is u < 0.5 and
// u is remapped to the chosen half.
inline float SampleTent(float u, float r) {
    float up = u;
    if (up == 1.0f) up = std::nextafter(1.0f, 0.0f);
    if (up < 0.5f) {
        float ur = std::min(up / 0.5f, OneMinusEpsilon);
        return -r + r * SampleLinear(ur, 0, 1);
    }
    float ur = std::min((up - 0.5f) / 0.5f, OneMinusEpsilon);
    return r * SampleLinear(ur, 1, 0);
}
inline vec2 SampleUniformDiskConcentric(vec2 u) {
    vec2 uo = vec2(2, 2) * u - vec2(1, 1);
    if (uo.x == 0 && uo.y == 0) return {0, 0};
    float theta, r;
    if (std::abs(uo.x) > std::abs(uo.y)) { r = uo.x; theta = PiOver4 * (uo.y / uo.x); }
    else { r = uo.y; theta = PiOver2 - PiOver4 * (uo.x / uo.y); }
    return r * vec2(std::cos(theta), std::sin(theta));
}
inline vec3 SampleCosineHemisphere(vec2 u) {
    vec2 d = SampleUniformDiskConcentric(u);
    float z = SafeSqrt(1 - d.x * d.x - d.y * d.y);
    return {d.x, d.y, z};
}
inline float CosineHemispherePDF(float c) { return c * InvPi; }

// ---- ThirdParty/pbrv4/spectrum.h ---------------------------------------------------------------
constexpr float Lambda_min = 360, Lambda_max = 830;
constexpr int NSpectrumSamples = 8;                 // spectrum.h:19
constexpr float CIE_Y_integral = 106.856895;        // spectrum.h:21

struct SampledSpectrum {                            // spectrum.h:52-249
    std::array<float, NSpectrumSamples> v{};
    SampledSpectrum() = default;
    explicit SampledSpectrum(float c) { v.fill(c); }
    float operator[](int i) const { return v[i]; }
    float& operator[](int i) { return v[i]; }
    SampledSpectrum& operator+=(const SampledSpectrum& s) { for (int i = 0; i < NSpectrumSamples; ++i) v[i] += s.v[i]; return *this; }
    SampledSpectrum& operator*=(const SampledSpectrum& s) { for (int i = 0; i < NSpectrumSamples; ++i) v[i] *= s.v[i]; return *this; }
    SampledSpectrum& operator*=(float a) { for (int i = 0; i < NSpectrumSamples; ++i) v[i] *= a; return *this; }
    SampledSpectrum operator+(const SampledSpectrum& s) const { SampledSpectrum r = *this; return r += s; }
    SampledSpectrum operator*(const SampledSpectrum& s) const { SampledSpectrum r = *this; return r *= s; }
    SampledSpectrum operator*(float a) const { SampledSpectrum r = *this; return r *= a; }
    float Average() const { float s = v[0]; for (int i = 1; i < NSpectrumSamples; ++i) s += v[i]; return s / NSpectrumSamples; }
    float MaxComponentValue() const { float m = v[0]; for (int i = 1; i < NSpectrumSamples; ++i) m = std::max(m, v[i]); return m; }
    bool IsBlack() const { for (float x : v) if (x != 0) return false; return true; }
};
inline SampledSpectrum operator*(float a, const SampledSpectrum& s) { return s * a; }
inline SampledSpectrum SafeDiv(const SampledSpectrum& a, const SampledSpectrum& b) {   // spectrum.h:643-649
    SampledSpectrum r;
    for (int i = 0; i < NSpectrumSamples; ++i) r[i] = (b[i] != 0) ? a[i] / b[i] : 0.f;
    return r;
}

struct SampledWavelengths {                         // spectrum.h:253-343
    std::array<float, NSpectrumSamples> lambda{}, pdf{};
    static SampledWavelengths SampleVisible(float u) {   // spectrum.h:322-336
        SampledWavelengths swl;
        for (int i = 0; i < NSpectrumSamples; ++i) {
            float up = u + float(i) / NSpectrumSamples;
            if (up > 1) up -= 1;
            swl.lambda[i] = SampleVisibleWavelengths(up);
            swl.pdf[i] = VisibleWavelengthsPDF(swl.lambda[i]);
        }
        return swl;
    }
    SampledSpectrum PDF() const { SampledSpectrum s; s.v = pdf; return s; }
    bool SecondaryTerminated() const { for (int i = 1; i < NSpectrumSamples; ++i) if (pdf[i] != 0) return false; return true; }
    void TerminateSecondary() {                     // spectrum.h:302-310
        if (SecondaryTerminated()) return;
        for (int i = 1; i < NSpectrumSamples; ++i) pdf[i] = 0;
        pdf[0] /= NSpectrumSamples;
    }
};

struct Spectrum {                                   // spectrum.h:38-49
    virtual ~Spectrum() = default;
    virtual float Query(float lambda) const = 0;
    virtual SampledSpectrum Sample(const SampledWavelengths& l) const {
        SampledSpectrum s;
        for (int i = 0; i < NSpectrumSamples; ++i) s[i] = Query(l.lambda[i]);
        return s;
    }
};
struct ConstantSpectrum : Spectrum {                // spectrum.h:355-373
    float c;
    explicit ConstantSpectrum(float c_) : c(c_) {}
    float Query(float) const override { return c; }
};
struct PiecewiseLinearSpectrum : Spectrum {         // spectrum.h:458-496; spectrum.cpp:60-72,134-165
    std::vector<float> lambdas, values;
    PiecewiseLinearSpectrum() = default;
    PiecewiseLinearSpectrum(const float* l, const float* v, int n) : lambdas(l, l + n), values(v, v + n) {}
    float Query(float lambda) const override {
        if (lambdas.empty() || lambda < lambdas.front() || lambda > lambdas.back()) return 0;
        int o = (int)FindInterval(lambdas.size(), [&](int i) { return lambdas[i] <= lambda; });
        float t = (lambda - lambdas[o]) / (lambdas[o + 1] - lambdas[o]);
        return Lerp(t, values[o], values[o + 1]);
    }
    void Scale(float s) { for (float& x : values) x *= s; }
    static PiecewiseLinearSpectrum* FromInterleaved(const float* samples, int count, bool normalize);
};
struct DenselySampledSpectrum : Spectrum {          // spectrum.h:376-456
    int lmin = 360, lmax = 830;
    std::vector<float> values;
    DenselySampledSpectrum() : values(471) {}
    explicit DenselySampledSpectrum(const Spectrum* s) : values(471) {
        if (s) for (int l = lmin; l <= lmax; ++l) values[l - lmin] = s->Query((float)l);
    }
    float Query(float lambda) const override {
        int off = (int)std::lround(lambda) - lmin;
        if (off < 0 || off >= (int)values.size()) return 0;
        return values[off];
    }
};
inline float InnerProduct(const Spectrum* f, const Spectrum* g) {      // spectrum.h:762-768
    float integral = 0;
    for (float l = Lambda_min; l <= Lambda_max; ++l) integral += f->Query(l) * g->Query(l);
    return integral;
}

// RGBSigmoidPolynomial: color.h:363-404 ; EvaluatePolynomial is Horner with FMA (helpers.h)
struct RGBSigmoidPolynomial {
    float c0 = 0, c1 = 0, c2 = 0;
    static float s(float x) {
        if (std::isinf(x)) return x > 0 ? 1 : 0;
        return .5f + x / (2 * std::sqrt(1 + (x * x)));
    }
    float operator()(float lambda) const { return s(std::fma(lambda, std::fma(lambda, c0, c1), c2)); }
};
// RGBToSpectrumTable::operator() uniform-rgb branch only (color.cpp:35-37).  The 64^3 coefficient
// file is not in the reference repo, so non-grey RGB cannot be restated.
inline bool GreyToSigmoid(float r, float g, float b, RGBSigmoidPolynomial* out) {
    if (!(r == g && g == b)) return false;
    *out = RGBSigmoidPolynomial{0, 0, (r - .5f) / std::sqrt(r * (1 - r))};
    return true;
}

struct XYZ { float X = 0, Y = 0, Z = 0; };
inline vec2 xy_of(XYZ c) { return {c.X / (c.X + c.Y + c.Z), c.Y / (c.X + c.Y + c.Z)}; }     // color.h:216
inline XYZ FromxyY(vec2 xy, float Y = 1) {                                                // color.h:219-224
    if (xy.y == 0) return {};
    return {xy.x * Y / xy.y, Y, (1 - xy.x - xy.y) * Y / xy.y};
}

// Global spectral tables (Spectra::Init, spectrum.cpp:2612-2640)
struct SpectraTables {
    std::unique_ptr<DenselySampledSpectrum> X, Y, Z;
    std::unique_ptr<PiecewiseLinearSpectrum> illumA, illumD50, illumD65, illumF1, illumF2, illumF11;
    static const SpectraTables& get();
};
XYZ SpectrumToXYZ(const Spectrum* s);              // spectrum.cpp:43-48

// RGBColorSpace sRGB (colorspace.cpp:13-28,82-100; colorspace.h:55-59)
struct RGBColorSpace {
    vec2 r, g, b, w;
    DenselySampledSpectrum illuminant;
    mat3 XYZFromRGB, RGBFromXYZ;
    RGBColorSpace(vec2 r_, vec2 g_, vec2 b_, const Spectrum* illum);
    vec3 ToRGB(vec3 xyz) const { return mul(RGBFromXYZ, xyz); }
    static const RGBColorSpace& sRGB();
};
// RGBAlbedoSpectrum / RGBIlluminantSpectrum for grey inputs (spectrum.h:535-559,593-638; spectrum.cpp:249-270)
struct RGBAlbedoSpectrum : Spectrum {
    RGBSigmoidPolynomial rsp;
    float Query(float l) const override { return rsp(l); }
};
struct RGBIlluminantSpectrum : Spectrum {
    float scale = 0; RGBSigmoidPolynomial rsp; const DenselySampledSpectrum* illuminant = nullptr;
    float Query(float l) const override { return illuminant ? scale * rsp(l) * illuminant->Query(l) : 0; }
    SampledSpectrum Sample(const SampledWavelengths& l) const override {
        if (!illuminant) return SampledSpectrum(0);
        SampledSpectrum s;
        for (int i = 0; i < NSpectrumSamples; ++i) s[i] = scale * rsp(l.lambda[i]);
        return s * illuminant->Sample(l);
    }
};
// RGBToSpectrumTable (color.h:405-432; operator() color.cpp:26-73).  The reference fills it from a data file that is not in
// its repository (color.cpp:107-166), so the oracle takes the table from the caller (orc_set_rgb_table), exactly as the
// reference's constructor does (zNodes, coeffs); without one only grey RGB (color.cpp:35-37) can be converted.
struct RGBToSpectrumTable {
    static constexpr int res = 64;
    std::vector<float> zNodes, coeffs;
    bool ready() const { return !coeffs.empty(); }
    RGBSigmoidPolynomial operator()(float r, float g, float b) const;
    static RGBToSpectrumTable& sRGB();
};
struct RGBUnboundedSpectrum : Spectrum {                                          // spectrum.h:561-590
    float scale = 1; RGBSigmoidPolynomial rsp;
    float Query(float l) const override { return scale * rsp(l); }
};
bool MakeRGBAlbedo(float r, float g, float b, RGBAlbedoSpectrum* out);
bool MakeRGBIlluminant(float r, float g, float b, RGBIlluminantSpectrum* out);
bool MakeRGBUnbounded(float r, float g, float b, RGBUnboundedSpectrum* out);

// Bradford white balance (color.h:600-629)
mat3 WhiteBalance(vec2 srcWhite, vec2 targetWhite);

// PixelSensor, XYZ constructor (pixelsensor.h:70-87)
struct PixelSensor {
    DenselySampledSpectrum r_bar, g_bar, b_bar;
    float imagingRatio;
    mat3 XYZFromSensorRGB;
    PixelSensor(const RGBColorSpace& out, const Spectrum* sensorIllum, float imagingRatio_);
    // measured sensor (pixelsensor.h:37-68): response curves r, g, b; XYZFromSensorRGB by least squares over the 24 Macbeth swatches
    PixelSensor(const Spectrum* r, const Spectrum* g, const Spectrum* b, const RGBColorSpace& out, const Spectrum* sensorIllum, float imagingRatio_);
    vec3 ToSensorRGB(SampledSpectrum L, const SampledWavelengths& lambda) const {
        L = SafeDiv(L, lambda.PDF());
        float r = (r_bar.Sample(lambda) * L).Average();
        float g = (g_bar.Sample(lambda) * L).Average();
        float b = (b_bar.Sample(lambda) * L).Average();
        return {imagingRatio * r, imagingRatio * g, imagingRatio * b};
    }
};

// Filters (filters.h:23-93,267-296)
struct FilterSample { vec2 p; float weight; };
struct Filter {
    virtual ~Filter() = default;
    virtual FilterSample Sample(vec2 u) const = 0;
};
struct BoxFilter : Filter {
    vec2 radius;
    explicit BoxFilter(vec2 r = vec2(0.5f, 0.5f)) : radius(r) {}
    FilterSample Sample(vec2 u) const override {
        return {vec2(Lerp(u.x, -radius.x, radius.x), Lerp(u.y, -radius.y, radius.y)), 1.0f};
    }
};
struct TriangleFilter : Filter {
    vec2 radius;
    explicit TriangleFilter(vec2 r) : radius(r) {}
    FilterSample Sample(vec2 u) const override {
        return {vec2(SampleTent(u.x, radius.x), SampleTent(u.y, radius.y)), 1.0f};
    }
};

// Gaussian (helpers.h:221-225)
inline float Gaussian(float x, float mu = 0, float sigma = 1) {
    return 1.0f / std::sqrt(2 * Pi * sigma * sigma) * std::exp(-std::pow(x - mu, 2.0f) / (2 * sigma * sigma));
}
// Continuous_Inversion_Sampler (RayTracer/Sampling.h:781-895): tabulated CDF (Riemann sums at the right edges, renormalised),
// binary search with the reference's float midpoint, linear interpolation inside the cell
struct ContinuousInversionSampler {
    int N; float a, b;
    std::vector<float> cdf;
    template <typename F>
    ContinuousInversionSampler(F pdf, float a_, float b_, float precision_N) : N((int)precision_N), a(a_), b(b_) {    // :784-806
        cdf.resize(N + 1);
        float delta_x = (b - a) / (float)N;
        float sum = 0;
        cdf[0] = 0.0f;
        for (int n = 1; n < N + 1; n++) {
            float current_x = std::clamp(a + delta_x * n, a, b);
            sum += delta_x * pdf(current_x);
            cdf[n] = sum;
        }
        float scaling_term = 1.0f / cdf[N];
        for (int n = 1; n < N; n++) cdf[n] *= scaling_term;
        cdf[N] = 1.0f;
    }
    float Sample(float U) const {                                                                                     // :809-848, use_U
        int index = -1, low = 0, high = N;
        while (low <= high) {
            int mid = low + (high - low) / 2.0f;
            // cdf[mid + 1] with mid == N reads past the table in the reference; it is only reached for U > 1
            if (mid < N && cdf[mid] < U && U <= cdf[mid + 1]) { index = mid; break; }
            if (cdf[mid] < U) low = mid + 1;
            else high = mid - 1;
        }
        if (index == -1) return 0;                                       // "couldn't find index" (U == 0)
        float t = std::clamp((U - cdf[index]) / (cdf[index + 1] - cdf[index]), 0.0f, 1.f);
        float delta_x = (b - a) / (float)N;
        return (a + delta_x * index) + t * (delta_x * (index + 1) - delta_x * index);
    }
};
// GaussianFilter (filters.h:96-163)
struct GaussianFilter : Filter {
    vec2 radius; float sigma, expX, expY;
    ContinuousInversionSampler inv_sampler_x, inv_sampler_y;
    GaussianFilter(vec2 r, float sigma_ = 0.5f)
        : radius(r), sigma(sigma_), expX(Gaussian(r.x, 0, sigma_)), expY(Gaussian(r.y, 0, sigma_)),
          inv_sampler_x([=, this](float x) { return EvaluateX(x); }, -r.x, r.x, 10000),
          inv_sampler_y([=, this](float y) { return EvaluateY(y); }, -r.y, r.y, 10000) {}
    float EvaluateX(float x) const { return std::max<float>(0, Gaussian(x, 0, sigma) - expX); }
    float EvaluateY(float y) const { return std::max<float>(0, Gaussian(y, 0, sigma) - expY); }
    float Evaluate(vec2 p) const { return EvaluateX(p.x) * EvaluateY(p.y); }
    FilterSample Sample(vec2 u) const override {                                                                      // :129-135
        FilterSample fs;
        fs.p.x = inv_sampler_x.Sample(u.x);
        fs.p.y = inv_sampler_y.Sample(u.y);
        fs.weight = Evaluate(fs.p) / (EvaluateX(fs.p.x) * EvaluateY(fs.p.y));
        return fs;
    }
};

}  // namespace orc
