// crt_spectrum.cuh -- spectral evaluation on the device (ThirdParty/pbrv4/spectrum.h, color.h, pixelsensor.h).
#pragma once
#include "crt_device_scene.h"

namespace crt {

struct Spec8 { float v[CRT_NLAMBDA]; };

// SampleVisibleWavelengths / VisibleWavelengthsPDF (RayTracer/Sampling.h:63-71).  The reference calls float
// atanh/cosh from the host libm; the device evaluates them in double and rounds once, which is the correctly
// rounded float result in all but ~1e-8 of cases (DESIGN.md "floating-point tolerance").
__device__ __noinline__ float sample_visible_wavelength(float u) {
    float x = 0.85691062f - 1.82750197f * u;
    float at = (float)atanh((double)x);
    return 538 - 138.888889f * at;
}
__device__ __noinline__ float visible_wavelength_pdf(float lambda) {
    if (lambda < 360 || lambda > 830) return 0;
    float ch = (float)cosh((double)(0.0072f * (lambda - 538)));
    // std::pow(float, int) is evaluated in double; the division too (C++ promotion), then narrowed
    return (float)((double)0.0039398042f / ((double)ch * (double)ch));
}
// SampledWavelengths::SampleVisible (spectrum.h:322-336)
CRT_D void sample_visible(float u, Spec8& lambda, Spec8& pdf) {
#pragma unroll
    for (int i = 0; i < CRT_NLAMBDA; ++i) {
        float up = u + float(i) / CRT_NLAMBDA;
        if (up > 1) up -= 1;
        lambda.v[i] = sample_visible_wavelength(up);
        pdf.v[i] = visible_wavelength_pdf(lambda.v[i]);
    }
}
// DenselySampledSpectrum::Sample (spectrum.h:386-398): value at lround(lambda) - 360
CRT_D float dense_lookup(const float* table, float lambda) {
    int off = (int)lroundf(lambda) - 360;
    if (off < 0 || off >= 471) return 0;
    return __ldg(&table[off]);
}
// PiecewiseLinearSpectrum::Query (spectrum.cpp:60-72) with FindInterval (helpers.h:160-172)
CRT_D float piecewise_query(const float* lambdas, const float* values, int n, float lambda) {
    if (n == 0 || lambda < __ldg(&lambdas[0]) || lambda > __ldg(&lambdas[n - 1])) return 0;
    int size = n - 2, first = 1;
    while (size > 0) {
        int half = size >> 1, middle = first + half;
        bool r = __ldg(&lambdas[middle]) <= lambda;
        first = r ? middle + 1 : first;
        size = r ? size - (half + 1) : half;
    }
    int o = first - 1;
    o = o < 0 ? 0 : (o > n - 2 ? n - 2 : o);
    float l0 = __ldg(&lambdas[o]), l1 = __ldg(&lambdas[o + 1]);
    float t = (lambda - l0) / (l1 - l0);
    return lerp_pbrt(t, __ldg(&values[o]), __ldg(&values[o + 1]));
}
// The same query with the interval found through the per-nanometre index table of DevSpectrum (crt_device_scene.h): o = the largest i in
// [0, n-2] with lambdas[i] <= lambda, which is exactly what FindInterval returns for lambda inside [lambdas[0], lambdas[n-1]].
CRT_D float piecewise_query_lut(const float* lambdas, const float* values, int n, const float* lut, int lut_base, int lut_n, float lambda) {
    if (n == 0 || lambda < __ldg(&lambdas[0]) || lambda > __ldg(&lambdas[n - 1])) return 0;
    int b = (int)floorf(lambda) - lut_base;
    b = b < 0 ? 0 : (b > lut_n - 1 ? lut_n - 1 : b);
    int o = __float_as_int(__ldg(&lut[b]));
    while (o + 1 < n && __ldg(&lambdas[o + 1]) <= lambda) ++o;
    o = o > n - 2 ? n - 2 : o;
    float l0 = __ldg(&lambdas[o]), l1 = __ldg(&lambdas[o + 1]);
    float t = (lambda - l0) / (l1 - l0);
    return lerp_pbrt(t, __ldg(&values[o]), __ldg(&values[o + 1]));
}
// RGBSigmoidPolynomial::operator() (color.h:376-399); EvaluatePolynomial is an FMA Horner chain (helpers.h:117-126)
CRT_D float sigmoid_eval(float c0, float c1, float c2, float lambda) {
    float x = fmaf(lambda, fmaf(lambda, c0, c1), c2);
    if (isinf(x)) return x > 0 ? 1.f : 0.f;
    return .5f + x / (2 * sqrtf(1 + (x * x)));
}
// Spectrum::operator()(lambda) for any spectrum of the scene.  Deliberately NOT inlined: a bounce of the path integrator evaluates
// spectra at ~40 places (8 wavelengths x emission, reflectance, light, eta, k); inlined, those copies made k_path_shade about three times
// as large and the kernel stalled on instruction fetch (ncu: "no instruction" = 68 % of its stall samples).  One copy, scalar arguments in
// registers -- the scene's pointers are passed by value so that the kernel-parameter struct is never copied to local memory.
__device__ __noinline__ float spectrum_query_nl(const DevSpectrum* spectra, const float* pool, const float* d65dense, int id, float lambda) {
    const DevSpectrum sp = spectra[id];
    switch (sp.kind) {
        case SPEC_CONSTANT: return sp.c0;
        case SPEC_PIECEWISE:
            if (sp.lut_n > 0) return piecewise_query_lut(pool + sp.offset, pool + sp.offset + sp.n, sp.n, pool + sp.offset + 2 * sp.n, sp.lut_base, sp.lut_n, lambda);
            return piecewise_query(pool + sp.offset, pool + sp.offset + sp.n, sp.n, lambda);
        case SPEC_DENSE: return dense_lookup(pool + sp.offset, lambda);
        case SPEC_SIGMOID: return sigmoid_eval(sp.c0, sp.c1, sp.c2, lambda);
        case SPEC_SIGMOID_UNBOUNDED: return sp.scale * sigmoid_eval(sp.c0, sp.c1, sp.c2, lambda);      // RGBUnboundedSpectrum::Query, spectrum.h:566
        default: return (sp.scale * sigmoid_eval(sp.c0, sp.c1, sp.c2, lambda)) * dense_lookup(d65dense, lambda);   // RGBIlluminantSpectrum::Sample
    }
}
CRT_D float spectrum_query(const DeviceScene& S, int id, float lambda) { return spectrum_query_nl(S.spectra, S.pool, S.d65dense, id, lambda); }
CRT_D void spectrum_sample(const DeviceScene& S, int id, const Spec8& lambda, Spec8& out) {
#pragma unroll
    for (int i = 0; i < CRT_NLAMBDA; ++i) out.v[i] = spectrum_query(S, id, lambda.v[i]);
}
// PixelSensor::ToSensorRGB with the XYZ sensor (pixelsensor.h:81-87): SafeDiv by the pdf, product with the
// sensor's response curves (the CIE observer for the XYZ sensor), Average (sequential sum / 8), times imagingRatio
CRT_D f3 to_sensor_rgb(const DeviceScene& S, const Spec8& L, const Spec8& lambda, const Spec8& pdf) {
    const float imagingRatio = S.imaging_ratio;
    float sx = 0, sy = 0, sz = 0;
#pragma unroll
    for (int i = 0; i < CRT_NLAMBDA; ++i) {
        float l = (pdf.v[i] != 0) ? L.v[i] / pdf.v[i] : 0.f;
        float x = dense_lookup(S.cieX, lambda.v[i]) * l;
        float y = dense_lookup(S.cieY, lambda.v[i]) * l;
        float z = dense_lookup(S.cieZ, lambda.v[i]) * l;
        if (i == 0) { sx = x; sy = y; sz = z; } else { sx += x; sy += y; sz += z; }
    }
    return mk3(imagingRatio * (sx / CRT_NLAMBDA), imagingRatio * (sy / CRT_NLAMBDA), imagingRatio * (sz / CRT_NLAMBDA));
}

}  // namespace crt
