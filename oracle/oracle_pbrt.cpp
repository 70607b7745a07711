// ORACLE (test infrastructure, never shipped, never on the product path).
// Out-of-line parts of oracle_pbrt.h.  See that header for scope and parity status.
#include "oracle_pbrt.h"

#include "../computational_ray_tracer_b200/data/spectral_tables.inc"

namespace orc {
const float* swatch_table(int i, int* n);

// spectrum.cpp:134-165
PiecewiseLinearSpectrum* PiecewiseLinearSpectrum::FromInterleaved(const float* samples, int count, bool normalize) {
    int n = count / 2;
    std::vector<float> lambda, v;
    if (samples[0] > Lambda_min) {
        lambda.push_back(Lambda_min - 1);
        v.push_back(samples[1]);
    }
    for (int i = 0; i < n; ++i) {
        lambda.push_back(samples[2 * i]);
        v.push_back(samples[2 * i + 1]);
    }
    if (lambda.back() < Lambda_max) {
        lambda.push_back(Lambda_max + 1);
        v.push_back(v.back());
    }
    auto* spec = new PiecewiseLinearSpectrum(lambda.data(), v.data(), (int)lambda.size());
    if (normalize) spec->Scale(CIE_Y_integral / InnerProduct(spec, SpectraTables::get().Y.get()));
    return spec;
}

// spectrum.cpp:2612-2640 (only the tables the render path can reach)
const SpectraTables& SpectraTables::get() {
    static SpectraTables* t = [] {
        auto* s = new SpectraTables;
        std::vector<float> lam(471);
        for (int i = 0; i < 471; ++i) lam[i] = float(360 + i);
        PiecewiseLinearSpectrum x(lam.data(), crt_tab_cie_x, 471), y(lam.data(), crt_tab_cie_y, 471), z(lam.data(), crt_tab_cie_z, 471);
        s->X = std::make_unique<DenselySampledSpectrum>(&x);
        s->Y = std::make_unique<DenselySampledSpectrum>(&y);
        s->Z = std::make_unique<DenselySampledSpectrum>(&z);
        return s;
    }();
    static bool illumInit = false;
    if (!illumInit) {
        illumInit = true;   // FromInterleaved(normalize) re-enters get() for Y, which is ready by now
        t->illumA.reset(PiecewiseLinearSpectrum::FromInterleaved(crt_tab_illum_a, crt_tab_illum_a_n, true));
        t->illumD50.reset(PiecewiseLinearSpectrum::FromInterleaved(crt_tab_illum_d50, crt_tab_illum_d50_n, true));
        t->illumD65.reset(PiecewiseLinearSpectrum::FromInterleaved(crt_tab_illum_d65, crt_tab_illum_d65_n, true));
        t->illumF1.reset(PiecewiseLinearSpectrum::FromInterleaved(crt_tab_illum_f1, crt_tab_illum_f1_n, true));
        t->illumF2.reset(PiecewiseLinearSpectrum::FromInterleaved(crt_tab_illum_f2, crt_tab_illum_f2_n, true));
        t->illumF11.reset(PiecewiseLinearSpectrum::FromInterleaved(crt_tab_illum_f11, crt_tab_illum_f11_n, true));
    }
    return *t;
}

XYZ SpectrumToXYZ(const Spectrum* s) {
    const auto& T = SpectraTables::get();
    XYZ r{InnerProduct(T.X.get(), s), InnerProduct(T.Y.get(), s), InnerProduct(T.Z.get(), s)};
    r.X /= CIE_Y_integral; r.Y /= CIE_Y_integral; r.Z /= CIE_Y_integral;
    return r;
}

// colorspace.cpp:13-28
RGBColorSpace::RGBColorSpace(vec2 r_, vec2 g_, vec2 b_, const Spectrum* illum) : r(r_), g(g_), b(b_), illuminant(illum) {
    XYZ W = SpectrumToXYZ(illum);
    w = xy_of(W);
    XYZ R = FromxyY(r), G = FromxyY(g), B = FromxyY(b);
    mat3 rgb;
    rgb.setcol(0, {R.X, R.Y, R.Z}); rgb.setcol(1, {G.X, G.Y, G.Z}); rgb.setcol(2, {B.X, B.Y, B.Z});
    vec3 C = mul(inverse(rgb), vec3(W.X, W.Y, W.Z));
    mat3 diag;
    diag.c[0][0] = C.x; diag.c[1][1] = C.y; diag.c[2][2] = C.z;
    XYZFromRGB = mul(rgb, diag);
    RGBFromXYZ = inverse(XYZFromRGB);
}
const RGBColorSpace& RGBColorSpace::sRGB() {   // colorspace.cpp:82-100
    static RGBColorSpace* cs = new RGBColorSpace(vec2(.64, .33), vec2(.3, .6), vec2(.15, .06), SpectraTables::get().illumD65.get());
    return *cs;
}

RGBToSpectrumTable& RGBToSpectrumTable::sRGB() { static RGBToSpectrumTable t; return t; }
RGBSigmoidPolynomial RGBToSpectrumTable::operator()(float r, float g, float b) const {      // color.cpp:26-73
    const float rgb[3] = {r, g, b};
    if (rgb[0] == rgb[1] && rgb[1] == rgb[2])                                                // :35-37
        return RGBSigmoidPolynomial{0, 0, (rgb[0] - .5f) / std::sqrt(rgb[0] * (1 - rgb[0]))};
    int maxc = (rgb[0] > rgb[1]) ? ((rgb[0] > rgb[2]) ? 0 : 2) : ((rgb[1] > rgb[2]) ? 1 : 2);  // :40-41
    float z = rgb[maxc];
    float x = rgb[(maxc + 1) % 3] * (res - 1) / z;
    float y = rgb[(maxc + 2) % 3] * (res - 1) / z;
    int xi = std::min((int)x, res - 2), yi = std::min((int)y, res - 2);                        // :47
    int zi = (int)FindInterval((size_t)res, [&](size_t i) { return zNodes[i] < z; });          // :48
    float dx = x - xi, dy = y - yi, dz = (z - zNodes[zi]) / (zNodes[zi + 1] - zNodes[zi]);
    float c[3];
    for (int i = 0; i < 3; ++i) {
        auto co = [&](int ddx, int ddy, int ddz) {                                             // :56-62
            size_t index = (size_t)maxc * 64 * 64 * 64 * 3 + (size_t)(zi + ddz) * 64 * 64 * 3 + (size_t)(yi + ddy) * 64 * 3 + (size_t)(xi + ddx) * 3 + i;
            return coeffs[index];
        };
        c[i] = Lerp(dz, Lerp(dy, Lerp(dx, co(0, 0, 0), co(1, 0, 0)), Lerp(dx, co(0, 1, 0), co(1, 1, 0))),
                    Lerp(dy, Lerp(dx, co(0, 0, 1), co(1, 0, 1)), Lerp(dx, co(0, 1, 1), co(1, 1, 1))));
    }
    return RGBSigmoidPolynomial{c[0], c[1], c[2]};
}
// RGBColorSpace::ToRGBCoeffs (colorspace.cpp:38-43): ClampZero, then the table; false if a non-grey RGB meets no table
static bool ToRGBCoeffs(float r, float g, float b, RGBSigmoidPolynomial* out) {
    r = std::max(0.0f, r); g = std::max(0.0f, g); b = std::max(0.0f, b);
    if (GreyToSigmoid(r, g, b, out)) return true;
    if (!RGBToSpectrumTable::sRGB().ready()) return false;
    *out = RGBToSpectrumTable::sRGB()(r, g, b);
    return true;
}
bool MakeRGBAlbedo(float r, float g, float b, RGBAlbedoSpectrum* out) {        // spectrum.cpp:249-254
    return ToRGBCoeffs(r, g, b, &out->rsp);
}
static bool scaled_coeffs(float r, float g, float b, float* scale, RGBSigmoidPolynomial* rsp) {   // spectrum.cpp:256-270
    float m = std::max(r, g);
    m = std::max(m, b);
    *scale = 2 * m;
    float s = *scale;
    return ToRGBCoeffs(s ? r / s : 0, s ? g / s : 0, s ? b / s : 0, rsp);
}
bool MakeRGBIlluminant(float r, float g, float b, RGBIlluminantSpectrum* out) { // spectrum.cpp:264-270
    out->illuminant = &RGBColorSpace::sRGB().illuminant;
    return scaled_coeffs(r, g, b, &out->scale, &out->rsp);
}
bool MakeRGBUnbounded(float r, float g, float b, RGBUnboundedSpectrum* out) {   // spectrum.cpp:256-262
    return scaled_coeffs(r, g, b, &out->scale, &out->rsp);
}

mat3 WhiteBalance(vec2 srcWhite, vec2 targetWhite) {   // color.h:600-629
    mat3 LMSFromXYZ, XYZFromLMS;
    LMSFromXYZ.setcol(0, {0.8951, -0.7502, 0.0389});
    LMSFromXYZ.setcol(1, {0.2664, 1.7135, -0.0685});
    LMSFromXYZ.setcol(2, {-0.1614, 0.0367, 1.0296});
    XYZFromLMS.setcol(0, {0.986993, 0.432305, -0.00852866});
    XYZFromLMS.setcol(1, {-0.147054, 0.51836, 0.0400428});
    XYZFromLMS.setcol(2, {0.159963, 0.0492912, 0.968487});
    XYZ s = FromxyY(srcWhite), d = FromxyY(targetWhite);
    vec3 srcLMS = mul(LMSFromXYZ, vec3(s.X, s.Y, s.Z)), dstLMS = mul(LMSFromXYZ, vec3(d.X, d.Y, d.Z));
    mat3 corr;
    corr.c[0][0] = dstLMS.x / srcLMS.x; corr.c[1][1] = dstLMS.y / srcLMS.y; corr.c[2][2] = dstLMS.z / srcLMS.z;
    return mul(mul(XYZFromLMS, corr), LMSFromXYZ);
}

PixelSensor::PixelSensor(const RGBColorSpace& out, const Spectrum* sensorIllum, float imagingRatio_)   // pixelsensor.h:70-79
    : r_bar(*SpectraTables::get().X), g_bar(*SpectraTables::get().Y), b_bar(*SpectraTables::get().Z), imagingRatio(imagingRatio_) {
    if (sensorIllum) {
        vec2 sourceWhite = xy_of(SpectrumToXYZ(sensorIllum));
        XYZFromSensorRGB = WhiteBalance(sourceWhite, out.w);
    }
}

// PixelSensor::ProjectReflectance (pixelsensor.h:104-117)
static void ProjectReflectance(const Spectrum* refl, const Spectrum* illum, const Spectrum* b1, const Spectrum* b2, const Spectrum* b3, float* out3) {
    float result[3] = {0, 0, 0};
    float g_integral = 0;
    for (float lambda = Lambda_min; lambda <= Lambda_max; ++lambda) {
        g_integral += b2->Query(lambda) * illum->Query(lambda);
        result[0] += b1->Query(lambda) * refl->Query(lambda) * illum->Query(lambda);
        result[1] += b2->Query(lambda) * refl->Query(lambda) * illum->Query(lambda);
        result[2] += b3->Query(lambda) * refl->Query(lambda) * illum->Query(lambda);
    }
    for (int c = 0; c < 3; ++c) out3[c] = result[c] / g_integral;
}
// LinearLeastSquares<3> (helpers.h:257-274)
static mat3 LinearLeastSquares3(const float A[][3], const float B[][3], int rows) {
    mat3 AtA, AtB;
    for (int i = 0; i < 3; ++i) for (int j = 0; j < 3; ++j) { AtA.c[i][j] = 0; AtB.c[i][j] = 0; }
    for (int i = 0; i < 3; ++i)
        for (int j = 0; j < 3; ++j)
            for (int r = 0; r < rows; ++r) {
                AtA.c[i][j] += A[r][i] * A[r][j];
                AtB.c[i][j] += A[r][i] * B[r][j];
            }
    mat3 AtAi = inverse(AtA);
    return transpose(mul(AtAi, AtB));
}
PixelSensor::PixelSensor(const Spectrum* r, const Spectrum* g, const Spectrum* b, const RGBColorSpace& out, const Spectrum* sensorIllum,
                         float imagingRatio_)                                                   // pixelsensor.h:37-68
    : r_bar(r), g_bar(g), b_bar(b), imagingRatio(imagingRatio_) {
    const auto& T = SpectraTables::get();
    constexpr int nSwatch = 24;
    std::vector<std::unique_ptr<PiecewiseLinearSpectrum>> swatches;
    for (int i = 0; i < nSwatch; ++i) { int n; const float* t = swatch_table(i, &n); swatches.emplace_back(PiecewiseLinearSpectrum::FromInterleaved(t, n, false)); }
    float rgbCamera[nSwatch][3];
    for (int i = 0; i < nSwatch; ++i) ProjectReflectance(swatches[i].get(), sensorIllum, &r_bar, &g_bar, &b_bar, rgbCamera[i]);
    float xyzOutput[nSwatch][3];
    float sensorWhiteG = InnerProduct(sensorIllum, &g_bar);
    float sensorWhiteY = InnerProduct(sensorIllum, T.Y.get());
    for (int i = 0; i < nSwatch; ++i) {
        float xyz[3];
        ProjectReflectance(swatches[i].get(), &out.illuminant, T.X.get(), T.Y.get(), T.Z.get(), xyz);
        float k = sensorWhiteY / sensorWhiteG;
        for (int c = 0; c < 3; ++c) xyzOutput[i][c] = k * xyz[c];                               // XYZ::operator*(float): a * X
    }
    XYZFromSensorRGB = LinearLeastSquares3(rgbCamera, xyzOutput, nSwatch);
}

const float* swatch_table(int i, int* n) {
    *n = crt_tab_swatch_offsets[i + 1] - crt_tab_swatch_offsets[i];
    return crt_tab_swatches + crt_tab_swatch_offsets[i];
}
const float* named_table(const char* name, int* n) {
    struct E { const char* k; const float* p; int n; };
    static const E tabs[] = {
        {"illum_a", crt_tab_illum_a, crt_tab_illum_a_n}, {"illum_d50", crt_tab_illum_d50, crt_tab_illum_d50_n},
        {"illum_d65", crt_tab_illum_d65, crt_tab_illum_d65_n}, {"illum_f1", crt_tab_illum_f1, crt_tab_illum_f1_n},
        {"illum_f2", crt_tab_illum_f2, crt_tab_illum_f2_n}, {"illum_f11", crt_tab_illum_f11, crt_tab_illum_f11_n},
        {"ag_eta", crt_tab_ag_eta, crt_tab_ag_eta_n}, {"ag_k", crt_tab_ag_k, crt_tab_ag_k_n},
        {"al_eta", crt_tab_al_eta, crt_tab_al_eta_n}, {"al_k", crt_tab_al_k, crt_tab_al_k_n},
        {"au_eta", crt_tab_au_eta, crt_tab_au_eta_n}, {"au_k", crt_tab_au_k, crt_tab_au_k_n},
        {"cu_eta", crt_tab_cu_eta, crt_tab_cu_eta_n}, {"cu_k", crt_tab_cu_k, crt_tab_cu_k_n},
        {"cuzn_eta", crt_tab_cuzn_eta, crt_tab_cuzn_eta_n}, {"cuzn_k", crt_tab_cuzn_k, crt_tab_cuzn_k_n},
        {"glass_bk7", crt_tab_glass_bk7, crt_tab_glass_bk7_n}, {"glass_baf10", crt_tab_glass_baf10, crt_tab_glass_baf10_n},
        {"glass_fk51a", crt_tab_glass_fk51a, crt_tab_glass_fk51a_n}, {"glass_lasf9", crt_tab_glass_lasf9, crt_tab_glass_lasf9_n},
        {"glass_sf5", crt_tab_glass_sf5, crt_tab_glass_sf5_n}, {"glass_sf10", crt_tab_glass_sf10, crt_tab_glass_sf10_n},
        {"glass_sf11", crt_tab_glass_sf11, crt_tab_glass_sf11_n},
    };
    for (const E& e : tabs)
        if (std::string(e.k) == name) { *n = e.n; return e.p; }
    *n = 0;
    return nullptr;
}

}  // namespace orc
