// crt_spectra.cpp -- host-side spectral tables and colour constants uploaded to the device at scene commit.
// Follows ThirdParty/pbrv4/spectrum.cpp:60-72,134-165,2612-2640 (tables), colorspace.cpp:13-28,82-100,
// color.h:600-629 (Bradford white balance) and pixelsensor.h:70-79 (XYZ sensor).  All arithmetic is fp32
// in the reference's order so the constants match a CPU build of the reference bit for bit.
#include <cstring>
#include <mutex>

#include "crt_host.h"
#include "../data/spectral_tables.inc"

namespace crt {

static inline size_t find_interval_le(const std::vector<float>& xs, float x) {     // helpers.h:160-172 with pred xs[i] <= x
    long size = (long)xs.size() - 2, first = 1;
    while (size > 0) {
        long half = size >> 1, middle = first + half;
        bool r = xs[middle] <= x;
        first = r ? middle + 1 : first;
        size = r ? size - (half + 1) : half;
    }
    long v = first - 1, hi = (long)xs.size() - 2;
    return (size_t)(v < 0 ? 0 : (v > hi ? hi : v));
}
float PiecewiseLinear::query(float lambda) const {                               // spectrum.cpp:60-72
    if (lambdas.empty() || lambda < lambdas.front() || lambda > lambdas.back()) return 0;
    size_t o = find_interval_le(lambdas, lambda);
    float t = (lambda - lambdas[o]) / (lambdas[o + 1] - lambdas[o]);
    return lerp_pbrt(t, values[o], values[o + 1]);
}
static float inner_product_with_Y(const PiecewiseLinear& s) {                     // spectrum.h:762-768 (f = s, g = Y)
    float integral = 0;
    for (float l = 360; l <= 830; ++l) integral += s.query(l) * crt_tab_cie_y[(int)l - 360];
    return integral;
}
PiecewiseLinear PiecewiseLinear::from_interleaved(const float* s, int count, bool normalize) {   // spectrum.cpp:134-165
    PiecewiseLinear p;
    int n = count / 2;
    if (s[0] > 360.0f) { p.lambdas.push_back(360.0f - 1); p.values.push_back(s[1]); }
    for (int i = 0; i < n; ++i) { p.lambdas.push_back(s[2 * i]); p.values.push_back(s[2 * i + 1]); }
    if (p.lambdas.back() < 830.0f) { p.lambdas.push_back(830.0f + 1); p.values.push_back(p.values.back()); }
    if (normalize) {
        float scale = 106.856895f / inner_product_with_Y(p);      // CIE_Y_integral, spectrum.h:21
        for (float& v : p.values) v *= scale;
    }
    return p;
}

static void xyz_of(const PiecewiseLinear& s, float* xyz) {                        // SpectrumToXYZ, spectrum.cpp:43-48
    float X = 0, Y = 0, Z = 0;
    // three separate InnerProduct(matching curve, s) loops; each accumulates f*g in wavelength order
    for (float l = 360; l <= 830; ++l) X += crt_tab_cie_x[(int)l - 360] * s.query(l);
    for (float l = 360; l <= 830; ++l) Y += crt_tab_cie_y[(int)l - 360] * s.query(l);
    for (float l = 360; l <= 830; ++l) Z += crt_tab_cie_z[(int)l - 360] * s.query(l);
    xyz[0] = X / 106.856895f; xyz[1] = Y / 106.856895f; xyz[2] = Z / 106.856895f;
}
static void from_xyY(float x, float y, float* out) {                              // XYZ::FromxyY, color.h:219-224
    const float Y = 1;
    if (y == 0) { out[0] = out[1] = out[2] = 0; return; }
    out[0] = x * Y / y; out[1] = Y; out[2] = (1 - x - y) * Y / y;
}
static void m3_mul_v3(const float* m, const float* v, float* out) {
    f3 r = mul_m3_v3(m, mk3(v[0], v[1], v[2]));
    out[0] = r.x; out[1] = r.y; out[2] = r.z;
}

// PixelSensor::ProjectReflectance (pixelsensor.h:104-117) with every spectrum but the swatch given at the integer wavelengths
static void project_reflectance(const PiecewiseLinear& refl, const float* illum, const float* b1, const float* b2, const float* b3, float* out3) {
    float result[3] = {0, 0, 0};
    float g_integral = 0;
    for (int i = 0; i < 471; ++i) {
        const float q = refl.query((float)(360 + i));
        g_integral += b2[i] * illum[i];
        result[0] += b1[i] * q * illum[i];
        result[1] += b2[i] * q * illum[i];
        result[2] += b3[i] * q * illum[i];
    }
    for (int c = 0; c < 3; ++c) out3[c] = result[c] / g_integral;
}
void measured_sensor_matrix(const float* r, const float* g, const float* b, const float* illum, float* out9) {   // pixelsensor.h:37-68
    const HostSpectra& h = host_spectra();
    const int nSwatch = 24;
    float rgbCamera[24][3], xyzOutput[24][3];
    float sensorWhiteG = 0, sensorWhiteY = 0;
    for (int i = 0; i < 471; ++i) sensorWhiteG += illum[i] * g[i];           // InnerProduct(sensorIllum, &g_bar)
    for (int i = 0; i < 471; ++i) sensorWhiteY += illum[i] * h.Y[i];         // InnerProduct(sensorIllum, &Spectra::Y())
    for (int s = 0; s < nSwatch; ++s) {
        int n = 0;
        const float* t = swatch_table(s, &n);
        PiecewiseLinear sw = PiecewiseLinear::from_interleaved(t, n, false);
        project_reflectance(sw, illum, r, g, b, rgbCamera[s]);
        float xyz[3];
        project_reflectance(sw, h.D65dense, h.X, h.Y, h.Z, xyz);
        const float k = sensorWhiteY / sensorWhiteG;
        for (int c = 0; c < 3; ++c) xyzOutput[s][c] = k * xyz[c];
    }
    // LinearLeastSquares (helpers.h:257-274), glm's [col][row] indexing kept as written there: m[3*col+row]
    float AtA[9] = {0, 0, 0, 0, 0, 0, 0, 0, 0}, AtB[9] = {0, 0, 0, 0, 0, 0, 0, 0, 0};
    for (int i = 0; i < 3; ++i)
        for (int j = 0; j < 3; ++j)
            for (int rr = 0; rr < nSwatch; ++rr) {
                AtA[3 * i + j] += rgbCamera[rr][i] * rgbCamera[rr][j];
                AtB[3 * i + j] += rgbCamera[rr][i] * xyzOutput[rr][j];
            }
    float AtAi[9], prod[9];
    m3_inverse(AtA, AtAi);
    m3_mul(AtAi, AtB, prod);
    for (int c = 0; c < 3; ++c) for (int rr = 0; rr < 3; ++rr) out9[3 * c + rr] = prod[3 * rr + c];      // glm::transpose
}

const HostSpectra& host_spectra() {
    static HostSpectra* hs = nullptr;
    static std::once_flag once;
    std::call_once(once, [] {
        auto* h = new HostSpectra;
        // DenselySampledSpectrum(PiecewiseLinear over the integer knots 360..830) reproduces the table values:
        // query at a knot is Lerp(0, v[o], v[o+1]) = (1-0)*v[o] + 0*v[o+1].
        for (int i = 0; i < 471; ++i) {
            h->X[i] = (1 - 0.0f) * crt_tab_cie_x[i] + 0.0f * crt_tab_cie_x[i < 470 ? i + 1 : i];
            h->Y[i] = (1 - 0.0f) * crt_tab_cie_y[i] + 0.0f * crt_tab_cie_y[i < 470 ? i + 1 : i];
            h->Z[i] = (1 - 0.0f) * crt_tab_cie_z[i] + 0.0f * crt_tab_cie_z[i < 470 ? i + 1 : i];
        }
        h->illum[0] = PiecewiseLinear::from_interleaved(crt_tab_illum_a, crt_tab_illum_a_n, true);
        h->illum[1] = PiecewiseLinear::from_interleaved(crt_tab_illum_d50, crt_tab_illum_d50_n, true);
        h->illum[2] = PiecewiseLinear::from_interleaved(crt_tab_illum_d65, crt_tab_illum_d65_n, true);
        h->illum[3] = PiecewiseLinear::from_interleaved(crt_tab_illum_f1, crt_tab_illum_f1_n, true);
        h->illum[4] = PiecewiseLinear::from_interleaved(crt_tab_illum_f2, crt_tab_illum_f2_n, true);
        h->illum[5] = PiecewiseLinear::from_interleaved(crt_tab_illum_f11, crt_tab_illum_f11_n, true);
        const PiecewiseLinear& d65 = h->illum[2];
        for (int l = 360; l <= 830; ++l) h->D65dense[l - 360] = d65.query((float)l);      // RGBColorSpace::illuminant
        // RGBColorSpace sRGB (colorspace.cpp:13-28, :82-100)
        float W[3];
        xyz_of(d65, W);
        float wsum = W[0] + W[1] + W[2];
        h->white[0] = W[0] / wsum; h->white[1] = W[1] / wsum;
        float R[3], G[3], B[3];
        from_xyY(.64, .33, R); from_xyY(.3, .6, G); from_xyY(.15, .06, B);
        const float rgb[9] = {R[0], R[1], R[2], G[0], G[1], G[2], B[0], B[1], B[2]};
        float rgb_inv[9], C[3];
        m3_inverse(rgb, rgb_inv);
        m3_mul_v3(rgb_inv, W, C);
        const float diag[9] = {C[0], 0, 0, 0, C[1], 0, 0, 0, C[2]};
        m3_mul(rgb, diag, h->XYZFromRGB);
        m3_inverse(h->XYZFromRGB, h->RGBFromXYZ);
        // PixelSensor(XYZ) white balance (pixelsensor.h:70-79; color.h:600-629): source white = xy of the sensor
        // illuminant (D65 as passed by RayTracerTestApp.h:149), target white = sRGB's
        const float LMSFromXYZ[9] = {0.8951, -0.7502, 0.0389, 0.2664, 1.7135, -0.0685, -0.1614, 0.0367, 1.0296};
        const float XYZFromLMS[9] = {0.986993, 0.432305, -0.00852866, -0.147054, 0.51836, 0.0400428, 0.159963, 0.0492912, 0.968487};
        float src[3], dst[3], srcLMS[3], dstLMS[3];
        from_xyY(h->white[0], h->white[1], src);       // sensor illuminant == colour-space illuminant -> same xy
        from_xyY(h->white[0], h->white[1], dst);
        m3_mul_v3(LMSFromXYZ, src, srcLMS);
        m3_mul_v3(LMSFromXYZ, dst, dstLMS);
        const float corr[9] = {dstLMS[0] / srcLMS[0], 0, 0, 0, dstLMS[1] / srcLMS[1], 0, 0, 0, dstLMS[2] / srcLMS[2]};
        float tmp[9];
        m3_mul(XYZFromLMS, corr, tmp);
        m3_mul(tmp, LMSFromXYZ, h->XYZFromSensorRGB);
        hs = h;
    });
    return *hs;
}

const float* swatch_table(int i, int* n) {
    if (i < 0 || i >= 24) { *n = 0; return nullptr; }
    *n = crt_tab_swatch_offsets[i + 1] - crt_tab_swatch_offsets[i];
    return crt_tab_swatches + crt_tab_swatch_offsets[i];
}
namespace {
struct Named { const char* k; const float* p; int n; };
const Named kNamed[] = {
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
}  // namespace
int named_table_count() { return (int)(sizeof kNamed / sizeof kNamed[0]); }
const float* named_table(const char* name, int* n) {
    for (const Named& e : kNamed)
        if (std::strcmp(e.k, name) == 0) { *n = e.n; return e.p; }
    *n = 0;
    return nullptr;
}

}  // namespace crt

extern "C" {
int crt_dense_table(int which, float* out) {
    const crt::HostSpectra& h = crt::host_spectra();
    const float* src = which == 0 ? h.X : which == 1 ? h.Y : which == 2 ? h.Z : which == 3 ? h.D65dense : nullptr;
    if (!src) { crt::set_error("dense_table: unknown table"); return 1; }
    std::memcpy(out, src, 471 * sizeof(float));
    return 0;
}
int crt_color_constants(float* sensor9, float* rgb_from_xyz9, float* xyz_from_rgb9, float* white2) {
    const crt::HostSpectra& h = crt::host_spectra();
    std::memcpy(sensor9, h.XYZFromSensorRGB, 36);
    std::memcpy(rgb_from_xyz9, h.RGBFromXYZ, 36);
    std::memcpy(xyz_from_rgb9, h.XYZFromRGB, 36);
    white2[0] = h.white[0]; white2[1] = h.white[1];
    return 0;
}
}
