"""RGB -> spectrum table (RGBToSpectrumTable, color.h:405-432 / color.cpp:26-166): host-side halves, no GPU.

The reference reads the table from a data file its repository does not contain.  The LOOKUP is pinned: the oracle's
restatement equals the reference's compiled operator() bit for bit (ref_pin_cases._rgb2spec / golden/ref_pin.npz), and here
the PRODUCT's lookup (csrc/crt_rgb2spec.cuh rgb2spec_lookup) equals the oracle's.  The table GENERATOR (restated Jakob-Hanika
optimiser) has no reference output to compare with; it is validated by round trips: RGB -> coefficients -> spectrum -> XYZ
under D65 by the renderer's own 1 nm quadrature -> RGB."""
import os
import subprocess
import sys

import numpy as np
import pytest

import oracle_lib as O
import ref_lib as R
import ref_pin_cases as P
from computational_ray_tracer_b200 import _capi, api


def _bits(a):
    return np.ascontiguousarray(a, np.float32).view(np.uint32)


def test_product_lookup_equals_oracle_lookup(oracle, crt_lib):
    scale, data = P.pseudo_rgb_table()
    L = O.lib()
    L.orc_set_rgb_table(O.fp(scale), O.fp(data))
    try:
        for rgb in P.rgb_inputs():
            want = np.zeros(3, np.float32)
            assert L.orc_rgb_coeffs(O.fp(rgb), O.fp(want)) == 0
            got = api.rgb2spec_lookup(scale, data, np.maximum(rgb, 0))
            assert np.array_equal(_bits(got), _bits(want)), rgb
    finally:
        L.orc_set_rgb_table(None, None)


def test_non_grey_without_a_table_is_refused(oracle):
    L = O.lib()
    L.orc_set_rgb_table(None, None)
    out = np.zeros(3, np.float32)
    assert L.orc_rgb_coeffs(O.fp(np.float32([0.2, 0.4, 0.6])), O.fp(out)) == -1
    assert L.orc_rgb_coeffs(O.fp(np.float32([0.4, 0.4, 0.4])), O.fp(out)) == 0        # grey bypasses the table (color.cpp:35-37)


def _roundtrip_model(crt_lib):
    def dense(w):
        a = np.zeros(471, np.float32); crt_lib.crt_dense_table(w, a.ctypes.data_as(_capi.f32p)); return a.astype(np.float64)
    X, Y, Z, D = dense(0), dense(1), dense(2), dense(3)
    m = [np.zeros(9, np.float32) for _ in range(3)]; w2 = np.zeros(2, np.float32)
    crt_lib.crt_color_constants(*[a.ctypes.data_as(_capi.f32p) for a in m], w2.ctypes.data_as(_capi.f32p))
    rgb_from_xyz = m[1].reshape(3, 3).T.astype(np.float64)
    lam = np.arange(360, 831, dtype=np.float64)

    def rgb_of(c):
        c = np.asarray(c, np.float64)
        x = c[0] * lam * lam + c[1] * lam + c[2]
        s = 0.5 + x / (2 * np.sqrt(1 + x * x))
        xyz = np.array([(X * s * D).sum(), (Y * s * D).sum(), (Z * s * D).sum()]) / (Y * D).sum()
        return rgb_from_xyz @ xyz
    return rgb_of


def test_cell_fit_round_trips(crt_lib):
    rgb_of = _roundtrip_model(crt_lib)
    rs = np.random.RandomState(0)
    cols = [(0.8, 0.2, 0.1), (0.1, 0.5, 0.9), (1.0, 0.0, 0.0), (0.0, 1.0, 0.0), (0.0, 0.0, 1.0), (0.05, 0.9, 0.05), (0.7, 0.7, 0.1)] + [tuple(r) for r in rs.rand(300, 3)]
    err = np.array([np.abs(rgb_of(api.rgb2spec_fit(c)) - np.array(c)).max() for c in cols])
    assert err[7:].max() < 1e-4, err[7:].max()          # random interior colours
    assert err[:7].max() < 2e-3, err[:7].max()          # saturated primaries sit on the gamut boundary: the coefficient clamp (|c| <= 200) binds
    c = np.array([api.rgb2spec_fit(c) for c in cols[:7]])
    assert np.isfinite(c).all()


def test_file_layout_round_trip_and_error_paths(crt_lib, tmp_path):
    scale, data = P.pseudo_rgb_table(3)
    path = tmp_path / "sRGB64binary"
    api.rgb2spec_save(path, scale, data)
    assert os.path.getsize(path) == 4 + 64 * 4 + data.size * 4          # color.cpp:117-150: int, 64 floats, 3*64^3*3 floats
    s2, d2 = api.rgb2spec_load(path)
    assert np.array_equal(_bits(s2), _bits(scale)) and np.array_equal(_bits(d2), _bits(data))
    with pytest.raises(_capi.CrtError):
        api.rgb2spec_load(tmp_path / "missing")
    (tmp_path / "short").write_bytes(b"\0" * 100)
    with pytest.raises(_capi.CrtError):
        api.rgb2spec_load(tmp_path / "short")


@pytest.mark.skipif(not R.available(), reason="compiled reference absent")
def test_reference_init_reads_the_file_we_write(crt_lib, tmp_path):
    """RGBToSpectrumTable::Init (color.cpp:107-166) opens ../rgb2spec/sRGB64binary relative to the working directory: run the
    compiled reference in a fresh process below such a tree and compare a non-grey albedo spectrum with the in-memory path."""
    scale, data = P.pseudo_rgb_table()
    (tmp_path / "rgb2spec").mkdir(); (tmp_path / "run").mkdir()
    api.rgb2spec_save(tmp_path / "rgb2spec" / "sRGB64binary", scale, data)
    R.build()
    code = ("import sys, numpy as np; sys.path.insert(0, %r); import ref_lib as R; L = R.lib();"
            "rgb = np.float32([0.2, 0.7, 0.4]); lam = np.float32([400, 500, 600, 700]); out = np.zeros(4, np.float32);"
            "L.ref_rgb_albedo_query(R.fp(rgb), R.fp(lam), 4, R.fp(out)); print(' '.join(str(int(v)) for v in out.view(np.uint32)))") % os.path.dirname(os.path.abspath(__file__))
    r = subprocess.run([sys.executable, "-c", code], cwd=tmp_path / "run", capture_output=True, text=True, check=True)
    from_file = np.array([int(v) for v in r.stdout.strip().splitlines()[-1].split()], np.uint32)
    L = R.lib()
    L.ref_set_rgb_table(R.fp(scale), R.fp(data))
    out = np.zeros(4, np.float32)
    L.ref_rgb_albedo_query(R.fp(np.float32([0.2, 0.7, 0.4])), R.fp(np.float32([400, 500, 600, 700])), 4, R.fp(out))
    assert np.array_equal(out.view(np.uint32), from_file)
    assert 0 < out.min() and out.max() < 1
