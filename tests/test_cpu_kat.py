"""Known-answer tests of the integer sampler stack (CPU): oracle and the product's host code against published
vectors and an independent pure-Python restatement.  The reference has no tests of its own; the
PCG32 vector is the official pcg32-demo output, everything else is cross-implementation agreement."""
import ctypes as C

import numpy as np
import pytest

import oracle_lib as O
from computational_ray_tracer_b200 import scenes
from computational_ray_tracer_b200._capi import f32p, i32p, u32p, u64p

M64 = (1 << 64) - 1


def murmur64a_py(key: bytes, seed: int) -> int:
    """MurmurHash64A (Austin Appleby, public domain) -- ThirdParty/pbrv4/hash.h:18-63."""
    m, r = 0xC6A4A7935BD1E995, 47
    h = (seed ^ (len(key) * m)) & M64
    nblocks = len(key) // 8
    for i in range(nblocks):
        k = int.from_bytes(key[8 * i:8 * i + 8], "little")
        k = (k * m) & M64
        k ^= k >> r
        k = (k * m) & M64
        h ^= k
        h = (h * m) & M64
    tail = key[8 * nblocks:]
    if tail:
        for i in range(len(tail) - 1, -1, -1):
            h ^= tail[i] << (8 * i)
        h = (h * m) & M64
    h ^= h >> r
    h = (h * m) & M64
    h ^= h >> r
    return h


def permutation_element_py(i, l, p):
    """PermutationElement (Util/HelperFunctions.h:175-203; Kensler's correlated multi-jittered sampling)."""
    M32 = 0xFFFFFFFF
    w = l - 1
    w |= w >> 1; w |= w >> 2; w |= w >> 4; w |= w >> 8; w |= w >> 16
    while True:
        i ^= p; i = (i * 0xe170893d) & M32
        i ^= p >> 16
        i ^= (i & w) >> 4
        i ^= p >> 8; i = (i * 0x0929eb3f) & M32
        i ^= p >> 23
        i ^= (i & w) >> 1; i = (i * (1 | p >> 27)) & M32
        i = (i * 0x6935fa69) & M32
        i ^= (i & w) >> 11; i = (i * 0x74dcb303) & M32
        i ^= (i & w) >> 2; i = (i * 0x9e501cc3) & M32
        i ^= (i & w) >> 2; i = (i * 0xc860a3df) & M32
        i &= w
        i ^= i >> 5
        if i < l:
            break
    return (i + p) % l


PCG32_DEMO = [0xa15c02b7, 0x7b47f409, 0xba1d3330, 0x83d2f293, 0xbfa4784b, 0xcbed606e]     # pcg32_srandom(42, 54)


def test_pcg32_official_vector(oracle, crt_lib):
    out = np.zeros(6, np.uint32)
    oracle.lib().orc_pcg32(2, 54, 42, 0, 6, O.up(out), None)
    assert [int(x) for x in out] == PCG32_DEMO
    out2 = np.zeros(6, np.uint32)
    assert crt_lib.crt_kat_pcg32(2, 54, 42, 0, 6, 0, out2.ctypes.data_as(u32p), None) == 0
    assert [int(x) for x in out2] == PCG32_DEMO
    r = scenes.PCG32.__new__(scenes.PCG32)           # the scene generator's PCG32 restatement, same stream
    r.state = 0; r.inc = (54 << 1) | 1; r.uniform_u32(); r.state = (r.state + 42) & M64; r.uniform_u32()
    assert [r.uniform_u32() for _ in range(6)] == PCG32_DEMO


@pytest.mark.parametrize("mode,seq,off,adv", [(0, 0, 0, 0), (1, 7, 0, 0), (1, 123456789, 0, 65536 * 3 + 5), (2, 9, 11, -17), (2, 2 ** 40 + 3, 5, 2 ** 33)])
def test_pcg32_streams_and_advance(oracle, crt_lib, mode, seq, off, adv):
    a = np.zeros(16, np.uint32); b = np.zeros(16, np.uint32)
    oracle.lib().orc_pcg32(mode, seq, off, adv, 16, O.up(a), None)
    assert crt_lib.crt_kat_pcg32(mode, seq, off, adv, 16, 0, b.ctypes.data_as(u32p), None) == 0
    assert np.array_equal(a, b)
    fa = np.zeros(16, np.float32); fb = np.zeros(16, np.float32)
    oracle.lib().orc_pcg32(mode, seq, off, adv, 16, None, O.fp(fa))
    assert crt_lib.crt_kat_pcg32(mode, seq, off, adv, 16, 0, None, fb.ctypes.data_as(f32p)) == 0
    assert np.array_equal(fa.view(np.uint32), fb.view(np.uint32))
    assert np.array_equal(fa, np.minimum(np.float32(1.0), a.astype(np.float32) * np.float32(2.0 ** -32)))   # rng.h:122-124 with pch.h:37
    if adv > 0 and mode == 1:    # Advance(k) == k draws
        c = np.zeros(adv + 16, np.uint32) if adv < 1 << 20 else None
        if c is not None:
            oracle.lib().orc_pcg32(mode, seq, off, 0, adv + 16, O.up(c), None)
            assert np.array_equal(c[adv:], a)


def test_murmur_and_hash_layout(oracle, crt_lib):
    rs = np.random.RandomState(0)
    for n in list(range(0, 33)) + [100, 255]:
        key = bytes(rs.randint(0, 256, n, dtype=np.uint8).tolist())
        seed = int(rs.randint(0, 2 ** 31)) * 7919
        want = murmur64a_py(key, seed)
        assert oracle.lib().orc_murmur64a(key, n, seed) == want
        out = C.c_uint64()
        assert crt_lib.crt_kat_hash(key, n, seed, 0, C.byref(out)) == 0
        assert out.value == want
    # Hash(ivec2, int) = MurmurHash64A over the 12 packed bytes, seed 0 (hash.h:96-104; samplers.h:47-51)
    for (x, y, s) in [(0, 0, 0), (3, 1080, 0), (1919, 1, 7), (-5, 17, 123456)]:
        key = np.array([x, y, s], np.int32).tobytes()
        assert oracle.lib().orc_hash_pixel_seed(x, y, s) == murmur64a_py(key, 0)
        key4 = np.array([x, y, 5, s], np.int32).tobytes()
        assert oracle.lib().orc_hash_pixel_dim_seed(x, y, 5, s) == murmur64a_py(key4, 0)


def test_mixbits(oracle):
    for v in [0, 1, 54, 2 ** 63 + 12345, M64]:
        assert oracle.lib().orc_mixbits(v) == scenes.mix_bits(v)


def test_permutation_element(oracle, crt_lib):
    rs = np.random.RandomState(1)
    i = rs.randint(0, 64, 512).astype(np.uint32); l = np.full(512, 64, np.uint32); p = rs.randint(0, 2 ** 31, 512).astype(np.uint32)
    l[256:] = 100; i[256:] = rs.randint(0, 100, 256)
    out = np.zeros(512, np.int32)
    assert crt_lib.crt_kat_permutation(i.ctypes.data_as(u32p), l.ctypes.data_as(u32p), p.ctypes.data_as(u32p), 512, 0, out.ctypes.data_as(i32p)) == 0
    for k in range(512):
        want = permutation_element_py(int(i[k]), int(l[k]), int(p[k]))
        assert out[k] == want
        assert oracle.lib().orc_permutation_element(int(i[k]), int(l[k]), int(p[k])) == want
    # it is a permutation of [0, l)
    for ll, pp in [(64, 12345), (100, 99), (7, 3)]:
        vals = sorted(oracle.lib().orc_permutation_element(k, ll, pp) for k in range(ll))
        assert vals == list(range(ll))


@pytest.mark.parametrize("kind,xs,ys,jitter", [(0, 4, 4, 0), (1, 4, 4, 1), (1, 8, 8, 1), (1, 3, 5, 0)])
def test_sampler_sequences_host_equals_oracle(oracle, crt_lib, kind, xs, ys, jitter):
    pattern = b"12p2121"
    nout = 1 + 2 + 2 + 2 + 1 + 2 + 1
    for (px, py, idx, dim, seed) in [(0, 1, 0, 0, 0), (17, 33, 5, 0, 0), (1919, 1080, xs * ys - 1, 3, 9)]:
        a = np.zeros(nout, np.float32); b = np.zeros(nout, np.float32)
        oracle.lib().orc_sampler_sequence(kind, xs, ys, jitter, seed, px, py, idx, dim, pattern, O.fp(a))
        assert crt_lib.crt_kat_sampler(kind, xs, ys, jitter, seed, px, py, idx, dim, pattern.replace(b"p", b"2"), 0, b.ctypes.data_as(f32p)) == 0
        assert np.array_equal(a.view(np.uint32), b.view(np.uint32))
        assert (a >= 0).all() and (a <= 1).all()


def test_stratified_samples_land_in_their_strata(oracle):
    xs, ys = 4, 4
    seen = set()
    for idx in range(xs * ys):
        a = np.zeros(3, np.float32)
        oracle.lib().orc_sampler_sequence(1, xs, ys, 1, 0, 5, 9, idx, 0, b"12", O.fp(a))
        seen.add((int(a[1] * xs), int(a[2] * ys)))
    assert len(seen) == xs * ys          # one sample per stratum of the first 2D dimension (samplers.h:109-123)


def test_gaussian_filter_host_matches_oracle(oracle, crt_lib):
    """GaussianFilter::Sample (filters.h:96-163): the product's host evaluation of the shared host/device routine against the oracle
    (which tests/test_cpu_ref_pin.py ties to the reference's compiled code): positions and weights bit for bit, NaNs in the same places."""
    import ctypes as C
    import ref_pin_cases as P
    from computational_ray_tracer_b200._capi import f32p
    u, params = P.gaussian_inputs()
    for rx, ry, sg in params:
        a = np.zeros((len(u), 3), np.float32); b = np.zeros((len(u), 3), np.float32)
        oracle.lib().orc_gaussian_filter_samples(rx, ry, sg, oracle.fp(u), len(u), oracle.fp(a))
        assert crt_lib.crt_kat_gaussian_filter(rx, ry, sg, u.ctypes.data_as(f32p), len(u), 0, b.ctypes.data_as(f32p)) == 0
        same = (a.view(np.uint32) == b.view(np.uint32)) | (np.isnan(a) & np.isnan(b))
        assert same.all(), (rx, ry, sg, int((~same).sum()))
        assert np.abs(a[:, 0]).max() <= rx * (1 + 1e-6) and np.abs(a[:, 1]).max() <= ry * (1 + 1e-6)


def test_measured_sensor_matrix_matches_the_reference(crt_lib):
    """PixelSensor's measured-sensor constructor (pixelsensor.h:37-68): the product's host computation of XYZFromSensorRGB against the
    matrix the reference's compiled code produced for the same response curves (tests/golden/ref_pin.npz, group `sensor`)."""
    import os
    import ref_pin_cases as P
    from computational_ray_tracer_b200._capi import f32p
    gold = np.load(os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden", "ref_pin.npz"))
    for i in range(len(P.SENSORS)):
        cur = np.ascontiguousarray(gold[f"sensor/sensor{i}.curves"]); ill = np.ascontiguousarray(gold[f"sensor/sensor{i}.illum"])
        want = gold[f"sensor/sensor{i}.matrix"]
        got = np.zeros(9, np.float32)
        r, g, b = (np.ascontiguousarray(cur[k]) for k in range(3))
        assert crt_lib.crt_measured_sensor_matrix(r.ctypes.data_as(f32p), g.ctypes.data_as(f32p), b.ctypes.data_as(f32p), ill.ctypes.data_as(f32p),
                                                  got.ctypes.data_as(f32p)) == 0
        assert np.array_equal(got.view(np.uint32), want.view(np.uint32)), (i, got, want)
