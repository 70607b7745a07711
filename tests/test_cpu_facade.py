"""include/crt/facade.hpp -- the reference's class shapes over the C ABI.

CPU: the header compiles warning-free, its host half (TriModel, Octtree_Model build, Shape::Bounds / Area, samplers, CameraBase::generateRay)
runs without a GPU and agrees with the Python binding of the same C ABI; device methods throw (there is no CPU fallback).
GPU (-m gpu): a C++ program that uses ONLY the facade renders the reference's Li (Tier A) and the path integrator, traverses single rays and
intersects analytic shapes; its outputs must equal the ctypes path byte for byte."""
import os
import subprocess

import numpy as np
import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
PKG = os.path.join(ROOT, "computational_ray_tracer_b200")

PROGRAM = r'''
#include "crt/facade.hpp"
#include <cstdio>
#include <cmath>
using namespace crt;

static Model grid_model(int n) {                      // a small height-field-like grid with normals
    Model model; Mesh m;
    for (int j = 0; j <= n; ++j) for (int i = 0; i <= n; ++i) {
        m.positions.push_back(i * 10.f - 5.f * n); m.positions.push_back(j * 10.f - 5.f * n); m.positions.push_back(500.f + (i * j) % 7);
        m.normals.push_back(0.f); m.normals.push_back(0.f); m.normals.push_back(-1.f);
    }
    for (int j = 0; j < n; ++j) for (int i = 0; i < n; ++i) {
        uint32_t a = j * (n + 1) + i, b = a + 1, c = a + n + 1, d = c + 1;
        for (uint32_t v : {a, c, b, b, c, d}) m.indices.push_back(v);
    }
    model.meshes.push_back(m);
    return model;
}
static void dump(FILE* f, const void* p, size_t bytes) { fwrite(p, 1, bytes, f); }

int main(int argc, char** argv) {
    const bool device = argc > 2;
    FILE* out = fopen(argv[1], "wb");
    Model model = grid_model(24);
    TriModel tm(model, identity(), false, true);
    TriModel copy = tm;                               // copies re-point their mesh descriptors at their own buffers
    Bounds3 b = copy.Bounds();
    Octtree_Model oct(tm);
    oct.CreateOcttree();
    oct.PrintInfo();
    Octtree_Model::node root = oct.GetNode(0);
    int32_t host_ints[2] = {oct.getTreeSize(), (int32_t)root.leaf};
    dump(out, host_ints, sizeof host_ints);
    dump(out, &b, sizeof b);
    // analytic shapes: Bounds / Area on the host
    Sphere sphere("s", translation(30, -20, 400), 60.f, -30.f, 45.f, 270.f);
    Cylinder cyl("c", translation(-80, 10, 450), 40.f, -50.f, 70.f, 200.f);
    Disk disk("d", translation(0, 90, 480), 10.f, 25.f, 70.f, 300.f);
    TriangleSimple tri("t", translation(0, 0, 350), vec3(-60, -40, 0), vec3(70, -30, 5), vec3(0, 65, -10));
    Shape* shapes[4] = {&sphere, &cyl, &disk, &tri};
    for (Shape* s : shapes) { Bounds3 sb = s->Bounds(); float a = s->Area(); dump(out, &sb, sizeof sb); dump(out, &a, sizeof a); }
    // samplers and camera rays on the host
    StratifiedSampler strat(4, 4, true, 3);
    IndependentSampler indep(16, 5);
    PerspectiveCamera cam(1, 1000, 45, vec3(0, 0, 0), vec3(0, 0, 1), vec3(0, 1, 0), 160, 90);
    PerspectiveCamera lens_cam(1, 1000, 45, vec3(0, 0, 0), vec3(0, 0, 1), vec3(0, 1, 0), 160, 90, 20.f, 500.f);
    for (Sampler* s : {(Sampler*)&strat, (Sampler*)&indep}) {
        s->StartPixelSample(ivec2(37, 21), 5, 0);
        float u = s->Get1D(); vec2 p = s->GetPixel2D(); vec2 q = s->Get2D(); float w = s->Get1D();
        float vals[6] = {u, p.x, p.y, q.x, q.y, w};
        dump(out, vals, sizeof vals);
        s->StartPixelSample(ivec2(37, 21), 5, 0);
        (void)s->Get1D(); (void)s->GetPixel2D();
        Ray r1 = cam.generateRay(vec2(37.5f, 21.5f), s), r2 = lens_cam.generateRay(vec2(37.5f, 21.5f), s);
        dump(out, &r1, sizeof r1); dump(out, &r2, sizeof r2);
    }
    if (!device) {
        bool threw = false;
        try { Ray r(vec3(0, 0, 0), vec3(0, 0, 1)); oct.Traverse(r); } catch (const Error& e) { threw = true; std::printf("no gpu: %s\n", e.what()); }
        fclose(out);
        return threw && oct.getTreeSize() > 1 && !root.leaf ? 0 : 1;
    }
    // ---- device: single rays, a batch, shape intersections, two renders
    const int W = 160, H = 90;
    std::vector<Ray> rays;
    for (int y = 1; y <= H; y += 7) for (int x = 0; x < W; x += 11) { strat.StartPixelSample(ivec2(x, y), 0, 0); rays.push_back(cam.generateRay(vec2(x + .5f, y + .5f), nullptr)); }
    std::vector<std::optional<LocalSurfaceInfo>> batch = oct.TraverseBatch(rays.data(), (int)rays.size());
    int hits = 0;
    for (size_t i = 0; i < rays.size(); ++i) {
        std::optional<LocalSurfaceInfo> one = oct.Traverse(rays[i]);                  // the reference's call shape, one ray at a time
        if (one.has_value() != batch[i].has_value()) return 3;
        int32_t found = one.has_value();
        dump(out, &found, 4);
        if (one) { if (std::memcmp(&*one, &*batch[i], sizeof(LocalSurfaceInfo)) != 0) return 4; dump(out, &*one, sizeof(LocalSurfaceInfo)); ++hits; }
    }
    std::printf("traverse: %d of %zu rays hit\n", hits, rays.size());
    for (Shape* s : shapes)
        for (size_t i = 0; i < rays.size(); i += 3) {
            std::optional<LocalSurfaceInfo> h = s->Intersect(rays[i]);
            int32_t found = h.has_value(), occluded = s->IntersectP(rays[i], 1e30f);
            if (found != occluded) return 5;
            dump(out, &found, 4);
            if (h) dump(out, &*h, sizeof(LocalSurfaceInfo));
        }
    Scene scene;
    scene.SetModel(oct);
    scene.Commit();
    GaussianFilter gauss(vec2(1.5f, 1.5f), 0.5f);
    Film film(W, H, nullptr), film_g(W, H, &gauss);
    Integrator li;                                                                     // mode 0: the reference's Li
    li.Render(scene, film, cam, strat, 0, 4);
    li.Render(scene, film_g, lens_cam, strat, 0, 4);
    std::vector<float> px = film.Pixels(), pg = film_g.Pixels();
    std::vector<uint8_t> rgb = film.ResolveRGB8();
    dump(out, px.data(), px.size() * 4); dump(out, pg.data(), pg.size() * 4); dump(out, rgb.data(), rgb.size());
    // path integrator with two analytic shapes in the scene
    Scene scene2;
    int grey = scene2.AddConstantSpectrum(0.5f), d65 = scene2.AddStdIlluminant(2);
    int lambert = scene2.AddMaterial(0, grey), light = scene2.AddMaterial(0, -1, -1, -1, d65, 30.f);
    scene2.SetModel(oct, {lambert});
    scene2.Add(sphere, lambert);
    scene2.Add(tri, light);
    scene2.Commit();
    Film film2(W, H);
    Integrator path; path.mode = 1; path.max_depth = 4;
    path.Render(scene2, film2, cam, strat, 0, 4);
    std::vector<float> p2 = film2.Pixels();
    dump(out, p2.data(), p2.size() * 4);
    fclose(out);
    return 0;
}
'''


def _build(tmp_path):
    src = tmp_path / "facade_demo.cpp"
    src.write_text(PROGRAM)
    exe = tmp_path / "facade_demo"
    subprocess.run(["g++", "-std=c++17", "-O1", "-Wall", "-Wextra", "-Werror", "-I", os.path.join(ROOT, "include"), str(src), "-o", str(exe),
                    "-L", PKG, "-l:libcrt_b200.so", f"-Wl,-rpath,{PKG}"], check=True)
    return exe


class _Reader:
    def __init__(self, path):
        self.b = open(path, "rb").read(); self.at = 0

    def take(self, dtype, n):
        a = np.frombuffer(self.b, dtype, n, self.at); self.at += a.nbytes
        return a


def _grid_meshes(n=24):
    pos, nrm, idx = [], [], []
    for j in range(n + 1):
        for i in range(n + 1):
            pos.append((i * 10.0 - 5.0 * n, j * 10.0 - 5.0 * n, 500.0 + (i * j) % 7)); nrm.append((0, 0, -1))
    for j in range(n):
        for i in range(n):
            a = j * (n + 1) + i; b = a + 1; c = a + n + 1; d = c + 1
            idx += [(a, c, b), (b, c, d)]
    return [dict(positions=np.float32(pos), normals=np.float32(nrm), indices=np.uint32(idx))]


SHAPES = [(0, (30, -20, 400), [60.0, -30.0, 45.0, 270.0]), (1, (-80, 10, 450), [40.0, -50.0, 70.0, 200.0]),
          (2, (0, 90, 480), [10.0, 25.0, 70.0, 300.0]), (3, (0, 0, 350), [-60, -40, 0, 70, -30, 5, 0, 65, -10])]


def _check_host_half(rd, crt_lib):
    """The part of the program's output that needs no GPU, against the Python binding of the same C ABI."""
    import ctypes as C
    from computational_ray_tracer_b200 import _capi, api, scenes
    ms = api.MeshSet(_grid_meshes())
    oc = api.Octtree_Model(ms)
    ints = rd.take(np.int32, 2)
    assert ints[0] == oc.getTreeSize() and ints[1] == 0
    assert np.array_equal(rd.take(np.float32, 6), oc.model_bounds())
    for kind, t, params in SHAPES:
        p = np.float32(params + [0] * (9 - len(params))); b = np.zeros(6, np.float32); a = np.zeros(1, np.float32)
        _capi.check(crt_lib.crt_shape_bounds(kind, scenes.translation(*t).reshape(-1).ctypes.data_as(_capi.f32p), p.ctypes.data_as(_capi.f32p), b.ctypes.data_as(_capi.f32p)))
        _capi.check(crt_lib.crt_shape_area(kind, p.ctypes.data_as(_capi.f32p), a.ctypes.data_as(_capi.f32p)))
        assert np.array_equal(rd.take(np.float32, 6), b) and rd.take(np.float32, 1)[0] == a[0] and a[0] > 0
    r2c, c2w = api.camera_matrices(0, 1.0, 1000.0, 45.0, (0, 0, 0), (0, 0, 1), (0, 1, 0), 160, 90)
    for kind, xs, ys, jitter, seed in ((1, 4, 4, 1, 3), (0, 16, 1, 1, 5)):
        want = np.zeros(6, np.float32)
        _capi.check(crt_lib.crt_kat_sampler(kind, xs, ys, jitter, seed, 37, 21, 5, 0, b"1221", 0, want.ctypes.data_as(_capi.f32p)))
        got = rd.take(np.float32, 6)
        assert np.array_equal(got, want)                       # Get1D, GetPixel2D, Get2D, Get1D: the stream of samplers.h
        r1 = rd.take(np.float32, 6); r2 = rd.take(np.float32, 6)
        xy = np.float32([37.5, 21.5]); out = np.zeros(6, np.float32)
        _capi.check(crt_lib.crt_camera_generate_rays(0, r2c.ctypes.data_as(_capi.f32p), c2w.ctypes.data_as(_capi.f32p), 0.0, 0.0, xy.ctypes.data_as(_capi.f32p), None, 1, 0,
                                                     out.ctypes.data_as(_capi.f32p)))
        assert np.array_equal(r1, out) and abs(np.linalg.norm(r1[3:]) - 1) < 1e-6
        u = np.ascontiguousarray(got[3:5]); out2 = np.zeros(6, np.float32)          # the thin lens consumed the Get2D() after GetPixel2D
        _capi.check(crt_lib.crt_camera_generate_rays(0, r2c.ctypes.data_as(_capi.f32p), c2w.ctypes.data_as(_capi.f32p), 20.0, 500.0, xy.ctypes.data_as(_capi.f32p),
                                                     u.ctypes.data_as(_capi.f32p), 1, 0, out2.ctypes.data_as(_capi.f32p)))
        assert np.array_equal(r2, out2) and np.abs(r2[:2]).max() > 0
    oc.close()
    return r2c, c2w


def test_facade_compiles_and_its_host_half_runs_without_a_gpu(tmp_path, crt_lib):
    exe = _build(tmp_path)
    out = tmp_path / "host.bin"
    r = subprocess.run([str(exe), str(out)], capture_output=True, text=True)
    import torch
    if torch.cuda.is_available():
        pytest.skip("a GPU is present: the no-device branch of the program cannot be exercised")
    assert r.returncode == 0, r.stdout + r.stderr
    assert "octree:" in r.stdout and "no gpu:" in r.stdout
    _check_host_half(_Reader(out), crt_lib)


@pytest.mark.gpu
def test_a_cpp_program_using_only_the_facade_matches_the_ctypes_path(tmp_path, gpu_ctx, crt_lib):
    from computational_ray_tracer_b200 import api, scenes
    exe = _build(tmp_path)
    out = tmp_path / "device.bin"
    r = subprocess.run([str(exe), str(out), "device"], capture_output=True, text=True)
    assert r.returncode == 0, r.stdout + r.stderr
    rd = _Reader(out)
    r2c, c2w = _check_host_half(rd, crt_lib)
    W, H = 160, 90
    ms = api.MeshSet(_grid_meshes()); oc = api.Octtree_Model(ms)
    sc = api.Scene(gpu_ctx); sc.set_model(oc); sc.commit()
    xs_, ys_ = np.meshgrid(np.arange(0, W, 11), np.arange(1, H + 1, 7))
    xy = np.stack([xs_.reshape(-1) + 0.5, ys_.reshape(-1) + 0.5], 1).astype(np.float32)
    rays = np.zeros((len(xy), 6), np.float32)
    from computational_ray_tracer_b200 import _capi
    _capi.check(crt_lib.crt_camera_generate_rays(0, r2c.ctypes.data_as(_capi.f32p), c2w.ctypes.data_as(_capi.f32p), 0.0, 0.0, xy.ctypes.data_as(_capi.f32p), None, len(xy), 1,
                                                 rays.ctypes.data_as(_capi.f32p)))                       # on the device: must equal the host rays
    g = sc.traverse_local_surface(rays)
    assert g["found"].mean() > 0.5
    for i in range(len(rays)):
        found = rd.take(np.int32, 1)[0]
        assert found == g["found"][i]
        if found:
            rec = rd.take(np.float32, 18)                      # LocalSurfaceInfo: tHit hitp u v du dv n wo
            want = np.concatenate([[0.0], g["hitp"][i], g["uv"][i], g["du"][i], g["dv"][i], g["n"][i], g["wo"][i]]).astype(np.float32)
            assert np.array_equal(rec.view(np.uint32), want.view(np.uint32))
    for kind, t, params in SHAPES:
        ssc = api.Scene(gpu_ctx); ssc.add_shape(kind, scenes.translation(*t), params); ssc.commit()
        h = ssc.shape_intersect(0, rays[::3])
        for i in range(len(rays[::3])):
            found = rd.take(np.int32, 1)[0]
            assert found == h["found"][i]
            if found:
                rec = rd.take(np.float32, 18)
                assert rec[0] == h["t"][i] and np.array_equal(rec[1:4], h["hitp"][i]) and np.array_equal(rec[4:6], h["uv"][i]) and np.array_equal(rec[12:15], h["n"][i])
        ssc.close()
    # Tier A films: box filter / pinhole, and Gaussian filter / thin lens; the 8-bit resolve
    film = api.Film(gpu_ctx, W, H)
    kw = dict(mode=0, sampler_kind=1, xs=4, ys=4, jitter=1, seed=3, spp_begin=0, spp_end=4, trace_mode=3)
    sc.render(film, api.make_config(W, H, r2c, c2w, **kw))
    assert np.array_equal(rd.take(np.float32, W * H * 4).view(np.uint32), film.download().reshape(-1).view(np.uint32))
    rgb8 = film.resolve(want_float=False)[0]
    film.clear()
    sc.render(film, api.make_config(W, H, r2c, c2w, lens_radius=20.0, focal_distance=500.0, filter_kind=2, filter_r=(1.5, 1.5), filter_sigma=0.5, **kw))
    assert np.array_equal(rd.take(np.float32, W * H * 4).view(np.uint32), film.download().reshape(-1).view(np.uint32))
    assert np.array_equal(rd.take(np.uint8, W * H * 3), rgb8.reshape(-1)) and rgb8.max() > 30
    # path integrator with a sphere and an emissive TriangleSimple added through Shape objects
    sc2 = api.Scene(gpu_ctx)
    grey = sc2.add_spectrum(0, c=0.5); d65 = sc2.add_spectrum(4, n=2)
    lam = sc2.add_material(type=0, refl=grey); light = sc2.add_material(type=0, refl=-1, emit=d65, emit_scale=30.0)
    sc2.set_model(oc, mesh_materials=[lam])
    sc2.add_shape(SHAPES[0][0], scenes.translation(*SHAPES[0][1]), SHAPES[0][2], material=lam)
    sc2.add_shape(SHAPES[3][0], scenes.translation(*SHAPES[3][1]), SHAPES[3][2], material=light)
    sc2.commit()
    film.clear()
    sc2.render(film, api.make_config(W, H, r2c, c2w, **dict(kw, mode=1, max_depth=4)))
    got = rd.take(np.float32, W * H * 4)
    assert np.array_equal(got.view(np.uint32), film.download().reshape(-1).view(np.uint32)) and got.reshape(-1, 4)[:, :3].max() > 0
    assert rd.at == len(rd.b)
    film.close(); sc2.close(); sc.close(); oc.close()
