#!/usr/bin/env python
"""Markdown tables of the round's committed bench lines (profiles/r02_*.json) for DESIGN.md / BASELINE.md.
usage: tools/make_tables.py bench|scaling"""
import json, os, sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
P = os.path.join(ROOT, "profiles")


def load(name):
    path = os.path.join(P, name)
    if not os.path.exists(path):
        return None
    lines = [ln for ln in open(path) if ln.startswith("{")]
    return json.loads(lines[-1]) if lines else None


def bench():
    print("| config | Mpaths/s | Mrays/s | e2e (scene re-uploaded every step) | e2e (scene resident) | ms/step | launches/step | host oracle, 16 cores (clean / faithful) | GPU ÷ CPU clean |")
    print("|---|---|---|---|---|---|---|---|---|")
    for label, name in (("C1", "r02_bench_C1.json"), ("C2 (headline)", "r02_bench.json"), ("C3", "r02_bench_C3.json"), ("C4", "r02_bench_C4.json"),
                        ("C5, 16 of 1024 spp", "r02_bench_C5_1gpu_16spp.json")):
        d = load(name)
        if not d: continue
        cb = d.get("cpu_baseline")
        cpu = f"{cb['value']:.3f} / {cb['faithful_value']:.3f}" if cb else "—"
        ratio = f"×{d['value'] / cb['value']:.0f}" if cb else "—"
        print(f"| {label} | **{d['value']:.1f}** | {d['mrays_per_s']:.0f} | {d['e2e']['value']:.1f} | {d['e2e']['resident']['value']:.1f} | {d['ms_per_step']:.2f} | "
              f"{d['gpu_launches'] // d['steps']} | {cpu} | {ratio} |")


def scaling():
    print("| config, partition | GPUs | Mpaths/s | Mrays/s | e2e resident | ms/step | vs 1 GPU |")
    print("|---|---|---|---|---|---|---|")
    for cfg, part, pat in (("C2", "spp", "r02_C2_n{n}.json"), ("C2", "tiles", "r02_C2_n{n}_tiles.json"), ("C4", "spp", "r02_C4_n{n}.json"),
                           ("C5 (4K, 1024 spp, 10 M triangles)", "spp", "r02_C5_n{n}_spp.json"), ("C5", "tiles", "r02_C5_n{n}_tiles.json")):
        base = load(pat.format(n=1).replace("_tiles", "")) or (load("r02_C2_n1.json") if cfg == "C2" else None)
        for n in (1, 2, 4, 8):
            d = load(pat.format(n=n))
            if not d: continue
            eff = f"{d['value'] / base['value']:.2f}× ({100 * d['value'] / base['value'] / n:.1f} %)" if base and base["metric"] == d["metric"] else "—"
            print(f"| {cfg}, {part} | {n} | {d['value']:.1f} | {d['mrays_per_s']:.0f} | {d['e2e']['resident']['value']:.1f} | {d['ms_per_step']:.2f} | {eff} |")


{"bench": bench, "scaling": scaling}[sys.argv[1]]()
