#!/usr/bin/env python
"""Print the markdown rows of BASELINE.md section 4 from profiles/r01_configs.json (output of tools/run_configs.py)."""
import json
import os
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
res = {r["config"]: r for r in json.load(open(sys.argv[1] if len(sys.argv) > 1 else os.path.join(ROOT, "profiles", "r01_configs.json")))}
print("| config | triangles | CPU clean Mpaths/s | 1×B200 Mpaths/s / Mrays/s (exact BFS kernel Mrays/s) | GPU ÷ CPU | rays/path | RMSE vs oracle | ordered vs BFS: id mismatches on primary rays; film bits over N rays | order-sensitive rays re-traced |")
print("|---|---|---|---|---|---|---|---|---|")
for k in sorted(res):
    r = res[k]
    cpu = f"{r['cpu_mpaths_s']:.3f} ({r['cpu_cores']} cores; {r['cpu_sample'].split(' in ')[0]})" if "cpu_mpaths_s" in r else "not run (oracle build of 10 M triangles ≈ 16 min)"
    rm = f"{r['rmse_mean_rgb']:.1e} on {r['oracle_pixels']} px" if "rmse_mean_rgb" in r else "— (GPU-only checks)"
    sp = f"×{r['speedup_vs_cpu']:.0f}" if "speedup_vs_cpu" in r else "—"
    eq = "identical" if r.get("films_bit_identical_across_trace_modes") else "DIFFERENT"
    print(f"| {k} {r['name']} | {r['triangles']:,} | {cpu} | {r['gpu_mpaths_s']:.1f} / {r['gpu_mrays_s']:.0f} ({r.get('bfs_kernel_mrays_s', 0):.0f}) | {sp} | {r['rays_per_path']:.2f} | {rm} | "
          f"{r['ordered_vs_bfs_mismatches']} of {r['primary_rays_checked']:,}; {eq} over {r.get('rays_in_that_check', 0):,} | {r['exact_retraced_rays']} |")
