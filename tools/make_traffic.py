#!/usr/bin/env python
"""profiles/<tag>_ncu_summary.json (tools/ncu_summary.py) -> profiles/traffic.json: DRAM bytes per launch of every
kernel of the path, keyed by the stage names bench.py reports.  usage: python tools/make_traffic.py r2 [c2|c3|c4]   -> profiles/traffic_<config>.json (profiles/<tag>_<config>_ncu_summary.json)"""
import json, os, re, sys
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
tag = sys.argv[1]
cfg = sys.argv[2] if len(sys.argv) > 2 else "c2"
GRID = {"c2": (45, 80, 128), "c3": (68, 120, 64), "c4": (135, 240, 16)}[cfg]
rows = json.load(open(os.path.join(ROOT, "profiles", f"{tag}_{cfg}_ncu_summary.json")))
RULES = [("tensorise_frames", r"tensorise_frames_kernel"), ("tc_enc1_fused", r"enc1_fused_kernel"),
         ("tc_enc2", r"enc_ws_kernel<.*ECfg<2,"), ("tc_enc3", r"enc_ws_kernel<.*ECfg<4,|shiftgemm_kernel<.*Cfg<0, 4,"),
         ("tc_enc4", r"enc_ws_kernel<.*ECfg<8,|shiftgemm_kernel<.*Cfg<0, 8,"),
         ("tc_dec0", r"shiftgemm_kernel<.*Cfg<1, 16, 128, \d+, \d+, 64>"), ("tc_dec1", r"shiftgemm_kernel<.*Cfg<1, 16, 128, \d+, \d+, 32>"),
         ("tc_dec2", r"shiftgemm_kernel<.*Cfg<1, 8,"), ("tc_dec3_head", r"shiftgemm_kernel<.*Cfg<2,"), ("ccl_bbox", r"ccl_bbox_kernel")]
traffic, sass = {}, {}
for r in rows:
    for name, pat in RULES:
        if re.search(pat, r["kernel"]) and "dram_bytes" in r:
            traffic[name] = int(r["dram_bytes"]); sass[name] = r["kernel"]
json.dump({"source": f"ncu --set full --clock-control none, CONFIG={cfg} tools/ncu_step.py ({GRID[2]} chains x 67 frames of {GRID[1]}x{GRID[0]} MB "
                     f"= {GRID[2] * 64} windows, one chunk), profiles/{tag}_{cfg}_ncu_summary.json", "grid": [GRID[0], GRID[1]],
           "windows_per_launch": GRID[2] * 64, "dram_bytes_per_launch": traffic, "dram_bytes_per_step": sum(traffic.values()), "sass_kernel": sass},
          open(os.path.join(ROOT, "profiles", f"traffic_{cfg}.json"), "w"), indent=1)
print(traffic)
