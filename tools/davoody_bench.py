"""Timing of the davoody table kernel for several tube sizes (GPU), with the reference's cost per placement beside it.

    python tools/davoody_bench.py [--ref]     # --ref also times the reference's first_order (oracle/_ref/libf1.so)
"""
import json
import os
import sys
import time

import numpy as np

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from cnt_film_monte_carlo_b200 import davoody as dv  # noqa: E402

with_ref = "--ref" in sys.argv
mc = {"zshift [m]": [1.5e-9, 10e-9, 11], "axis shift 1 [m]": [-10e-9, 10e-9, 11], "axis shift 2 [m]": [-10e-9, 10e-9, 11],
      "theta [degrees]": [0, 180, 21]}
axes = dv.table_axes(mc)
peak = dv.fp64_peak(0)
print(json.dumps({"fp64_fma_peak_tflops_measured": peak}), flush=True)
for spec in [(4, 2, 10), (4, 2, 20), (4, 2, 40), (6, 5, 10), (4, 2, 80)]:
    t = dv.Tube(*spec)
    t0 = time.time()
    x = dv.Transfer(t, t)
    t_x = time.time() - t0
    x.table(*axes)  # warm-up
    ms = []
    for _ in range(3):
        x.table(*axes)
        ms.append(x.info()["last_kernel_ms"])
    info = x.info()
    n_place = 21 * 11 * 11 * 11
    site_pairs = t.sites * t.sites
    passes = -(-info["donor_kcm"] // info["kcm_per_pass"])
    # FP64 operations of stage 1 per site pair and pass: 3 sub, 3 mul, 2 add, sqrt, div (counted as 1 each) + 2 fma per K_cm
    flop = n_place * site_pairs * passes * (10 + 4 * info["kcm_per_pass"])
    line = {"tube": spec, "sites": t.sites, "tube_build_s": t.build_seconds, "transfer_setup_s": t_x, "pairs": info["pairs"],
            "donor_kcm": info["donor_kcm"], "kcm_per_pass": info["kcm_per_pass"], "threads": info["threads"], "table_kernel_ms": min(ms),
            "site_pairs_per_s": n_place * site_pairs * passes / (min(ms) * 1e-3), "fp64_tflops_counted": flop / (min(ms) * 1e-3) / 1e12}
    line["roofline"] = {"bound": "fp64", "achieved": line["fp64_tflops_counted"], "peak": peak, "unit": "TFLOP/s",
                        "frac": line["fp64_tflops_counted"] / peak}
    if with_ref and t.sites <= 1200:
        from oracle import f1
        r = f1.RefTube(*spec)
        t0 = time.time()
        f1.first_order(r, r, 2e-9, 1e-9, -1e-9, 0.4)
        line["ref_s_per_placement"] = time.time() - t0
        line["ref_table_s_one_thread"] = line["ref_s_per_placement"] * n_place
    print(json.dumps(line), flush=True)
