"""profiles/<round>_ncu_post_kernels.md from gpurun_out/<round>_post_raw.csv (tools/profile_post.py under ncu --set full)."""
import csv
import os
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
R = sys.argv[1] if len(sys.argv) > 1 else "r01"
rows = list(csv.reader(open(os.path.join(ROOT, "gpurun_out", R + "_post_raw.csv"))))
hdr, units = rows[0], rows[1]
ix = {h: i for i, h in enumerate(hdr)}
want = ["gpu__time_duration.sum", "dram__bytes_read.sum", "dram__bytes_write.sum",
        "gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed", "sm__throughput.avg.pct_of_peak_sustained_elapsed",
        "launch__grid_size", "launch__registers_per_thread"]
B, HW = 256, 76800
# algorithmic bytes per launch: what the kernel must read and write once
alg = {"preprocess_u8_kernel": B * HW * (1 + 4), "seg_post_kernel": B * (3 * HW * 4 + HW), "seg_counts_kernel": B * HW * 2,
       "seg_loss_kernel": B * HW * (12 + 1 + 4 + 12)}
out = ["# %s: ncu --set full --clock-control none, bandwidth-bound post-processing kernels on a 256-frame batch" % R,
       "# (tools/profile_post.py, warm pass; HBM peak: 6650 GB/s fallback of B200_PROFILING.md when MEASURED_PEAKS.json is absent)", ""]
for r in rows[2:]:
    k = r[ix["Kernel Name"]].split("(")[0]
    out.append("## " + k)
    vals = {}
    for w in want:
        if w in ix:
            out.append("- %s: %s %s" % (w, r[ix[w]], units[ix[w]]))
            vals[w] = (r[ix[w]], units[ix[w]])
    t = float(vals["gpu__time_duration.sum"][0].replace(",", ""))
    u = vals["gpu__time_duration.sum"][1]
    t_us = t if u in ("us", "usecond") else (t / 1000 if u.startswith("n") else t * 1000)
    if k in alg:
        gbs = alg[k] / 1e3 / t_us
        out.append("- algorithmic bytes per launch: %.1f MB -> %.0f GB/s achieved = %.0f %% of 6650 GB/s" % (alg[k] / 1e6, gbs, 100 * gbs / 6650))
    out.append("")
open(os.path.join(ROOT, "profiles", R + "_ncu_post_kernels.md"), "w").write("\n".join(out))
print("\n".join(l for l in out if l.startswith("## ") or "algorithmic" in l or "time_duration" in l))
