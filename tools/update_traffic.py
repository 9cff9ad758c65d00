#!/usr/bin/env python
"""profiles/roofline_traffic.json from the two `ncu --set full` captures of tools/gpu.sh prof:
   python tools/update_traffic.py gpurun_out/prof_celltile.ncu-rep gpurun_out/prof_celltile_mixed.ncu-rep
Records the hash of the kernel sources and the commit the captures were taken at: bench.py drops the
traffic figure when the tree's kernel sources differ (bench.roofline_traffic)."""
import csv, io, json, os, subprocess, sys
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import bench  # noqa: E402


def first_kernel(rep):
    txt = subprocess.run(["ncu", "-i", rep, "--page", "raw", "--csv"], capture_output=True, text=True).stdout
    rows = list(csv.reader(io.StringIO(txt)))
    d = dict(zip(rows[0], rows[2])); u = dict(zip(rows[0], rows[1]))
    def num(k):
        v = float(d[k].replace(",", ""))
        return v * {"Kbyte": 1e3, "Mbyte": 1e6, "Gbyte": 1e9, "byte": 1.0}.get(u[k], 1.0)
    return d["Kernel Name"], num("dram__bytes_read.sum"), num("dram__bytes_write.sum"), d["gpu__time_duration.sum"] + " " + u["gpu__time_duration.sum"]


k, r, w, t = first_kernel(sys.argv[1])
km, rm, wm, tm = first_kernel(sys.argv[2])
head = subprocess.run(["git", "-C", ROOT, "rev-parse", "--short", "HEAD"], capture_output=True, text=True).stdout.strip()
out = {
    "kernel": k + "  [FP64, AUTO at N=1M]", "dram_bytes_read": r, "dram_bytes_write": w, "dram_bytes_per_launch": r + w,
    "duration_under_ncu": t,
    "note": "below the algorithmic 651 MB: the kernel streams the library's 16-bit mirror list (2 B per pair) "
            "instead of the reference's 4-byte list",
    "source": "profiles/r02_celltile.summary.txt (ncu --set full --clock-control none, config C, one launch)",
    "kernel_source_sha": bench.kernel_source_sha(), "git_head": head,
    "mixed_kernel": {"kernel": km + "  [LJ_PREC_MIXED, wide tiles]", "dram_bytes_per_launch": rm + wm,
                     "duration_under_ncu": tm, "source": "profiles/r02_celltile_mixed.summary.txt"},
}
json.dump(out, open(os.path.join(ROOT, "profiles", "roofline_traffic.json"), "w"), indent=1)
print(json.dumps(out, indent=1))
