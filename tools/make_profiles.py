"""Turn the gpurun_out/ artefacts of tools/capture_profiles.sh <tag> into tracked summaries under profiles/."""
import csv
import json
import shutil
import subprocess
import sys

sys.path.insert(0, "tools")
from ncu_summary import KEYS, launches  # noqa: E402

tag = sys.argv[1]
G = "gpurun_out"


def full(rep, dst, header, pick=None):
    out = subprocess.run(["ncu", "-i", rep, "--page", "raw", "--csv"], capture_output=True, text=True).stdout
    rows = list(csv.reader(out.splitlines()))
    hdr, units = rows[0], rows[1]
    with open(dst, "w") as f:
        f.write(header)
        for idx, r in enumerate(rows[2:]):
            if pick is not None and idx not in pick:
                continue
            d = dict(zip(hdr, r))
            f.write(f"\n## launch {idx}: {d['Kernel Name'][:70]}\n")
            for k in KEYS:
                if k in d:
                    f.write(f"{k:80s} {d[k]:>18s} {units[hdr.index(k)]}\n")
            for k in hdr:
                if ("average_warps_issue_stalled" in k and k.endswith("per_issue_active.ratio")
                        and "not_issued" not in k and d[k] and float(d[k]) > 0.05):
                    f.write(f"{k:80s} {d[k]:>18s}\n")


cmd = "bench.py --steps 10 --warmup 3 --no-cpu-baseline"
launches(f"{G}/launches_{tag}.csv", f"profiles/{tag}_launches.txt")
full(f"{G}/prof_binned_{tag}.ncu-rep", f"profiles/{tag}_density_binned_full.txt",
     f"# ncu --set full --clock-control none --import-source on -k regex:k_density_binned -s 10 -c 2 over {cmd}\n"
     "# dominant kernel of the timed step: a wave of 32 evaluations per launch\n")
full(f"{G}/prof_dense32_{tag}.ncu-rep", f"profiles/{tag}_density_dense32_full.txt",
     f"# ncu --set full --clock-control none --import-source on -k regex:k_density_dense32 -c 1 over {cmd}\n"
     "# the dense fp32 brute-force kernel (north-star kernel): 8 evaluations per launch\n")
full(f"{G}/prof_small_{tag}.ncu-rep", f"profiles/{tag}_other_kernels_full.txt",
     f"# ncu --set full --clock-control none, the seven other kernels of one wave, over {cmd}\n")
for name in (f"bench_{tag}_n1.json", f"bench_{tag}_reference.json"):
    shutil.copy(f"{G}/{name}", f"profiles/{name}")
d = json.load(open(f"{G}/bench_{tag}_n1.json"))
r = d["roofline"]
print("value", d["value"], "e2e", d["e2e"]["value"], "frac", r["frac"], "fp64 util", r["fp64_pipe_utilisation"],
      "cpu", d["cpu_baseline"]["value"] if d.get("cpu_baseline") else None)

# fine-grid capture (optional): 128x128 cells + cell moments
import os  # noqa: E402

if os.path.exists(f"{G}/prof_c5_{tag}.ncu-rep"):
    full(f"{G}/prof_c5_{tag}.ncu-rep", f"profiles/{tag}_c5_fine_grid_full.txt",
         "# C5_N=3000000 ncu --set full --clock-control none --import-source on "
         "-k regex:'k_density_binned|k_cell_moments|k_scan_wide' -c 3 over tools/c5_probe.py\n"
         "# fine-grid path: 3 M + 3 M points, R = 1024, 128x128 cells, cells inside the support summed from moments\n")

# DRAM traffic of the dominant kernel per launch (bench.py's roofline.traffic reads this file)
out = subprocess.run(["ncu", "-i", f"{G}/prof_binned_{tag}.ncu-rep", "--page", "raw", "--csv"],
                     capture_output=True, text=True).stdout
rows = list(csv.reader(out.splitlines()))
hdr, units, first = rows[0], rows[1], rows[2]


def _bytes(key):
    v = float(first[hdr.index(key)].replace(",", ""))
    u = units[hdr.index(key)].lower()
    return v * {"byte": 1.0, "kbyte": 1e3, "mbyte": 1e6, "gbyte": 1e9}.get(u, 1.0)


json.dump({"k_density_binned": {"dram_bytes_read": _bytes("dram__bytes_read.sum"),
                                "dram_bytes_write": _bytes("dram__bytes_write.sum"), "evals_per_launch": 64,
                                "source": f"profiles/{tag}_density_binned_full.txt (ncu --set full, launch 0)"}},
          open("profiles/traffic.json", "w"), indent=1)
