"""Turn the gpurun_out/ artefacts of tools/capture_profiles.sh <tag> into tracked summaries under profiles/."""
import csv
import json
import os
import shutil
import subprocess
import sys

sys.path.insert(0, "tools")
from ncu_summary import KEYS, launches  # noqa: E402

tag = sys.argv[1]
G = "gpurun_out"


def raw(rep):
    out = subprocess.run(["ncu", "-i", rep, "--page", "raw", "--csv"], capture_output=True, text=True).stdout
    rows = list(csv.reader(out.splitlines()))
    return rows[0], rows[1], rows[2:]


def full(rep, dst, header, pick=None):
    hdr, units, rows = raw(rep)
    with open(dst, "w") as f:
        f.write(header)
        for idx, r in enumerate(rows):
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


def dram(rep):
    hdr, units, rows = raw(rep)
    first = rows[0]

    def _bytes(key):
        v = float(first[hdr.index(key)].replace(",", ""))
        u = units[hdr.index(key)].lower()
        return v * {"byte": 1.0, "kbyte": 1e3, "mbyte": 1e6, "gbyte": 1e9}.get(u, 1.0)

    return _bytes("dram__bytes_read.sum"), _bytes("dram__bytes_write.sum")


cmd = "bench.py --steps 10 --warmup 3 --no-cpu-baseline --no-extras"
launches(f"{G}/launches_{tag}.csv", f"profiles/{tag}_launches.txt")
full(f"{G}/prof_binned32_{tag}.ncu-rep", f"profiles/{tag}_density_binned32_full.txt",
     f"# ncu --set full --clock-control none --import-source on -k regex:k_density_binned32 -s 10 -c 2 over {cmd}\n"
     "# dominant kernel of the timed step (default algorithm): a wave of 64 evaluations per launch\n")
full(f"{G}/prof_binned_{tag}.ncu-rep", f"profiles/{tag}_density_binned_full.txt",
     f"# ncu --set full --clock-control none --import-source on -k 'regex:^k_density_binned$' -s 10 -c 2 over bench.py --algo binned ...\n"
     "# the fp64 culled kernel (algo=\"binned\"): a wave of 64 evaluations per launch\n")
full(f"{G}/prof_small_{tag}.ncu-rep", f"profiles/{tag}_other_kernels_full.txt",
     f"# ncu --set full --clock-control none, the other kernels of one wave (incl. the fp64 redo launch), over {cmd}\n")
if os.path.exists(f"{G}/prof_c5_{tag}.ncu-rep"):
    full(f"{G}/prof_c5_{tag}.ncu-rep", f"profiles/{tag}_c5_fine_grid_full.txt",
         "# C5_N=3000000 ncu --set full --clock-control none -k 'regex:^k_density_binned$|k_cell_moments|k_scan_wide|k_scale_count' "
         "-s 4 -c 4 over tools/c5_probe.py\n# fine-grid path: 3 M + 3 M points, R = 1024, 128 x 128 cells, moment cells\n")
if os.path.exists(f"{G}/prof_tsv_{tag}.ncu-rep"):
    full(f"{G}/prof_tsv_{tag}.ncu-rep", f"profiles/{tag}_tsv_kernels_full.txt",
         "# ncu --set full --clock-control none --import-source on -k regex:k_tsv -s 6 -c 3 over tools/tsv_probe.py 64\n"
         "# the device table parser: one wave of 16 tables of 100 000 rows (2.84 MB of text each) per launch\n")
for name in (f"bench_{tag}_n1.json", f"bench_{tag}_n1_binned.json", f"bench_{tag}_reference.json", f"fp32_probe_{tag}.log",
             f"band_sort_probe_{tag}.log", f"match_probe_{tag}.log", f"h2d_probe_{tag}.jsonl", f"wave_probe_{tag}.log",
             f"tail_probe_{tag}.log", f"ingest_probe_{tag}.log", f"tsv_probe_{tag}.log"):
    if os.path.exists(f"{G}/{name}"):
        shutil.copy(f"{G}/{name}", f"profiles/{name}")
d = json.load(open(f"{G}/bench_{tag}_n1.json"))
r = d["roofline"]
print("value", d["value"], "e2e", d["e2e"]["value"], "frac", r["frac"], "pipe util", r["pipe_utilisation"],
      "cpu", d["cpu_baseline"]["value"] if d.get("cpu_baseline") else None)

# DRAM traffic of the density kernels per launch (bench.py's roofline.traffic reads this file)
traffic = {}
for key, rep, src in (("k_density_binned32", f"{G}/prof_binned32_{tag}.ncu-rep", f"profiles/{tag}_density_binned32_full.txt"),
                      ("k_density_binned", f"{G}/prof_binned_{tag}.ncu-rep", f"profiles/{tag}_density_binned_full.txt")):
    rd, wr = dram(rep)
    traffic[key] = {"dram_bytes_read": rd, "dram_bytes_write": wr, "evals_per_launch": 64,
                    "source": f"{src} (ncu --set full, launch 0)"}
json.dump(traffic, open("profiles/traffic.json", "w"), indent=1)
subprocess.run([sys.executable, "tools/sass_loops.py"], stdout=open(f"profiles/{tag}_sass_density_loops.txt", "w"), check=True)
