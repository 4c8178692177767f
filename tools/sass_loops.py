"""Trimmed SASS of the candidate loops of the two culled density kernels, with per-opcode counts, from the built
libkdejs.so (cuobjdump; no GPU needed):  python tools/sass_loops.py > profiles/<tag>_sass_density_loops.txt"""
import collections
import os
import re
import subprocess
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
SO = os.path.join(ROOT, "poppunk-mod_b200", "libkdejs.so")
sass = subprocess.run(["cuobjdump", "-sass", SO], capture_output=True, text=True, check=True).stdout
funcs, cur = {}, None
for ln in sass.splitlines():
    m = re.search(r"Function : (\S+)", ln)
    if m:
        cur = m.group(1)
        funcs[cur] = []
    elif cur and re.match(r"\s+/\*[0-9a-f]{4}\*/", ln):
        funcs[cur].append(re.sub(r"\s*/\* 0x[0-9a-f]+ \*/\s*$", "", ln).rstrip())
addr = lambda l: int(re.search(r"/\*([0-9a-f]{4})\*/", l).group(1), 16)


def opcode(l):
    t = l.split("*/", 1)[1].split()
    op = t[1] if t[0].startswith("@") else t[0]
    return op.rstrip(";")


def loops(lines, marker, per_candidate):
    """innermost backward-branch loops whose body holds `per_candidate` marker instructions per candidate"""
    A = [addr(l) for l in lines]
    out = []
    for i, l in enumerate(lines):
        m = re.search(r"BRA\s+(0x[0-9a-f]+)", l)
        if not m:
            continue
        tgt = int(m.group(1), 16)
        if tgt < A[i] and tgt in A:
            body = lines[A.index(tgt): i + 1]
            n = sum(marker in opcode(b) for b in body)
            if per_candidate <= n <= 4 * per_candidate:
                out.append((tgt, A[i], body, n // per_candidate))
    return out


print("# Candidate loops of the culled density kernels (sm_100a SASS of the shipped libkdejs.so; tools/sass_loops.py)")
print("# A lane tests one candidate against the 16 nodes of a 4x4 tile; loops are unrolled over 1-4 candidates per lane")
print("# (fp32, warp-per-tile phase: two trips of two candidates each).\n")
for name, lines in sorted(funcs.items()):
    if "k_density_binned32ILb0" in name:
        title, marker, per = "k_density_binned32<false>  (fp32 terms: FFMA.SAT + packed FADD2)", "FFMA.SAT", 16
    elif "k_density_binnedILb0" in name:
        title, marker, per = "k_density_binned<false>  (fp64: DFMA / DADD, clamp = VIMNMX on the high word)", "VIMNMX", 16
    else:
        continue
    print(f"## {title}\n## {name}: {len(lines)} instructions, only sm_100a code in the binary")
    hist_all = collections.Counter(opcode(l).split(".")[0] for l in lines)
    print("## whole kernel, static:", ", ".join(f"{k} {v}" for k, v in hist_all.most_common(14)))
    printed = 0
    for tgt, end, body, trips in loops(lines, marker, per):
        if len(body) > 330:             # outer loops (whole tiles): not candidate loops
            continue
        fam = collections.Counter(opcode(b).split(".")[0] for b in body)
        print(f"\n### loop 0x{tgt:04x}..0x{end:04x}: {len(body)} instructions for {trips} candidate(s) per lane = {len(body) / trips:.1f} per candidate "
              "(static count; run-switch and fold code inside the body executes once per run / per fold block)")
        print("### by opcode family:", ", ".join(f"{k} {v}" for k, v in fam.most_common()))
        if len(body) <= (80 if "binned32" not in name else 280) and printed < 3:
            printed += 1
            for b in body:
                print(b)
    print()
