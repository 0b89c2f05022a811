#!/usr/bin/env python
"""Attribute the per-instruction counters of an ncu report (source page, SASS) to the phases
of step_kernel by joining them with nvdisasm's line info of the same cubin.

    python scripts/phase_profile.py gpurun_out/prof_step.ncu-rep [mangled-kernel-prefix]

Inlined helpers (ldn/stn/mat3_*) carry the helper's line number, so an instruction is
attributed to the phase of the most recent instruction whose line lies in the kernel body.
"""
import collections, csv, pathlib, re, subprocess, sys, tempfile

rep = sys.argv[1]
prefix = sys.argv[2] if len(sys.argv) > 2 else ".text._ZN7b200sim11step_kernelIfLi8EEE"
root = pathlib.Path(__file__).resolve().parent.parent
src = (root / "jaxsim_b200/csrc/b200sim_kernels.cuh").read_text().split("\n")
# phase boundaries = the "// ====...= name" banner comments inside the kernel body
kstart = next(i for i, l in enumerate(src) if "__global__" in l and "step_kernel" in l) + 1
bounds = [(kstart, "prologue")]
for i, l in enumerate(src[kstart:], start=kstart + 1):
    m = re.match(r"\s*// =+ (.*)$", l)
    if m:
        bounds.append((i, m.group(1).strip()))
    m = re.match(r"\s*auto (\w+) = \[&\]", l)
    if m:
        bounds.append((i, "lambda " + m.group(1)))
bounds.sort()

def phase_of(line):
    name = bounds[0][1]
    for b, n in bounds:
        if line >= b:
            name = n
    return name

with tempfile.TemporaryDirectory() as td:
    subprocess.run(["cuobjdump", "-xelf", "all", str(root / "jaxsim_b200/csrc/libb200sim.so")], cwd=td, check=True, capture_output=True)
    cubin = next(pathlib.Path(td).glob("*.cubin"))
    sass = subprocess.run(["nvdisasm", "-g", "-c", str(cubin)], capture_output=True, text=True).stdout.split("\n")
start = next(i for i, l in enumerate(sass) if l.startswith(prefix))
end = next(i for i in range(start + 1, len(sass)) if sass[i].startswith("//--------------------- .text."))
cur, seq = None, []
for l in sass[start:end]:
    m = re.search(r'//## File "([^"]+)", line (\d+)', l)
    if m:
        cur = (m.group(1).split("/")[-1], int(m.group(2)))
    m2 = re.match(r"\s+/\*([0-9a-f]{4,})\*/\s+(.*?);", l)
    if m2:
        seq.append((m2.group(2), cur))
raw = subprocess.run(["ncu", "-i", rep, "--page", "source", "--csv", "--print-source", "sass"], capture_output=True, text=True).stdout
rows = list(csv.reader(raw.split("\n")))
hdr = rows[1]
ix = {h: k for k, h in enumerate(hdr)}
data = [r for r in rows[2:] if len(r) == len(hdr)][:len(seq)]
assert len(data) == len(seq), (len(data), len(seq))
agg = collections.OrderedDict()
phase = "prologue"
for (ins, srcl), r in zip(seq, data):
    if srcl and srcl[0] == "b200sim_kernels.cuh" and srcl[1] >= kstart:
        phase = phase_of(srcl[1])
    a = agg.setdefault(phase, [0, 0, 0, 0])
    a[0] += int(r[ix["Instructions Executed"]]); a[1] += int(r[ix["# Samples"]]); a[2] += 1
    a[3] += int(r[ix["Thread Instructions Executed"]])
ti, ts = sum(v[0] for v in agg.values()), sum(v[1] for v in agg.values())
print(f"{len(seq)} SASS instructions, {ti} warp instructions executed, {ts} samples")
print(f"{'phase':58s} {'inst%':>6s} {'smpl%':>6s} {'static':>6s} {'lanes':>5s}")
for name, v in agg.items():
    print(f"{name[:58]:58s} {100*v[0]/ti:6.1f} {100*v[1]/max(ts,1):6.1f} {v[2]:6d} {v[3]/max(v[0],1):5.1f}")
