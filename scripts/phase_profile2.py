#!/usr/bin/env python
"""Per-phase attribution of an ncu report of step2_kernel (source page, SASS) joined with nvdisasm's line info
of the in-tree library: instructions, stall samples, lanes, shared-memory / global wavefronts per phase.

    python scripts/phase_profile2.py gpurun_out/prof_x.ncu-rep [kernel-substring] [envs]
"""
import collections, csv, pathlib, re, subprocess, sys, tempfile

rep = sys.argv[1]
pat = sys.argv[2] if len(sys.argv) > 2 else "step2_kernelIfLi8ELi384"
envs = float(sys.argv[3]) if len(sys.argv) > 3 else 0
root = pathlib.Path(__file__).resolve().parent.parent
fname = "b200sim_step2.cuh"
src = (root / "jaxsim_b200/csrc" / fname).read_text().split("\n")
kstart = next(i for i, l in enumerate(src) if "__global__" in l and "step2_kernel" in l) + 1
bounds = [(kstart, "prologue")]
for i, l in enumerate(src[kstart:], start=kstart + 1):
    m = re.match(r"\s*// =+ (.*)$", l)
    if m:
        bounds.append((i, m.group(1).strip()))
bounds.sort()

def phase_of(line):
    name = bounds[0][1]
    for b, n in bounds:
        if line >= b:
            name = n
    return name

with tempfile.TemporaryDirectory() as td:
    subprocess.run(["cuobjdump", "-xelf", "all", str(root / "jaxsim_b200/csrc/libb200sim.so")], cwd=td, check=True, capture_output=True)
    sass = []
    for cubin in pathlib.Path(td).glob("*.cubin"):
        sass += subprocess.run(["nvdisasm", "-g", "-c", str(cubin)], capture_output=True, text=True).stdout.split("\n")
start = next(i for i, l in enumerate(sass) if l.startswith(".text.") and pat in l)
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
def I(r, k):
    try:
        return int(r[ix[k]])
    except Exception:
        return 0
for (ins, srcl), r in zip(seq, data):
    if srcl and srcl[0] == fname and srcl[1] >= kstart:
        phase = phase_of(srcl[1])
    a = agg.setdefault(phase, [0] * 8)
    a[0] += I(r, "Instructions Executed"); a[1] += I(r, "# Samples"); a[2] += 1
    a[3] += I(r, "Thread Instructions Executed")
    a[4] += I(r, "L1 Wavefronts Shared"); a[5] += I(r, "L1 Wavefronts Shared Excessive")
    a[6] += I(r, "L1 Tag Requests Global"); a[7] += I(r, "L2 Theoretical Sectors Global")
ti, ts = sum(v[0] for v in agg.values()), sum(v[1] for v in agg.values())
tw = sum(v[4] for v in agg.values())
print(f"{len(seq)} SASS instructions, {ti} warp instructions executed, {ts} samples, {tw} shared wavefronts" + (f" ({tw/envs:.0f}/env, {ti/envs:.0f} warp-inst/env)" if envs else ""))
print(f"{'phase':58s} {'inst%':>6s} {'smpl%':>6s} {'static':>6s} {'lanes':>5s} {'shwf%':>6s} {'excess%':>7s} {'gtag%':>6s}")
tg = max(1, sum(v[6] for v in agg.values()))
for name, v in agg.items():
    print(f"{name[:58]:58s} {100*v[0]/ti:6.1f} {100*v[1]/max(ts,1):6.1f} {v[2]:6d} {v[3]/max(v[0],1):5.1f} {100*v[4]/max(tw,1):6.1f} {100*v[5]/max(v[4],1):7.1f} {100*v[6]/tg:6.1f}")
