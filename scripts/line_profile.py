#!/usr/bin/env python
"""Per-source-line attribution of an ncu report (source page, SASS) joined with nvdisasm's line info of the
in-tree library: warp instructions executed, stall samples, shared-memory wavefronts per source line.

    python scripts/line_profile.py gpurun_out/prof_x.ncu-rep <kernel-substring> [top]
"""
import collections, csv, pathlib, re, subprocess, sys, tempfile

rep, pat = sys.argv[1], sys.argv[2]
top = int(sys.argv[3]) if len(sys.argv) > 3 else 45
root = pathlib.Path(__file__).resolve().parent.parent
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
def I(r, k):
    try:
        return int(r[ix[k]])
    except Exception:
        return 0
agg = collections.defaultdict(lambda: [0, 0, 0, 0, 0, 0])
stall_cols = [h for h in hdr if h.startswith("stall_") and "Not Issued" not in h]
stalls = collections.defaultdict(collections.Counter)
for (ins, srcl), r in zip(seq, data):
    a = agg[srcl]
    a[0] += I(r, "Instructions Executed"); a[1] += I(r, "# Samples"); a[2] += 1
    a[3] += I(r, "Thread Instructions Executed"); a[4] += I(r, "L1 Wavefronts Shared"); a[5] += I(r, "L1 Wavefronts Shared Excessive")
    for h in stall_cols:
        stalls[srcl][h] += I(r, h)
ti = sum(v[0] for v in agg.values()); ts = sum(v[1] for v in agg.values()); tw = sum(v[4] for v in agg.values())
print(f"{len(seq)} SASS instructions, {ti} warp instructions executed, {ts} samples, {tw} shared wavefronts")
src_cache = {}
def text(srcl):
    if not srcl:
        return ""
    f = next(iter(root.glob("jaxsim_b200/csrc/" + srcl[0])), None)
    if f is None:
        return ""
    if f not in src_cache:
        src_cache[f] = f.read_text().split("\n")
    return src_cache[f][srcl[1] - 1].strip()[:70]
print(f"{'line':28s} {'inst%':>6s} {'smpl%':>6s} {'static':>6s} {'lanes':>5s} {'shwf%':>6s} {'exc%':>5s}  top stalls | source")
for srcl, v in sorted(agg.items(), key=lambda kv: -kv[1][1])[:top]:
    st = ", ".join(f"{k[6:]} {100*c/max(v[1],1):.0f}" for k, c in stalls[srcl].most_common(3))
    name = f"{srcl[0]}:{srcl[1]}" if srcl else "?"
    print(f"{name[-28:]:28s} {100*v[0]/ti:6.1f} {100*v[1]/max(ts,1):6.1f} {v[2]:6d} {v[3]/max(v[0],1):5.1f} {100*v[4]/max(tw,1):6.1f} {100*v[5]/max(v[4],1):5.0f}  {st} | {text(srcl)}")
