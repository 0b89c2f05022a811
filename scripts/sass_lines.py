"""Static SASS statistics of one kernel per source line (needs -lineinfo): instruction count and scalar
shared-memory accesses.  Usage: python scripts/sass_lines.py <cubin> <kernel-substring> [top]"""
import collections
import re
import subprocess
import sys

cubin, pat = sys.argv[1], sys.argv[2]
top = int(sys.argv[3]) if len(sys.argv) > 3 else 40
dis = subprocess.run(["nvdisasm", "-g", cubin], capture_output=True, text=True).stdout
fn = None
cur = None
tot = collections.Counter()
sc = collections.Counter()
ops = collections.Counter()
for line in dis.splitlines():
    m = re.match(r"\s*\.section\s+\.text\.(\S+?),", line)
    if m:
        fn = m.group(1)
        continue
    m = re.search(r'//## File "([^"]+)", line (\d+)', line)
    if m:
        cur = (m.group(1).split("/")[-1], int(m.group(2)))
        continue
    if fn and pat in fn:
        m = re.match(r"\s+/\*[0-9a-f]+\*/\s+(?:@!?U?P\d+\s+)?([A-Z0-9_.]+)", line)
        if m:
            op = m.group(1)
            tot[cur] += 1
            ops[op.split(".")[0]] += 1
            if op in ("LDS", "STS"):
                sc[cur] += 1
print("instructions:", sum(tot.values()))
print("scalar LDS/STS by line:", sorted(sc.items(), key=lambda x: -x[1])[:top])
print("instructions by line:")
for k, v in sorted(tot.items(), key=lambda x: -x[1])[:top]:
    print("  ", k, v)
print("opcodes:", ops.most_common(25))
