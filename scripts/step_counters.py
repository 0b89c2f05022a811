#!/usr/bin/env python
"""Turn an ncu counter capture of consecutive step launches into profiles/r02_step_counters.json (read by bench.py).

Capture (on the GPU box; single pass, no kernel replay, caches left alone, so the writes one launch leaves in L2 are
counted when they drain during the following launches -- a steady-state figure):

    ncu --profile-from-start off --cache-control none --clock-control none -k regex:step -c 40 --csv \
        --metrics dram__bytes_read.sum,dram__bytes_write.sum,smsp__sass_thread_inst_executed_op_ffma_pred_on.sum,\
smsp__sass_thread_inst_executed_op_fadd_pred_on.sum,smsp__sass_thread_inst_executed_op_fmul_pred_on.sum,gpu__time_duration.sum \
        --log-file gpurun_out/step_counters.csv python bench.py --steps 40 --warmup 5 --no-cpu-baseline --no-graph --no-extras --profile

    python scripts/step_counters.py gpurun_out/step_counters.csv icub_like_f32_B4096 4096
"""
import csv
import json
import pathlib
import sys

src, key, batch = sys.argv[1], sys.argv[2], int(sys.argv[3])
rows = [r for r in csv.reader(open(src)) if len(r) > 10]
hdr = rows[0]
ix = {h: i for i, h in enumerate(hdr)}
acc = {}
launches = set()
for r in rows[1:]:
    name, val, unit = r[ix["Metric Name"]], r[ix["Metric Value"]].replace(",", ""), r[ix["Metric Unit"]]
    v = float(val)
    if "byte" in unit:
        v *= {"byte": 1, "Kbyte": 1e3, "Mbyte": 1e6, "Gbyte": 1e9}[unit]
    acc[name] = acc.get(name, 0.0) + v
    launches.add(r[ix["ID"]])
n = len(launches)
rd, wr = acc["dram__bytes_read.sum"] / n, acc["dram__bytes_write.sum"] / n
flops = (2 * acc["smsp__sass_thread_inst_executed_op_ffma_pred_on.sum"] + acc["smsp__sass_thread_inst_executed_op_fadd_pred_on.sum"]
         + acc["smsp__sass_thread_inst_executed_op_fmul_pred_on.sum"]) / n / batch
out = pathlib.Path(__file__).resolve().parent.parent / "profiles" / "r02_step_counters.json"
d = json.loads(out.read_text()) if out.exists() else {}
d[key] = {"launches": n, "dram_bytes_read_per_launch": rd, "dram_bytes_write_per_launch": wr, "dram_bytes_per_launch": rd + wr,
          "flops_per_env_step": flops, "mean_kernel_us_under_ncu": acc.get("gpu__time_duration.sum", 0) / n / 1e3,
          "source": f"ncu counters over {n} consecutive launches ({pathlib.Path(src).name}), scripts/step_counters.py"}
out.write_text(json.dumps(d, indent=1) + "\n")
print(json.dumps(d[key], indent=1))
