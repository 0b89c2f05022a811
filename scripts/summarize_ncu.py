#!/usr/bin/env python
"""Summarise an ncu report (.ncu-rep) into profiles/<name>.md (+ traffic json).

    python scripts/summarize_ncu.py gpurun_out/prof_step.ncu-rep profiles/r01_step_kernel \
        --key icub_like_f32_B4096
"""
import csv
import io
import json
import pathlib
import subprocess
import sys

KEEP = [
    "gpu__time_duration.sum", "dram__bytes_read.sum", "dram__bytes_write.sum", "gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed",
    "dram__throughput.avg.pct_of_peak_sustained_elapsed", "sm__throughput.avg.pct_of_peak_sustained_elapsed",
    "sm__inst_executed.sum", "smsp__inst_executed.sum", "sm__warps_active.avg.pct_of_peak_sustained_active",
    "smsp__issue_active.avg.pct_of_peak_sustained_active", "smsp__inst_executed.avg.per_cycle_active",
    "sm__cycles_elapsed.avg", "sm__cycles_active.avg", "launch__registers_per_thread", "launch__shared_mem_per_block_dynamic",
    "launch__grid_size", "launch__block_size", "launch__occupancy_limit_shared_mem", "launch__occupancy_limit_registers",
    "launch__waves_per_multiprocessor", "l1tex__data_bank_conflicts_pipe_lsu_mem_shared.sum",
    "l1tex__data_pipe_lsu_wavefronts_mem_shared.sum", "smsp__pcsamp_warps_issue_stalled_long_scoreboard",
    "lts__t_sector_hit_rate.pct", "l1tex__t_sector_hit_rate.pct", "sm__pipe_fma_cycles_active.avg.pct_of_peak_sustained_active",
    "sm__inst_executed_pipe_lsu.sum", "sm__inst_executed_pipe_fma.sum", "sm__inst_executed_pipe_alu.sum",
    "sm__sass_thread_inst_executed_op_ffma_pred_on.sum", "sm__sass_thread_inst_executed_op_fadd_pred_on.sum",
    "sm__sass_thread_inst_executed_op_fmul_pred_on.sum", "smsp__thread_inst_executed_per_inst_executed.ratio",
]


def main():
    rep, out = sys.argv[1], pathlib.Path(sys.argv[2])
    key = sys.argv[sys.argv.index("--key") + 1] if "--key" in sys.argv else None
    raw = subprocess.run(["ncu", "-i", rep, "--page", "raw", "--csv"], capture_output=True, text=True, check=True).stdout
    rows = list(csv.reader(io.StringIO(raw)))
    hdr, units, data = rows[0], rows[1], rows[2:]
    idx = {h: i for i, h in enumerate(hdr)}
    lines = [f"# ncu summary of `{pathlib.Path(rep).name}`", "",
             "Captured with `ncu --set full --clock-control none --import-source on` (replayed, cold caches:",
             "durations here are NOT bench values).", ""]
    traffic = []
    for r in data:
        name = r[idx["Kernel Name"]]
        lines += [f"## {name[:100]}  (id {r[idx['ID']]})", "", "| metric | value | unit |", "|---|---|---|"]
        for k in hdr:
            if k in KEEP or k.startswith("smsp__pcsamp_warps_issue_stalled") or k.startswith("smsp__average_warps_issue_stalled") or k.startswith("smsp__average_warp"):
                v = r[idx[k]]
                if v in ("", "0", "0.0") and k not in KEEP:
                    continue
                lines.append(f"| {k} | {v} | {units[idx[k]]} |")
        lines.append("")
        try:
            def f(k):
                return float(r[idx[k]].replace(",", ""))
            rd, wr = f("dram__bytes_read.sum"), f("dram__bytes_write.sum")
            mult = {"byte": 1, "Kbyte": 1e3, "Mbyte": 1e6, "Gbyte": 1e9}
            rd *= mult.get(units[idx["dram__bytes_read.sum"]], 1)
            wr *= mult.get(units[idx["dram__bytes_write.sum"]], 1)
            traffic.append(rd + wr)
            lines += [f"DRAM traffic this launch: read {rd/1e6:.2f} MB + write {wr/1e6:.2f} MB = {(rd+wr)/1e6:.2f} MB", ""]
        except Exception as e:  # noqa: BLE001
            lines += [f"(traffic parse failed: {e})", ""]
    out.with_suffix(".md").write_text("\n".join(lines))
    if key and traffic:
        tp = out.parent / "traffic_per_launch.json"
        d = json.loads(tp.read_text()) if tp.exists() else {}
        d[key] = sum(traffic) / len(traffic)
        tp.write_text(json.dumps(d, indent=1))
    print("\n".join(lines[:80]))


if __name__ == "__main__":
    main()
