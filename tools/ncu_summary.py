"""Summarises an `ncu --set full` report (read here with `ncu -i ... --page raw --csv`) into the JSON files bench.py / DESIGN.md cite:
per launch duration, DRAM bytes, achieved occupancy, issue-slot utilisation, tensor-pipe activity, shared-memory bank conflicts.
usage: python tools/ncu_summary.py <report.ncu-rep> <out.json> [--phase-traffic]"""
import csv
import json
import subprocess
import sys

rep, out = sys.argv[1], sys.argv[2]
raw = subprocess.run(["ncu", "-i", rep, "--page", "raw", "--csv"], capture_output=True, text=True).stdout
rows = list(csv.reader(raw.splitlines()))
h = rows[0]
want = {"gpu__time_duration.sum": "duration_us", "dram__bytes_read.sum": "dram_read_bytes", "dram__bytes_write.sum": "dram_write_bytes",
        "sm__warps_active.avg.pct_of_peak_sustained_active": "achieved_occupancy_pct", "smsp__issue_active.avg.pct_of_peak_sustained_active": "issue_active_pct",
        "sm__pipe_tensor_cycles_active.avg.pct_of_peak_sustained_active": "tensor_pipe_active_pct",
        "sm__inst_executed_pipe_tensor.sum": "tensor_instructions", "l1tex__data_bank_conflicts_pipe_lsu_mem_shared.sum": "smem_bank_conflicts",
        "smsp__inst_executed.sum": "warp_instructions", "launch__grid_size": "grid", "launch__block_size": "block", "launch__registers_per_thread": "regs",
        "sm__throughput.avg.pct_of_peak_sustained_elapsed": "sm_throughput_pct", "lts__t_bytes.sum": "l2_bytes"}
units = rows[1]
launches = []
for r in rows[2:]:
    if len(r) < len(h):
        continue
    d = {"kernel": r[h.index("Kernel Name")].split("(")[0]}
    for m, k in want.items():
        if m in h:
            try:
                v = float(r[h.index(m)].replace(",", ""))
            except ValueError:
                continue
            u = units[h.index(m)]
            if k == "duration_us":
                v = v / 1e3 if u in ("ns", "nsecond") else (v if u in ("us", "usecond") else v * 1e3 if u in ("ms", "msecond") else v)
            if k.endswith("_bytes"):
                v *= {"byte": 1, "Kbyte": 1e3, "Mbyte": 1e6, "Gbyte": 1e9}.get(u, 1)
            d[k] = v
    launches.append(d)
res = {"report": rep, "launches": launches}
if "--phase-traffic" in sys.argv:
    tot = sum(l.get("dram_read_bytes", 0) + l.get("dram_write_bytes", 0) for l in launches)
    res["critic_phase_dram_bytes"] = tot
    res["how"] = ("ncu --set full --cache-control none --clock-control none of the five launches of one critic phase (k_gather_mb, k_dense_tc x2, k_tail, "
                  "k_dense_tc_t<1>) in a warm run (L2 NOT flushed between kernel replays); dram__bytes_read.sum + dram__bytes_write.sum summed")
json.dump(res, open(out, "w"), indent=1)
for l in launches:
    print({k: (round(v, 2) if isinstance(v, float) else v) for k, v in l.items()})
if "critic_phase_dram_bytes" in res:
    print("critic phase DRAM bytes:", res["critic_phase_dram_bytes"])
