#!/usr/bin/env python
"""Summarise an `ncu --set full` report (read here, on the CPU box: `ncu -i ... --page raw --csv`) into the small JSON /
markdown files committed under profiles/ -- the numbers DESIGN.md and bench.py's `roofline.traffic` cite.

    python tools/ncu_summary.py gpurun_out/prof_fwd.ncu-rep profiles/r2_prof_render_fwd_summary.json [--traffic-key render_forward_kernel]

With --traffic-key the kernel's DRAM bytes per launch are also written into profiles/render_forward_traffic.json, stamped with
the sha256 of the compositing sources (bench.py refuses a stale figure)."""
import csv
import io
import json
import os
import subprocess
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
KEEP = [
    "gpu__time_duration.sum", "smsp__inst_executed.sum", "sm__inst_executed.avg.per_cycle_elapsed", "sm__inst_executed.avg.per_cycle_active",
    "smsp__issue_active.avg.pct_of_peak_sustained_active", "sm__warps_active.avg.pct_of_peak_sustained_active",
    "launch__registers_per_thread", "launch__grid_size", "launch__block_size", "launch__shared_mem_per_block_static",
    "launch__shared_mem_per_block_dynamic", "launch__occupancy_limit_registers", "launch__occupancy_limit_shared_mem",
    "launch__occupancy_limit_warps", "launch__waves_per_multiprocessor", "dram__bytes_read.sum", "dram__bytes_write.sum",
    "dram__throughput.avg.pct_of_peak_sustained_elapsed", "gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed",
    "lts__t_sectors_srcunit_tex_op_red.sum", "lts__t_bytes.sum", "l1tex__data_bank_conflicts_pipe_lsu_mem_shared.sum",
    "sm__cycles_active.avg", "sm__cycles_elapsed.avg", "sm__pipe_fma_cycles_active.avg.pct_of_peak_sustained_active",
    "sm__pipe_alu_cycles_active.avg.pct_of_peak_sustained_active", "sm__inst_executed_pipe_xu.sum", "sm__inst_executed_pipe_lsu.sum",
    "sm__inst_executed_pipe_fma.sum", "sm__inst_executed_pipe_alu.sum", "sm__pipe_tensor_cycles_active.avg.pct_of_peak_sustained_active",
]
STALLS = "smsp__pcsamp_warps_issue_stalled_"


def main():
    rep, out = sys.argv[1], sys.argv[2]
    tkey = sys.argv[sys.argv.index("--traffic-key") + 1] if "--traffic-key" in sys.argv else None
    raw = subprocess.run(["ncu", "-i", rep, "--page", "raw", "--csv"], capture_output=True, text=True, check=True).stdout
    rows = list(csv.reader(io.StringIO(raw)))
    hdr, units, vals = rows[0], rows[1], rows[-1]
    col = {h: i for i, h in enumerate(hdr)}
    num = lambda s: float(s.replace(",", "")) if s not in ("", "n/a") else None
    summ = {"report": os.path.basename(rep), "kernel": vals[col["Kernel Name"]], "metrics": {}, "stall_samples": {}}
    for k in KEEP:
        if k in col:
            summ["metrics"][k] = {"value": num(vals[col[k]]), "unit": units[col[k]]}
    for h in hdr:
        if h.startswith(STALLS) and not h.endswith("_not_issued"):
            v = num(vals[col[h]])
            if v:
                summ["stall_samples"][h[len(STALLS):]] = v
    m = summ["metrics"]
    g = lambda k: m.get(k, {}).get("value")
    scale = {"byte": 1.0, "Kbyte": 1e3, "Mbyte": 1e6, "Gbyte": 1e9}
    dram = sum((g(k) or 0.0) * scale.get(m[k]["unit"], 1.0) for k in ("dram__bytes_read.sum", "dram__bytes_write.sum") if k in m)
    summ["derived"] = {"dram_bytes_per_launch": dram,
                       "l2_red_bytes": (g("lts__t_sectors_srcunit_tex_op_red.sum") or 0.0) * 32.0,
                       "issue_slots_used_of_elapsed": (g("sm__inst_executed.avg.per_cycle_elapsed") or 0.0) / 4.0}
    with open(out, "w") as f:
        json.dump(summ, f, indent=1)
    print(json.dumps({"kernel": summ["kernel"][:60], "us": g("gpu__time_duration.sum"), **summ["derived"]}))
    if tkey:
        sys.path.insert(0, ROOT)
        import importlib.util
        spec = importlib.util.spec_from_file_location("bench", os.path.join(ROOT, "bench.py"))
        bench = importlib.util.module_from_spec(spec)
        spec.loader.exec_module(bench)
        tp = os.path.join(ROOT, "profiles", "render_forward_traffic.json")
        try:
            j = json.load(open(tp))
            if "dram_bytes_per_launch" in j:          # r1 flat layout -> per-kernel layout
                j = {}
        except Exception:
            j = {}
        j[tkey] = {"dram_bytes_per_launch": dram, "source_sha256": bench._source_hash(tkey), "from": os.path.basename(out),
                   "how": "ncu --set full --clock-control none, dram__bytes_read.sum + dram__bytes_write.sum, one launch, C2 scene seed 1314"}
        with open(tp, "w") as f:
            json.dump(j, f, indent=1)


if __name__ == "__main__":
    main()
