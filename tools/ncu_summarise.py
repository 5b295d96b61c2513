"""Turn an `ncu --set full` report of the fused kernel into the two committed artefacts:
  profiles/<name>.csv       selected raw metrics of the launch (name,unit,value)
  profiles/ncu_summary.json per-clip / per-frame figures bench.py reports as roofline.traffic, keyed by bench.source_sha16() so
                            that a summary captured from other kernel sources is never reported
usage: python tools/ncu_summarise.py report.ncu-rep clips name "comment"   (run where the report's sources are checked out)"""
import csv
import io
import json
import os
import subprocess
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import bench  # noqa: E402

rep, clips, name, comment = sys.argv[1], int(sys.argv[2]), sys.argv[3], sys.argv[4]
raw = subprocess.run(["ncu", "-i", rep, "--page", "raw", "--csv"], capture_output=True, text=True, check=True).stdout
rows = list(csv.reader(io.StringIO(raw)))
hdr, units, vals = rows[0], rows[1], rows[2]
d = {h: (u, v) for h, u, v in zip(hdr, units, vals)}
KEEP = ["gpu__time_duration.sum", "dram__bytes_read.sum", "dram__bytes_write.sum",
        "gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed", "lts__t_sector_hit_rate.pct",
        "l1tex__t_sector_hit_rate.pct", "smsp__issue_active.avg.pct_of_peak_sustained_active",
        "sm__warps_active.avg.pct_of_peak_sustained_active", "smsp__inst_executed.sum",
        "l1tex__data_pipe_lsu_wavefronts.avg.pct_of_peak_sustained_elapsed", "l1tex__data_pipe_lsu_wavefronts_mem_shared.sum",
        "sm__inst_executed_pipe_lsu.avg.pct_of_peak_sustained_active", "sm__pipe_fma_cycles_active.avg.pct_of_peak_sustained_active",
        "sm__inst_executed_pipe_alu.avg.pct_of_peak_sustained_active", "sm__inst_executed_pipe_xu.avg.pct_of_peak_sustained_active",
        "sm__inst_executed_pipe_tensor.avg.pct_of_peak_sustained_active", "sm__pipe_tensor_cycles_active.avg.pct_of_peak_sustained_active",
        "launch__registers_per_thread", "launch__grid_size", "launch__block_size", "launch__shared_mem_per_block_dynamic",
        "launch__occupancy_limit_registers", "launch__occupancy_limit_shared_mem", "sm__maximum_warps_per_active_cycle_pct",
        "sm__inst_executed_pipe_fma.avg.pct_of_peak_sustained_active", "l1tex__data_pipe_lsu_wavefronts.sum.pct_of_peak_sustained_elapsed",
        "l1tex__data_bank_conflicts_pipe_lsu_mem_shared.sum", "sm__throughput.avg.pct_of_peak_sustained_elapsed",
        "lts__t_bytes.sum", "lts__t_sectors_srcunit_tex_op_read.sum", "lts__t_sectors_srcunit_tex_op_write.sum"]
KEEP += sorted(k for k in d if k.startswith("smsp__average_warps_issue_stalled") and k.endswith("_per_issue_active.ratio"))
KEEP += sorted(k for k in d if k.startswith("smsp__average_warp_latency_issue_stalled") and k.endswith(".ratio"))
out = os.path.join(ROOT, "profiles", name + ".csv")
with open(out, "w") as fh:
    fh.write(f"# {comment}\n")
    for k in KEEP:
        if k in d:
            fh.write(f"{k},{d[k][0]},{d[k][1]}\n")


def val(k, scale=1.0):
    u, v = d[k]
    v = float(v.replace(",", ""))
    mult = {"Gbyte": 1e9, "Mbyte": 1e6, "Kbyte": 1e3, "byte": 1.0, "Tbyte": 1e12}.get(u, 1.0)
    return v * mult * scale


frames = clips * 130
s = {"source": f"profiles/{name}.csv ({comment})", "src_sha16": bench.source_sha16(),
     "dram_bytes_per_clip": (val("dram__bytes_read.sum") + val("dram__bytes_write.sum")) / clips,
     "dram_read_bytes_per_clip": val("dram__bytes_read.sum") / clips,
     "dram_write_bytes_per_clip": val("dram__bytes_write.sum") / clips,
     "issue_active_pct": val("smsp__issue_active.avg.pct_of_peak_sustained_active"),
     "warp_inst_per_frame": val("smsp__inst_executed.sum") / frames,
     "smem_wavefronts_per_frame": val("l1tex__data_pipe_lsu_wavefronts_mem_shared.sum") / frames,
     "lsu_data_pipe_pct": val("l1tex__data_pipe_lsu_wavefronts.avg.pct_of_peak_sustained_elapsed"),
     "l2_hit_pct": val("lts__t_sector_hit_rate.pct"), "kernel_ms": val("gpu__time_duration.sum") * (1.0 if d["gpu__time_duration.sum"][0] == "ms" else 1e-3 if d["gpu__time_duration.sum"][0] == "us" else 1e3)}
for k in ("sm__inst_executed_pipe_tensor.avg.pct_of_peak_sustained_active", "sm__pipe_tensor_cycles_active.avg.pct_of_peak_sustained_active"):
    if k in d:
        s["tensor_pipe_pct"] = val(k)
        break
json.dump(s, open(os.path.join(ROOT, "profiles", "ncu_summary.json"), "w"), indent=1)
print(json.dumps(s, indent=1))
