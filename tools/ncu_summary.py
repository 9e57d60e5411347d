#!/usr/bin/env python
"""Turn an `ncu --set full` report into the small JSON bench.py attaches to its roofline
(profiles/ncu_dominant_kernel.json) and keep the raw CSV page beside it.
    python tools/ncu_summary.py gpurun_out/prof.ncu-rep profiles/<name>_raw.csv <variant> "<command that was profiled>" """
import csv, json, subprocess, sys

rep, raw_out, variant, source = sys.argv[1:5]
raw = subprocess.run(["ncu", "-i", rep, "--page", "raw", "--csv"], stdout=subprocess.PIPE, text=True, check=True).stdout
open(raw_out, "w").write(raw)
rows = list(csv.reader(raw.splitlines()))
h = rows[0]
pick = int(sys.argv[5]) if len(sys.argv) > 5 else 0   # which captured launch (0 = first)
v = rows[2 + pick]
col = {k: i for i, k in enumerate(h)}


def f(name, default=None):
    try:
        return float(v[col[name]].replace(",", ""))
    except Exception:
        return default


import hashlib, os
_h = hashlib.sha256()
for _f in ("gkm_index.cu", "gkm_index.h", "gkm_index_dev.h"):   # bench.py:kernel_source_hash -- the capture belongs to this source
    _h.update(open(os.path.join(os.path.dirname(os.path.dirname(os.path.abspath(__file__))), "gkmqc_b200", "csrc", _f), "rb").read())
out = {
    "variant": variant,
    "source_sha256": _h.hexdigest()[:16],
    "source": source + "; " + raw_out,
    "kernel": v[col["Kernel Name"]],
    "grid": v[col["Grid Size"]] if "Grid Size" in col else None,
    "block": v[col["Block Size"]] if "Block Size" in col else None,
    "duration_ms": f("gpu__time_duration.sum"),
    "registers_per_thread": f("launch__registers_per_thread"),
    "dyn_smem_bytes_per_block": f("launch__shared_mem_per_block_dynamic"),
    "dram_bytes_read": f("dram__bytes_read.sum"),
    "dram_bytes_write": f("dram__bytes_write.sum"),
    "l1tex_throughput_pct": f("l1tex__throughput.avg.pct_of_peak_sustained_elapsed"),
    "l1tex_lsu_wavefronts_pct": f("l1tex__data_pipe_lsu_wavefronts.avg.pct_of_peak_sustained_elapsed"),
    "lts_throughput_pct": f("lts__throughput.avg.pct_of_peak_sustained_elapsed"),
    "lts_sector_hit_rate_pct": f("lts__t_sector_hit_rate.pct"),
    "l1_global_load_sectors": f("l1tex__t_sectors_pipe_lsu_mem_global_op_ld.sum"),
    "l2_read_sectors_from_l1": f("lts__t_sectors_srcunit_tex_op_read.sum"),
    "shared_atomic_instructions": f("smsp__inst_executed_op_shared_atom.sum"),
    "alu_pipe_pct_of_peak": f("sm__inst_executed_pipe_alu.avg.pct_of_peak_sustained_active"),
    "lsu_pipe_pct_of_peak": f("sm__inst_executed_pipe_lsu.avg.pct_of_peak_sustained_active"),
    "issue_slots_busy_pct": f("smsp__issue_active.avg.pct_of_peak_sustained_active"),
    "warps_active_pct": f("sm__warps_active.avg.pct_of_peak_sustained_active"),
    "warp_instructions": f("smsp__inst_executed.sum"),
}
for k in ("dram_bytes_read", "dram_bytes_write"):
    # the raw page reports these in scaled units (Mbyte / Kbyte ...): convert through the unit row
    unit = rows[1][col["dram__bytes_read.sum" if k.endswith("read") else "dram__bytes_write.sum"]].lower()
    scale = {"byte": 1.0, "kbyte": 1e3, "mbyte": 1e6, "gbyte": 1e9}.get(unit, 1.0)
    if out[k] is not None:
        out[k] *= scale
out["dram_bytes_per_launch"] = (out["dram_bytes_read"] or 0) + (out["dram_bytes_write"] or 0)
su = rows[1][col["launch__shared_mem_per_block_dynamic"]].lower() if "launch__shared_mem_per_block_dynamic" in col else ""
if out["dyn_smem_bytes_per_block"] is not None:
    out["dyn_smem_bytes_per_block"] *= 1e3 if su.startswith("kbyte") else 1e6 if su.startswith("mbyte") else 1.0
ms_unit = rows[1][col["gpu__time_duration.sum"]].lower()
out["duration_ms"] = out["duration_ms"] * {"ns": 1e-6, "us": 1e-3, "ms": 1.0, "s": 1e3}.get(ms_unit, 1.0)
print(json.dumps(out, indent=1))
