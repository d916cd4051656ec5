"""Headline counters of every kernel in an .ncu-rep (read with `ncu -i ... --page raw --csv`) as JSON.

usage: python tools/ncu_summary.py report.ncu-rep [candidates_per_launch] > summary.json
"""
import csv
import io
import json
import subprocess
import sys

WANT = {
    "gpu__time_duration.sum": "time",
    "dram__bytes_read.sum": "dram_read",
    "dram__bytes_write.sum": "dram_write",
    "dram__throughput.avg.pct_of_peak_sustained_elapsed": "dram_pct_of_peak",
    "sm__throughput.avg.pct_of_peak_sustained_elapsed": "sm_pct_of_peak",
    "smsp__issue_active.avg.pct_of_peak_sustained_active": "issue_active_pct",
    "sm__warps_active.avg.pct_of_peak_sustained_active": "warps_active_pct",
    "smsp__inst_executed.sum": "warp_instructions",
    "sm__inst_executed_pipe_fp64.avg.pct_of_peak_sustained_active": "fp64_pipe_pct",
    "SM_C.TriageCompute.smsp__pipe_tensor_subpipe_dmma_cycles_active.avg": "dmma_cycles_active_per_smsp",
    "TPC.TriageCompute.sm__pipe_tensor_cycles_active_realtime.avg.pct_of_peak_sustained_elapsed": "tensor_pipe_pct",
    "TPC.TriageCompute.sm__pipe_fp64_cycles_active_realtime.avg.pct_of_peak_sustained_elapsed": "fp64_pipe_cycles_pct",
    "l1tex__data_pipe_lsu_wavefronts.avg.pct_of_peak_sustained_elapsed": "lsu_data_pipe_pct",
    "l1tex__data_pipe_lsu_wavefronts_mem_shared.sum": "smem_wavefronts",
    "l1tex__data_pipe_lsu_wavefronts_mem_shared.sum.pct_of_peak_sustained_elapsed": "smem_wavefronts_pct_of_peak",
    "l1tex__data_bank_conflicts_pipe_lsu_mem_shared.sum": "smem_bank_conflicts",
    "lts__t_sector_hit_rate.pct": "l2_hit_pct",
    "l1tex__t_sector_hit_rate.pct": "l1_hit_pct",
    "launch__registers_per_thread": "registers",
    "launch__grid_size": "grid",
    "launch__block_size": "block",
    "launch__shared_mem_per_block_dynamic": "dyn_smem",
    "launch__shared_mem_per_block_static": "static_smem",
    "launch__occupancy_limit_shared_mem": "occupancy_limit_smem",
    "launch__occupancy_limit_registers": "occupancy_limit_regs",
    "sm__maximum_warps_per_active_cycle_pct": "theoretical_occupancy_pct",
}
SCALE = {"byte": 1.0, "Kbyte": 1e3, "Mbyte": 1e6, "Gbyte": 1e9, "ns": 1e-9, "us": 1e-6, "ms": 1e-3, "s": 1.0, "usecond": 1e-6,
         "msecond": 1e-3, "nsecond": 1e-9, "second": 1.0}


def main():
    rep = sys.argv[1]
    per = float(sys.argv[2]) if len(sys.argv) > 2 else None
    out = subprocess.run(["ncu", "-i", rep, "--page", "raw", "--csv"], capture_output=True, text=True, check=True).stdout
    rows = list(csv.reader(io.StringIO(out)))
    hdr, units = rows[0], rows[1]
    res = []
    for vals in rows[2:]:
        rec = {"kernel": vals[hdr.index("Kernel Name")]}
        for i, h in enumerate(hdr):
            if h in WANT and vals[i] != "":
                v = float(vals[i].replace(",", ""))
                u = units[i]
                if u in SCALE:
                    v *= SCALE[u]
                    u = "B" if "byte" in u else "s"
                rec[WANT[h]] = v
                rec[WANT[h] + "_unit"] = u
        if per and "dram_read" in rec:
            rec["candidates_per_launch"] = per
            rec["dram_bytes_per_candidate"] = (rec["dram_read"] + rec["dram_write"]) / per
            if "smem_wavefronts" in rec:
                rec["smem_wavefronts_per_candidate"] = rec["smem_wavefronts"] / per
            if "warp_instructions" in rec:
                rec["warp_instructions_per_candidate"] = rec["warp_instructions"] / per
        res.append(rec)
    json.dump(res, sys.stdout, indent=1)
    print()


if __name__ == "__main__":
    main()
