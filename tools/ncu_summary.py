"""Condense an `ncu --page raw --csv` export into the few numbers DESIGN.md / bench.py cite.

  ncu -i X.ncu-rep --page raw --csv > X_raw.csv ; python tools/ncu_summary.py X_raw.csv > profiles/X_summary.json
Bytes are converted to plain bytes, durations to microseconds. One record per profiled launch.
"""
import csv
import json
import sys

KEEP = {
    "gpu__time_duration.sum": "duration_us",
    "dram__bytes_read.sum": "dram_read_bytes",
    "dram__bytes_write.sum": "dram_write_bytes",
    "lts__t_sectors_srcunit_tex_op_read.sum": "l2_to_l1_read_sectors",
    "l1tex__t_sector_hit_rate.pct": "l1_hit_pct",
    "lts__t_sector_hit_rate.pct": "l2_hit_pct",
    "l1tex__data_pipe_lsu_wavefronts.sum.pct_of_peak_sustained_elapsed": "lsu_data_pipe_pct",
    "lts__throughput.avg.pct_of_peak_sustained_elapsed": "l2_throughput_pct",
    "dram__throughput.avg.pct_of_peak_sustained_elapsed": "dram_throughput_pct",
    "sm__throughput.avg.pct_of_peak_sustained_elapsed": "sm_throughput_pct",
    "smsp__issue_active.avg.pct_of_peak_sustained_active": "issue_active_pct",
    "sm__warps_active.avg.pct_of_peak_sustained_active": "warps_active_pct",
    "launch__registers_per_thread": "registers",
    "smsp__inst_executed.sum": "warp_instructions",
    "sm__inst_executed_pipe_fp64.avg.pct_of_peak_sustained_active": "pipe_fp64_pct",
    "sm__inst_executed_pipe_fma.avg.pct_of_peak_sustained_active": "pipe_fma_pct",
    "sm__inst_executed_pipe_alu.avg.pct_of_peak_sustained_active": "pipe_alu_pct",
    "sm__inst_executed_pipe_lsu.avg.pct_of_peak_sustained_active": "pipe_lsu_pct",
    "sm__inst_executed_pipe_tensor.sum": "pipe_tensor_inst",
    "sm__inst_executed_pipe_tensor.avg.pct_of_peak_sustained_active": "pipe_tensor_pct",
    "smsp__average_warps_issue_stalled_long_scoreboard_per_issue_active.ratio": "stall_long_scoreboard",
    "smsp__average_warp_latency_per_inst_issued.ratio": "warp_latency_per_inst",
    "sm__cycles_elapsed.avg": "sm_cycles",
}
SCALE = {"byte": 1, "Kbyte": 1e3, "Mbyte": 1e6, "Gbyte": 1e9, "Tbyte": 1e12, "ns": 1e-3, "us": 1, "usecond": 1, "ms": 1e3,
         "msecond": 1e3, "second": 1e6, "nsecond": 1e-3}


def main(path):
    rows = list(csv.reader(open(path)))
    hdr, units, data = rows[0], rows[1], [r for r in rows[2:] if len(r) == len(rows[0])]
    out = []
    for r in data:
        rec = {"kernel": r[hdr.index("Kernel Name")], "grid": r[hdr.index("Grid Size")], "block": r[hdr.index("Block Size")]}
        for k, name in KEEP.items():
            if k in hdr:
                i = hdr.index(k)
                try:
                    v = float(r[i].replace(",", ""))
                except ValueError:
                    continue
                rec[name] = v * SCALE.get(units[i], 1)
        out.append(rec)
    json.dump({"source": path, "launches": out}, sys.stdout, indent=1)


if __name__ == "__main__":
    main(sys.argv[1])
