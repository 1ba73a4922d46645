"""Summarise ncu reports into profiles/ (tracked): per-kernel key metrics as JSON + a launch list.
Usage: python tools/ncu_summary.py <tag> <report.ncu-rep> [<report2.ncu-rep> ...]"""
import csv
import io
import json
import os
import subprocess
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
KEYS = {
    "gpu__time_duration.sum": "duration",
    "dram__bytes_read.sum": "dram_read",
    "dram__bytes_write.sum": "dram_write",
    "gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed": "dram_pct",
    "lts__t_sector_hit_rate.pct": "l2_hit_pct",
    "lts__t_sectors.sum": "l2_sectors",
    "lts__t_sectors_srcunit_tex_op_red.sum": "l2_red_sectors",
    "lts__throughput.avg.pct_of_peak_sustained_elapsed": "l2_throughput_pct",
    "l1tex__t_sector_hit_rate.pct": "l1_hit_pct",
    "l1tex__throughput.avg.pct_of_peak_sustained_elapsed": "l1_throughput_pct",
    "l1tex__t_sectors_pipe_lsu_mem_global_op_ld.sum": "l1_global_load_sectors",
    "l1tex__t_sectors_pipe_lsu_mem_global_op_ld_lookup_hit.sum": "l1_global_load_hit_sectors",
    "l1tex__t_sectors_pipe_lsu_mem_global_op_red.sum": "l1_red_sectors",
    "l1tex__data_pipe_lsu_wavefronts_mem_shared.sum": "smem_wavefronts",
    "l1tex__data_bank_conflicts_pipe_lsu_mem_shared.sum": "smem_bank_conflicts",
    "l1tex__m_xbar2l1tex_read_bytes.sum": "l2_to_l1_bytes",
    "l1tex__m_l1tex2xbar_write_bytes.sum": "l1_to_l2_bytes",
    "sm__throughput.avg.pct_of_peak_sustained_elapsed": "sm_throughput_pct",
    "sm__warps_active.avg.pct_of_peak_sustained_active": "achieved_occupancy_pct",
    "smsp__issue_active.avg.pct_of_peak_sustained_active": "issue_active_pct",
    "smsp__inst_executed.sum": "warp_instructions",
    "launch__registers_per_thread": "registers",
    "launch__grid_size": "grid",
    "launch__block_size": "block",
    "sm__cycles_elapsed.max": "cycles",
    "sm__pipe_tensor_cycles_active.avg.pct_of_peak_sustained_active": "tensor_pipe_pct",
    # what bounds the gather kernels: the L1 data pipe (128 B/clk/SM, forward) and the L1 -> crossbar
    # request path the reductions leave the SM through (backward)
    "l1tex__data_pipe_lsu_wavefronts.avg.pct_of_peak_sustained_elapsed": "l1_data_pipe_pct",
    "l1tex__data_pipe_lsu_wavefronts.sum": "l1_data_pipe_wavefronts",
    "l1tex__m_l1tex2xbar_req_cycles_active.avg.pct_of_peak_sustained_elapsed": "l1_to_xbar_req_busy_pct",
    "l1tex__m_l1tex2xbar_write_bytes.sum.pct_of_peak_sustained_elapsed": "l1_to_xbar_write_pct",
    "l1tex__t_output_wavefronts_pipe_lsu_mem_global_op_red.sum": "l1_red_wavefronts",
    "smsp__average_warps_issue_stalled_long_scoreboard_per_issue_active.ratio": "stall_long_scoreboard_ratio",
    "smsp__average_warps_issue_stalled_lg_throttle_per_issue_active.ratio": "stall_lg_throttle_ratio",
}
STALLS = "smsp__pcsamp_warps_issue_stalled_"


def to_num(txt):
    try:
        return float(txt.replace(",", ""))
    except Exception:
        return txt


def scale(value, unit):
    mult = {"Kbyte": 1e3, "Mbyte": 1e6, "Gbyte": 1e9, "Tbyte": 1e12, "byte": 1.0,
            "us": 1e-6, "ms": 1e-3, "ns": 1e-9, "s": 1.0, "usecond": 1e-6, "msecond": 1e-3, "nsecond": 1e-9, "second": 1.0}
    return value * mult[unit] if isinstance(value, float) and unit in mult else value


def summarise(report):
    raw = subprocess.run(["ncu", "-i", report, "--page", "raw", "--csv"], stdout=subprocess.PIPE, text=True).stdout
    rows = list(csv.reader(io.StringIO(raw)))
    hdr, units = rows[0], rows[1]
    out = []
    for r in rows[2:]:
        k = {"kernel": r[hdr.index("Kernel Name")].split("(")[0].replace("void ", "")}
        for i, h in enumerate(hdr):
            if h in KEYS:
                k[KEYS[h]] = scale(to_num(r[i]), units[i])
        stalls = {h[len(STALLS):]: to_num(r[i]) for i, h in enumerate(hdr) if h.startswith(STALLS) and "not_issued" not in h}
        tot = sum(v for v in stalls.values() if isinstance(v, float)) or 1.0
        k["stall_pct"] = {n: round(100 * v / tot, 1) for n, v in sorted(stalls.items(), key=lambda kv: -kv[1])[:6]}
        if "dram_read" in k and "dram_write" in k:
            k["dram_bytes_per_launch"] = k["dram_read"] + k["dram_write"]
        out.append(k)
    return out


def main():
    tag = sys.argv[1]
    blob = {}
    for rep in sys.argv[2:]:
        blob[os.path.basename(rep)] = summarise(rep)
    path = os.path.join(ROOT, "profiles", f"ncu_{tag}.json")
    json.dump(blob, open(path, "w"), indent=1)
    print(json.dumps(blob, indent=1))


if __name__ == "__main__":
    main()
