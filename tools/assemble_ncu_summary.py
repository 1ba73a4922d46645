"""profiles/ncu_summary.json <- the gather kernels' entries of a per-round capture file (profiles/ncu_<tag>.json made by
tools/ncu_summary.py), stamped with the hash of the sources they were built from.
Usage: python tools/assemble_ncu_summary.py <tag>"""
import json
import os
import sys

sys.path.insert(0, os.path.dirname(os.path.abspath(__file__)))
import build_hash  # noqa: E402

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
FIELDS = ("dram_bytes_per_launch", "l1_hit_pct", "l2_hit_pct", "l1_throughput_pct", "l2_throughput_pct", "l1_data_pipe_pct",
          "l1_to_xbar_req_busy_pct", "achieved_occupancy_pct", "issue_active_pct", "registers", "stall_pct", "tensor_pipe_pct",
          "warp_instructions")


def main():
    tag = sys.argv[1]
    blob = json.load(open(os.path.join(ROOT, "profiles", f"ncu_{tag}.json")))
    path = os.path.join(ROOT, "profiles", "ncu_summary.json")
    summary = json.load(open(path))
    want = {"msda_fwd_fast_kernel<float, 32, 0, float>": ("f32", "forward"),
            "msda_bwd_fast_kernel<float, 32, 0, float>": ("f32", "backward"),
            "msda_fwd_fast_kernel<__nv_bfloat16, 32, 0, float>": ("bf16", "forward"),
            "msda_bwd_fast_kernel<__nv_bfloat16, 32, 0, float>": ("bf16", "backward")}
    for kernels in blob.values():
        for k in kernels:
            name = k["kernel"].replace("msda::", "")
            if name in want:
                dt, kind = want[name]
                entry = {"kernel": name, "duration_s_under_ncu": k.get("duration")}
                entry.update({f: k[f] for f in FIELDS if f in k})
                summary.setdefault(dt, {})[kind] = entry
    summary["gather_kernels_build"] = {"sha256_16": build_hash.gather_kernels_hash(), "capture": f"profiles/ncu_{tag}.json",
                                       "files": ["csrc/msda_common.cuh", "csrc/msda_forward.cu", "csrc/msda_backward.cu"]}
    json.dump(summary, open(path, "w"), indent=1)
    print(json.dumps({k: summary[k] for k in ("f32", "bf16", "gather_kernels_build")}, indent=1)[:3000])


if __name__ == "__main__":
    main()
