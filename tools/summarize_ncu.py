"""Turns the two `ncu --set full` captures of tools/gpu_profile.sh into profiles/ncu_full_r1.md and
profiles/roofline_traffic_r1.json (DRAM bytes per launch, read by bench.py for `roofline.traffic`).
Usage: python tools/summarize_ncu.py gpurun_out/conv32_full_r1.ncu-rep gpurun_out/dist_full_r1.ncu-rep"""
import csv
import io
import json
import subprocess
import sys

KEYS = ["gpu__time_duration.sum", "dram__bytes_read.sum", "dram__bytes_write.sum", "l1tex__m_xbar2l1tex_read_bytes.sum",
        "lts__t_sector_hit_rate.pct", "sm__pipe_tensor_cycles_active.avg.pct_of_peak_sustained_elapsed",
        "sm__throughput.avg.pct_of_peak_sustained_elapsed", "sm__warps_active.avg.pct_of_peak_sustained_active",
        "launch__grid_size", "launch__block_size", "launch__registers_per_thread", "launch__occupancy_limit_shared_mem",
        "launch__occupancy_limit_registers"]


def raw(path):
    out = subprocess.run(["ncu", "-i", path, "--page", "raw", "--csv"], capture_output=True, text=True).stdout
    rows = list(csv.reader(io.StringIO(out)))
    hdr, units, vals = rows[0], rows[1], rows[2]
    d = {h: (v, u) for h, u, v in zip(hdr, units, vals)}
    stalls = sorted(((float(v.replace(",", "")), h) for h, (v, u) in d.items()
                     if h.startswith("smsp__average_warps_issue_stalled") and h.endswith("_per_issue_active.ratio")
                     and v not in ("", "n/a")), reverse=True)
    return d, stalls


def to_bytes(v, u):
    x = float(v.replace(",", ""))
    return x * {"byte": 1, "Kbyte": 1e3, "Mbyte": 1e6, "Gbyte": 1e9}.get(u, 1)


def section(title, d, stalls):
    lines = ["## " + title, "", "| metric | value |", "|---|---|"]
    for k in KEYS:
        if k in d:
            lines.append(f"| `{k}` | {d[k][0]} {d[k][1]} |")
    top = ", ".join(f"{h.split('stalled_')[1].split('_per_issue')[0]} {v:.1f}" for v, h in stalls[:4])
    lines += ["", "Top stall reasons (warps per issue): " + top + ".", ""]
    return lines


def main(conv_rep, dist_rep):
    cd, cs = raw(conv_rep)
    dd, ds = raw(dist_rep)
    lines = ["# ncu --set full, round 1 (one launch each, --clock-control none; regenerate with tools/summarize_ncu.py)", ""]
    lines += section("tc_gemm_kernel<32,bf16,raw+stats>: conv 3x3 32->32 @64x32, batch 256 (tools/probe_conv.py)", cd, cs)
    lines += section("part_distance_tc_kernel<bool,euclid,mean>: 4096 q x 16384 g x 6 parts x 512 (tools/probe_eval.py)", dd, ds)
    open("profiles/ncu_full_r1.md", "w").write("\n".join(lines) + "\n")
    traffic = {
        "conv3x3_c32_bytes_per_launch": int(to_bytes(*cd["dram__bytes_read.sum"]) + to_bytes(*cd["dram__bytes_write.sum"])),
        "part_distance_bytes_per_launch": int(to_bytes(*dd["dram__bytes_read.sum"]) + to_bytes(*dd["dram__bytes_write.sum"])),
        "source": "profiles/ncu_full_r1.md (dram__bytes_read.sum + dram__bytes_write.sum, one launch)"}
    json.dump(traffic, open("profiles/roofline_traffic_r1.json", "w"), indent=1)
    print("\n".join(lines))


if __name__ == "__main__":
    main(sys.argv[1], sys.argv[2])
