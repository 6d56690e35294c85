"""Turns `ncu --set full` captures into profiles/ncu_full_r2.md and profiles/roofline_traffic_r2.json (DRAM bytes per
launch, read by bench.py for `roofline.traffic`).
Usage: python tools/summarize_ncu_r2.py key=path.ncu-rep:"title" ...   (key conv32 feeds bench.py)"""
import csv
import io
import json
import subprocess
import sys

KEYS = ["gpu__time_duration.sum", "dram__bytes_read.sum", "dram__bytes_write.sum", "l1tex__m_xbar2l1tex_read_bytes.sum",
        "lts__t_sector_hit_rate.pct", "sm__pipe_tensor_cycles_active.avg.pct_of_peak_sustained_elapsed",
        "sm__throughput.avg.pct_of_peak_sustained_elapsed", "sm__warps_active.avg.pct_of_peak_sustained_active",
        "l1tex__data_pipe_lsu_wavefronts_mem_shared.sum", "smsp__inst_executed.sum",
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


def main(argv):
    lines = ["# ncu --set full, round 2 (one launch each, --clock-control none; regenerate with tools/summarize_ncu_r2.py)", ""]
    traffic = {"source": "profiles/ncu_full_r2.md (dram__bytes_read.sum + dram__bytes_write.sum, one launch)"}
    for arg in argv:
        key, rest = arg.split("=", 1)
        path, title = rest.split(":", 1)
        try:
            d, stalls = raw(path)
        except Exception as e:  # a capture that did not happen must not hide the others
            lines += [f"## {title}", "", f"(capture unavailable: {e})", ""]
            continue
        lines += ["## " + title, "", "| metric | value |", "|---|---|"]
        for k in KEYS:
            if k in d:
                lines.append(f"| `{k}` | {d[k][0]} {d[k][1]} |")
        top = ", ".join(f"{h.split('stalled_')[1].split('_per_issue')[0]} {v:.1f}" for v, h in stalls[:4])
        lines += ["", "Top stall reasons (warps per issue): " + top + ".", ""]
        try:
            traffic[key + "_dram_bytes_per_launch"] = int(to_bytes(*d["dram__bytes_read.sum"]) + to_bytes(*d["dram__bytes_write.sum"]))
        except KeyError:
            pass
    if "conv32_dram_bytes_per_launch" in traffic:
        traffic["conv3x3_c32_bs256_dram_bytes_per_launch"] = traffic["conv32_dram_bytes_per_launch"]
    open("profiles/ncu_full_r2.md", "w").write("\n".join(lines) + "\n")
    json.dump(traffic, open("profiles/roofline_traffic_r2.json", "w"), indent=1)
    print("\n".join(lines))


if __name__ == "__main__":
    main(sys.argv[1:])
