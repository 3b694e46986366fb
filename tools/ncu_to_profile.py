"""Turns an `ncu -i <rep> --page raw --csv` export into profiles/<name>.json (+ a markdown summary on stdout).
bench.py reads roofline.traffic from that JSON (dram__bytes_read.sum + dram__bytes_write.sum per launch of the dominant kernel).
usage: python tools/ncu_to_profile.py <raw.csv> <out.json> "<source description>" """
import csv
import json
import sys

KEYS = {"gpu__time_duration.sum": "time_us", "dram__bytes_read.sum": "dram_bytes_read", "dram__bytes_write.sum": "dram_bytes_write",
        "smsp__inst_executed.sum": "warp_instructions", "smsp__issue_active.avg.pct_of_peak_sustained_active": "issue_active_pct",
        "sm__warps_active.avg.pct_of_peak_sustained_active": "warps_active_pct", "launch__registers_per_thread": "registers",
        "gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed": "dram_throughput_pct",
        "smsp__thread_inst_executed_per_inst_executed.ratio": "active_threads_per_instruction",
        "sm__inst_executed_pipe_alu.avg.pct_of_peak_sustained_active": "alu_pipe_pct", "sm__inst_executed_pipe_fma.avg.pct_of_peak_sustained_active": "fma_pipe_pct"}
UNIT = {"Mbyte": 1e6, "Kbyte": 1e3, "Gbyte": 1e9, "byte": 1.0}


def short(name):
    n = name.split("(")[0].replace("void ", "").replace("lvb::", "")
    base = n.split("<")[0]
    if base == "subsense_tail_pass":
        return "subsense_tail_pass2" if n.rstrip(">").endswith("1") or ", 1>" in n else "subsense_tail_pass1"
    return base


def main():
    rows = list(csv.reader(open(sys.argv[1])))
    hdr, units = rows[0], rows[1]
    out = {"source": sys.argv[3] if len(sys.argv) > 3 else sys.argv[1], "kernels": {}}
    for r in rows[2:]:
        k = short(r[hdr.index("Kernel Name")])
        d = {}
        for m, key in KEYS.items():
            if m in hdr and r[hdr.index(m)] not in ("", "n/a"):
                v = float(r[hdr.index(m)].replace(",", ""))
                u = units[hdr.index(m)]
                if key.startswith("dram_bytes"):
                    v *= UNIT.get(u, 1.0)
                if key == "time_us" and u in ("ns", "nsecond"):
                    v /= 1e3
                if key == "time_us" and u in ("ms", "msecond"):
                    v *= 1e3
                d[key] = v
        stalls = sorted(((float(r[i].replace(",", "")), h.replace("smsp__average_warps_issue_stalled_", "").replace("_per_issue_active.ratio", ""))
                         for i, h in enumerate(hdr) if h.startswith("smsp__average_warp") and "per_issue_active" in h and "not_issued" not in h and r[i] not in ("", "n/a")), reverse=True)
        d["top_stalls"] = {h: v for v, h in stalls[:5]}
        out["kernels"][k] = d
    json.dump(out, open(sys.argv[2], "w"), indent=1)
    print("| kernel | time (us) | DRAM read (MB) | DRAM write (MB) | warp instr (M) | issue active % | warps active % | regs | top stalls |")
    print("|---|---|---|---|---|---|---|---|---|")
    for k, d in out["kernels"].items():
        print(f"| `{k}` | {d.get('time_us', 0):.1f} | {d.get('dram_bytes_read', 0) / 1e6:.1f} | {d.get('dram_bytes_write', 0) / 1e6:.1f} | {d.get('warp_instructions', 0) / 1e6:.1f} | "
              f"{d.get('issue_active_pct', 0):.1f} | {d.get('warps_active_pct', 0):.1f} | {int(d.get('registers', 0))} | " + ", ".join(f"{h} {v:.2f}" for h, v in d["top_stalls"].items()) + " |")


if __name__ == "__main__":
    main()
