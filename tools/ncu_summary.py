#!/usr/bin/env python3
"""Turn an .ncu-rep (read here, no GPU needed) into the text summary committed under profiles/:
  ncu_summary.py <report.ncu-rep> [--top 25]  -> key raw metrics per captured launch + the SASS
  lines that collected the most warp-stall samples."""
import csv
import io
import subprocess
import sys

KEYS = ["gpu__time_duration.sum", "dram__bytes_read.sum", "dram__bytes_write.sum",
        "dram__bytes.sum.per_second", "gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed",
        "lts__t_sector_hit_rate.pct", "l1tex__t_sector_hit_rate.pct",
        "lts__throughput.avg.pct_of_peak_sustained_elapsed",
        "sm__throughput.avg.pct_of_peak_sustained_elapsed",
        "smsp__issue_active.avg.pct_of_peak_sustained_active",
        "sm__warps_active.avg.pct_of_peak_sustained_active", "sm__maximum_warps_per_active_cycle_pct",
        "launch__registers_per_thread", "launch__occupancy_limit_registers", "launch__grid_size",
        "launch__block_size", "launch__waves_per_multiprocessor", "smsp__inst_executed.sum",
        "sm__cycles_elapsed.max", "smsp__cycles_active.avg"]


def page(rep, name):
    out = subprocess.run(["ncu", "-i", rep, "--page", name, "--csv"], capture_output=True, text=True).stdout
    return list(csv.reader(io.StringIO(out)))


def main():
    rep = sys.argv[1]
    top = int(sys.argv[sys.argv.index("--top") + 1]) if "--top" in sys.argv else 25
    rows = page(rep, "raw")
    hdr, units = rows[0], rows[1]
    print("# %s" % rep)
    for n, r in enumerate(rows[2:]):
        print("\n## launch %d: %s" % (n, r[hdr.index("Kernel Name")] if "Kernel Name" in hdr else ""))
        for k in KEYS:
            if k in hdr:
                i = hdr.index(k)
                print("%-62s %s %s" % (k, r[i], units[i]))
        stalls = []
        for i, h in enumerate(hdr):
            if "issue_stalled" in h and h.endswith("per_warp_active.pct"):
                try:
                    stalls.append((float(r[i]), h))
                except ValueError:
                    pass
        print("warp stall reasons (% of warp-active cycles, top 6):")
        for v, h in sorted(stalls, reverse=True)[:6]:
            print("  %6.1f  %s" % (v, h.replace("smsp__average_warps_issue_stalled_", "").replace("_per_warp_active.pct", "")))
    src = page(rep, "source")
    heads = [k for k, row in enumerate(src) if row and row[0] == "Address"]
    if heads:
        hk = heads[0]
        end = heads[1] - 1 if len(heads) > 1 else len(src)
        h = src[hk]
        ia, isamp = h.index("Source"), h.index("# Samples")
        body = [b for b in src[hk + 1:end] if len(b) > isamp and b[isamp].isdigit()]
        tot = sum(int(b[isamp]) for b in body) or 1
        print("\n## SASS hot spots (launch 0): %d instructions, %d stall samples" % (len(body), tot))
        order = sorted(range(len(body)), key=lambda n: -int(body[n][isamp]))[:top]
        for n in sorted(order):
            print("  #%-4d %5.1f%%  %s" % (n, 100.0 * int(body[n][isamp]) / tot, body[n][ia].strip()[:90]))
        print("\n## memory instructions in program order")
        for n, b in enumerate(body):
            t = b[ia]
            if any(x in t for x in ("LDG", "STG", "SHFL", "BAR", "CALL", "EXIT")):
                print("  #%-4d %5.1f%%  %s" % (n, 100.0 * int(b[isamp]) / tot, t.strip()[:90]))


if __name__ == "__main__":
    main()
