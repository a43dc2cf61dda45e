"""Turn an .ncu-rep into the short text summary committed under profiles/ (run here, no GPU needed).

    python profiles/summarize_ncu.py gpurun_out/decode_full.ncu-rep > profiles/r01_decode_kernel_ncu.txt
"""
import csv
import io
import subprocess
import sys

KEYS = ["gpu__time_duration.sum", "dram__bytes_read.sum", "dram__bytes_write.sum",
        "gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed", "dram__cycles_active.avg.pct_of_peak_sustained_elapsed",
        "sm__throughput.avg.pct_of_peak_sustained_elapsed", "smsp__issue_active.avg.pct_of_peak_sustained_active",
        "sm__warps_active.avg.pct_of_peak_sustained_active", "launch__registers_per_thread", "launch__grid_size",
        "launch__block_size", "launch__shared_mem_per_block_dynamic", "smsp__inst_executed.sum",
        "sm__pipe_fma_cycles_active.avg.pct_of_peak_sustained_active", "sm__pipe_alu_cycles_active.avg.pct_of_peak_sustained_active",
        "sm__inst_executed_pipe_lsu.avg.pct_of_peak_sustained_active", "sm__pipe_tensor_cycles_active.avg.pct_of_peak_sustained_active",
        "l1tex__data_bank_conflicts_pipe_lsu_mem_shared.sum", "lts__t_sector_hit_rate.pct", "sm__cycles_elapsed.max",
        "smsp__cycles_active.avg", "smsp__warps_eligible.avg.per_cycle_active"]


def page(rep, name):
    out = subprocess.run(["ncu", "-i", rep, "--page", name, "--csv"], capture_output=True, text=True).stdout
    return list(csv.reader(io.StringIO(out)))


def main():
    rep = sys.argv[1]
    rows = page(rep, "raw")
    hdr, units = rows[0], rows[1]
    for r in rows[2:]:
        if len(r) != len(hdr):
            continue
        print(f"== launch {r[hdr.index('ID')]}: {r[hdr.index('Kernel Name')][:90]}")
        for k in KEYS:
            if k in hdr:
                i = hdr.index(k)
                print(f"   {k:72s} {r[i]:>18s} {units[i]}")
        stalls = []
        for i, h in enumerate(hdr):
            if "issue_stalled" in h and h.endswith("per_issue_active.ratio"):
                try:
                    stalls.append((float(r[i].replace(",", "")), h.split("issue_stalled_")[1].split("_per_issue")[0]))
                except ValueError:
                    pass
        print("   top warp stalls (avg warps stalled per issue-active cycle): " +
              ", ".join(f"{n}={v:.2f}" for v, n in sorted(stalls, reverse=True)[:8]))
    src = page(rep, "source")
    if len(src) > 2:
        h = src[1]
        n = len(h)
        data = [r for r in src[2:] if len(r) == n]
        iS, isrc = h.index("# Samples"), h.index("Source")
        f = lambda x: int(x) if x.isdigit() else 0
        tot = sum(f(r[iS]) for r in data) or 1
        seen, out = set(), []
        for r in sorted(data, key=lambda r: -f(r[iS])):
            if r[isrc] in seen:
                continue
            seen.add(r[isrc])
            out.append(f"   {100 * f(r[iS]) / tot:5.2f}%  {r[isrc].strip()[:100]}")
            if len(out) == 14:
                break
        print("== hottest SASS (share of stall samples, duplicates across launches merged)")
        print("\n".join(out))


if __name__ == "__main__":
    main()
