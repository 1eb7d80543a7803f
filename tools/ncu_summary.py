"""Text summary of an .ncu-rep (read in the GPU-less container): per kernel the headline sections, the DRAM bytes
against the algorithmic bytes, the stall-reason samples and the hottest SASS lines.
usage: python tools/ncu_summary.py report.ncu-rep [algorithmic_bytes_per_launch] > profiles/<name>.txt"""
import collections
import csv
import io
import subprocess
import sys


def ncu(path, page):
    return subprocess.run(["ncu", "-i", path, "--page", page, "--csv"], capture_output=True, text=True).stdout


def num(s):
    try:
        return float(s.replace(",", ""))
    except ValueError:
        return None


def main():
    path = sys.argv[1]
    algo = float(sys.argv[2]) if len(sys.argv) > 2 else None
    rows = list(csv.reader(io.StringIO(ncu(path, "raw"))))
    hdr, units = rows[0], rows[1]
    want = ["gpu__time_duration.sum", "launch__grid_size", "launch__block_size", "launch__registers_per_thread",
            "launch__cluster_max_active", "sm__warps_active.avg.pct_of_peak_sustained_active", "dram__bytes_read.sum", "dram__bytes_write.sum",
            "dram__throughput.avg.pct_of_peak_sustained_elapsed", "lts__throughput.avg.pct_of_peak_sustained_elapsed",
            "lts__t_sector_hit_rate.pct", "l1tex__throughput.avg.pct_of_peak_sustained_elapsed",
            "sm__throughput.avg.pct_of_peak_sustained_elapsed", "smsp__issue_active.avg.pct_of_peak_sustained_active",
            "smsp__inst_executed.sum", "l1tex__data_pipe_lsu_wavefronts_mem_shared.sum",
            "l1tex__data_bank_conflicts_pipe_lsu_mem_shared.sum", "sm__inst_executed_pipe_lsu.avg.pct_of_peak_sustained_active",
            "sm__inst_executed_pipe_fma.avg.pct_of_peak_sustained_active"]
    print(f"# {path}  (ncu --set full --clock-control none; per-launch values, cold-cache and serialised)")
    for r in rows[2:]:
        d = dict(zip(hdr, r))
        print(f"\n== {d.get('Kernel Name', '?')}")
        for k in want:
            if k in d:
                print(f"   {k:72s} {d[k]:>16s} {units[hdr.index(k)]}")
        rd, wr = num(d.get("dram__bytes_read.sum", "")), num(d.get("dram__bytes_write.sum", ""))
        if rd is not None and wr is not None:
            scale = {"Gbyte": 1e9, "Mbyte": 1e6, "Kbyte": 1e3, "byte": 1.0, "Tbyte": 1e12}
            rd *= scale.get(units[hdr.index("dram__bytes_read.sum")], 1.0)
            wr *= scale.get(units[hdr.index("dram__bytes_write.sum")], 1.0)
            line = f"   DRAM traffic per launch: {rd + wr:.4e} B (read {rd:.4e}, write {wr:.4e})"
            if algo:
                line += f" = {(rd + wr) / algo:.3f} x the algorithmic {algo:.4e} B"
            print(line)
        stalls = sorted(((num(d[k]) or 0.0, k.replace("smsp__pcsamp_warps_issue_stalled_", "")) for k in hdr
                         if k.startswith("smsp__pcsamp_warps_issue_stalled_") and not k.endswith("_not_issued")), reverse=True)
        tot = sum(v for v, _ in stalls) or 1.0
        print("   warp-state samples: " + ", ".join(f"{k} {100 * v / tot:.0f}%" for v, k in stalls[:8]))
    src = list(csv.reader(io.StringIO(ncu(path, "source"))))
    heads = [i for i, r in enumerate(src) if r and r[0] == "Address"]
    names = [r[1] for r in src if r and r[0] == "Kernel Name"]
    for n, hi in enumerate(heads):
        ix = {k: i for i, k in enumerate(src[hi])}
        end = heads[n + 1] - 1 if n + 1 < len(heads) else len(src)
        top = []
        for r in src[hi + 1:end]:
            if len(r) < len(src[hi]):
                continue
            s = num(r[ix["# Samples"]])
            if s:
                top.append((s, r[ix["Source"]][:100]))
        top.sort(reverse=True)
        total = sum(s for s, _ in top) or 1.0
        print(f"\n-- hottest SASS lines, kernel {n} ({names[n][:90] if n < len(names) else ''}): share of {int(total)} samples")
        for s, t in top[:12]:
            print(f"   {100 * s / total:5.1f}%  {t}")


if __name__ == "__main__":
    main()
