"""Per-kernel table from an `ncu --set full` report: `python tools/ncu_table.py report.ncu-rep [layer names ...]`.

Reads `ncu -i report --page raw --csv` (ncu must be on PATH; no GPU needed) and prints, per profiled launch: duration,
share, DRAM bytes (read + write), DRAM / L2 / tensor-pipe / issue utilisation, shared-memory wavefront utilisation,
registers, threads.  Optional layer names label the rows in launch order.  Also prints a JSON dict {label: DRAM bytes}.
"""
import csv
import io
import json
import subprocess
import sys

rep = sys.argv[1]
labels = sys.argv[2:]
raw = subprocess.run(["ncu", "-i", rep, "--page", "raw", "--csv"], check=True, capture_output=True, text=True).stdout
rows = list(csv.reader(io.StringIO(raw)))
hdr = rows[0]
data = [r for r in rows[2:] if len(r) == len(hdr)]  # rows[1] = units


def col(name_part, exact=False):
    for i, h in enumerate(hdr):
        if (h == name_part) if exact else h.endswith(name_part):
            return i
    return None


def num(r, i):
    if i is None:
        return float("nan")
    try:
        return float(r[i].replace(",", ""))
    except ValueError:
        return float("nan")


units = rows[1]
c_name = col("Kernel Name", True)
c_t = col("gpu__time_duration.sum", True)
c_rd, c_wr = col("dram__bytes_read.sum", True), col("dram__bytes_write.sum", True)
c_dram = col("gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed", True)
c_l2 = col("lts__throughput.avg.pct_of_peak_sustained_elapsed", True)
c_tc = col("sm__pipe_tensor_cycles_active.avg.pct_of_peak_sustained_active", True)
c_issue = col("sm__issue_active.avg.pct_of_peak_sustained_elapsed", True)
c_smem = col("l1tex__data_pipe_lsu_wavefronts_mem_shared.sum.pct_of_peak_sustained_elapsed", True)
c_tcsm = col("l1tex__data_pipe_tc_wavefronts_mem_shared.sum.pct_of_peak_sustained_elapsed", True)
c_regs = col("launch__registers_per_thread", True)
c_thr = col("launch__block_size", True)


def scale(i, want):
    """value multiplier so that column i is in unit `want` (ns->us, byte/Kbyte/Mbyte/Gbyte -> GB)"""
    u = units[i].lower() if i is not None else ""
    if want == "us":
        return {"ns": 1e-3, "us": 1.0, "ms": 1e3, "s": 1e6}.get(u, 1e-3)
    return {"byte": 1e-9, "kbyte": 1e-6, "mbyte": 1e-3, "gbyte": 1.0}.get(u, 1e-9)


tot = sum(num(r, c_t) * scale(c_t, "us") for r in data)
print(f"{'layer':<20}{'kernel':<50}{'us':>9}{'share%':>8}{'DRAM GB':>9}{'dram%':>7}{'L2%':>7}{'tensor%':>9}"
      f"{'issue%':>8}{'smem-lsu%':>10}{'smem-tc%':>9}{'regs':>6}{'thr':>6}")
traffic = {}
for k, r in enumerate(data):
    lab = labels[k] if k < len(labels) else f"#{k}"
    us = num(r, c_t) * scale(c_t, "us")
    gb = num(r, c_rd) * scale(c_rd, "GB") + num(r, c_wr) * scale(c_wr, "GB")
    traffic[lab] = int(gb * 1e9)
    print(f"{lab:<20}{r[c_name][:48]:<50}{us:>9.1f}{100 * us / tot:>8.1f}{gb:>9.3f}{num(r, c_dram):>7.1f}"
          f"{num(r, c_l2):>7.1f}{num(r, c_tc):>9.1f}{num(r, c_issue):>8.1f}{num(r, c_smem):>10.1f}{num(r, c_tcsm):>9.1f}"
          f"{num(r, c_regs):>6.0f}{num(r, c_thr):>6.0f}")
print(f"{'total':<70}{tot:>9.1f}")
print(json.dumps(traffic))
