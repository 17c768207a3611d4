"""Markdown table of the judged metrics from one or more `ncu --set full` reports (read with `ncu -i ... --page raw
--csv` on the CPU box).  Usage: ncu_summary.py rep1.ncu-rep [rep2.ncu-rep ...] > profiles/xxx.md"""
import csv
import io
import subprocess
import sys

COLS = [("gpu__time_duration.sum", "us", 1.0), ("sm__cycles_active.avg", "SM cycles", 1.0),
        ("sm__pipe_tensor_cycles_active.avg.pct_of_peak_sustained_active", "tensor pipe %", 1.0),
        ("smsp__issue_active.avg.pct_of_peak_sustained_active", "issue %", 1.0),
        ("dram__bytes_read.sum", "DRAM rd MB", 1.0), ("dram__bytes_write.sum", "DRAM wr MB", 1.0),
        ("gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed", "DRAM %", 1.0),
        ("lts__throughput.avg.pct_of_peak_sustained_elapsed", "L2 %", 1.0),
        ("l1tex__m_l1tex2xbar_write_bytes_mem_global_op_tma_st.sum", "TMA st MB", 1.0),
        ("l1tex__m_xbar2l1tex_read_bytes_mem_global_op_tma_ld.sum", "TMA ld MB", 1.0),
        ("launch__registers_per_thread", "regs", 1.0), ("launch__grid_size", "grid", 1.0)]

print("| kernel | " + " | ".join(c[1] for c in COLS) + " |")
print("|---|" + "---|" * len(COLS))
for rep in sys.argv[1:]:
    out = subprocess.run(["ncu", "-i", rep, "--page", "raw", "--csv"], capture_output=True, text=True).stdout
    rows = list(csv.reader(io.StringIO(out)))
    h, units = rows[0], rows[1]
    for r in rows[2:]:
        name = r[h.index("Kernel Name")]
        name = name.replace("apla::", "").replace("void ", "").split("(CUtensorMap")[0].split("(apla")[0].split("(const")[0]
        vals = []
        for key, _, _ in COLS:
            if key in h:
                v = r[h.index(key)]
                try:
                    f = float(v.replace(",", ""))
                    u = units[h.index(key)]
                    if u == "byte":
                        f /= 1e6
                    elif u == "Kbyte":
                        f /= 1e3
                    elif u == "Gbyte":
                        f *= 1e3
                    elif u == "ns":
                        f /= 1e3
                    elif u == "ms":
                        f *= 1e3
                    vals.append(f"{f:.1f}" if f < 1e5 else f"{f:.0f}")
                except ValueError:
                    vals.append(v)
            else:
                vals.append("-")
        print(f"| `{name[:48]}` | " + " | ".join(vals) + " |")
