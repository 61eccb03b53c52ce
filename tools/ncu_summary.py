"""Summarise an .ncu-rep (read offline with `ncu -i ... --page raw --csv`) into a markdown table.
usage: python tools/ncu_summary.py gpurun_out/prof.ncu-rep "title" > profiles/xxx.md"""
import csv, io, subprocess, sys
rep, title = sys.argv[1], (sys.argv[2] if len(sys.argv) > 2 else sys.argv[1])
raw = subprocess.run(["ncu", "-i", rep, "--page", "raw", "--csv"], capture_output=True, text=True).stdout
rows = list(csv.reader(io.StringIO(raw)))
hdr, units, data = rows[0], rows[1], rows[2:]
def col(name):
    for i, h in enumerate(hdr):
        if h == name:
            return i
    return None
M = [("Kernel Name", "kernel", None), ("Grid Size", "grid", None), ("gpu__time_duration.sum", "us", 1e-3),
     ("dram__bytes_read.sum", "dram rd MB", None), ("dram__bytes_write.sum", "dram wr MB", None),
     ("sm__pipe_tensor_subpipe_hmma_cycles_active.avg.pct_of_peak_sustained_active", "tensor pipe %", None),
     ("sm__throughput.avg.pct_of_peak_sustained_elapsed", "SM %", None),
     ("gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed", "DRAM %", None),
     ("sm__warps_active.avg.pct_of_peak_sustained_active", "warps active %", None),
     ("launch__registers_per_thread", "regs", None), ("launch__shared_mem_per_block_dynamic", "dyn smem KB", None),
     ("launch__occupancy_limit_shared_mem", "CTA/SM (smem)", None)]
idx = [(col(n), label, scale) for n, label, scale in M]
print(f"# {title}\n")
print("Read with `ncu -i <rep> --page raw --csv`; capture: `ncu --set full --clock-control none --import-source on`.\n")
print("| " + " | ".join(l for _, l, _ in idx) + " |")
print("|" + "---|" * len(idx))
for r in data:
    cells = []
    for i, label, scale in idx:
        if i is None:
            cells.append("n/a"); continue
        v = r[i]
        if label == "kernel":
            v = "`" + v.split("(")[0].replace("void ", "").replace("cvae::", "")[:48] + "`"
        elif label in ("dram rd MB", "dram wr MB"):
            u = units[i]
            f = float(v.replace(",", ""))
            f = f * {"byte": 1e-6, "Kbyte": 1e-3, "Mbyte": 1.0, "Gbyte": 1e3}.get(u, 1.0)
            v = f"{f:.2f}"
        elif label == "us":
            u = units[i]
            f = float(v.replace(",", "")) * {"ns": 1e-3, "us": 1.0, "ms": 1e3}.get(u, 1.0)
            v = f"{f:.1f}"
        else:
            try:
                v = f"{float(v.replace(',', '')):.1f}" if "." in v else v
            except ValueError:
                pass
        cells.append(v)
    print("| " + " | ".join(cells) + " |")
