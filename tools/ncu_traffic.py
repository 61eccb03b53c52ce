"""dram__bytes_read.sum + dram__bytes_write.sum per launch, averaged per kernel family, from an `ncu --set full` report
of one training step (tools/profile_step.py) -> the JSON bench.py reads for `roofline.traffic`.
usage: python tools/ncu_traffic.py gpurun_out/prof.ncu-rep "capture description" > profiles/r02_ncu_traffic.json"""
import csv, io, json, subprocess, sys

rep, source = sys.argv[1], (sys.argv[2] if len(sys.argv) > 2 else sys.argv[1])
raw = subprocess.run(["ncu", "-i", rep, "--page", "raw", "--csv"], capture_output=True, text=True).stdout
rows = list(csv.reader(io.StringIO(raw)))
hdr, units, data = rows[0], rows[1], rows[2:]
col = {h: i for i, h in enumerate(hdr)}
scale = {"byte": 1.0, "Kbyte": 1e3, "Mbyte": 1e6, "Gbyte": 1e9}


def val(r, name):
    i = col[name]
    return float(r[i].replace(",", "")) * scale.get(units[i], 1.0)


fam = {}
for r in data:
    name = r[col["Kernel Name"]]
    if "conv_wa_kernel" in name or "conv_pipe_kernel" in name:
        f = "conv_gemm"
    elif "conv_wgrad" in name:
        f = "conv_wgrad"
    else:
        continue
    fam.setdefault(f, []).append(val(r, "dram__bytes_read.sum") + val(r, "dram__bytes_write.sum"))
print(json.dumps({"source": source, "launches": {k: len(v) for k, v in fam.items()},
                  "bytes_per_launch": {k: sum(v) / len(v) for k, v in fam.items()}}, indent=1))
