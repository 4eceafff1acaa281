"""profiles/gemm_traffic.json: mean DRAM bytes per launch of the tcgen05 GEMM over the twelve GEMM shapes of one
cfg2 layer, from an `ncu --set full` capture of scripts/gemm_micro.py (ITERS=1: 4 launches per shape, the last
of each group is taken).  bench.py reports it as roofline.traffic."""
import csv
import io
import json
import subprocess
import sys

rep, out = sys.argv[1], sys.argv[2]
txt = subprocess.run(["ncu", "-i", rep, "--page", "raw", "--csv"], capture_output=True, text=True).stdout
rows = list(csv.reader(io.StringIO(txt)))
hdr, units = rows[0], rows[1]
col = {h: i for i, h in enumerate(hdr)}
scale = {"byte": 1, "Kbyte": 1e3, "Mbyte": 1e6, "Gbyte": 1e9}
launches = []
for r in rows[2:]:
    if "gemm_tc" not in r[col["Kernel Name"]]:
        continue
    b = sum(float(r[col[k]].replace(",", "")) * scale[units[col[k]]] for k in ("dram__bytes_read.sum", "dram__bytes_write.sum"))
    launches.append({"kernel": r[col["Kernel Name"]][:40], "us": float(r[col["gpu__time_duration.sum"]]),
                     "dram_bytes": b, "tensor_pct": float(r[col["sm__pipe_tensor_cycles_active.avg.pct_of_peak_sustained_elapsed"]])})
per_shape = launches[3::4][:12]
json.dump({"source": rep, "mean_dram_bytes_per_launch": sum(l["dram_bytes"] for l in per_shape) / len(per_shape),
           "launches": per_shape}, open(out, "w"), indent=1)
print(len(launches), "gemm launches,", len(per_shape), "shapes")
