#!/bin/bash
# ncu launch list of one steady-state step of the bench command (gpu__time_duration.sum per launch), summarised by
# scripts/summarize_launches.py.  Only a window of launches is profiled (--launch-skip / --launch-count): the first
# forward pass issues ~735 launches (one-off table builds), every later step 315, so launches [1300, 2000) hold two
# complete steps; profiling all ~2000 launches took 146 s of box time, the window takes a third of that.
mkdir -p gpurun_out
timeout 300 ncu --metrics gpu__time_duration.sum --clock-control none --launch-skip ${SKIP:-1300} --launch-count ${COUNT:-700} \
  --csv --log-file gpurun_out/launches.csv python bench.py --steps 1 --warmup 3 --no-cpu-baseline --no-decode ${BENCH_ARGS} \
  > gpurun_out/ncu_bench.log 2>&1
echo "exit $?"; wc -l gpurun_out/launches.csv
python - <<'PY'
import csv, subprocess, sys
lines = [l for l in open("gpurun_out/launches.csv") if not l.startswith("==")]
names = [r["Kernel Name"] for r in csv.DictReader(lines) if r.get("Metric Name") == "gpu__time_duration.sum"]
starts = [i for i, n in enumerate(names) if "convert_batched" in n]      # first launch of every forward pass
print("step starts at", starts)
if len(starts) >= 2:
    subprocess.run([sys.executable, "scripts/summarize_launches.py", "gpurun_out/launches.csv", "40", str(starts[0]),
                    str(starts[1])])
PY
