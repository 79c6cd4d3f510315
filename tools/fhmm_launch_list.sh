#!/bin/bash
# Per-kernel durations of one cy2p_fhmm_loss_grad call at config-3 shapes (serialised under ncu).
# usage (under gpurun): tools/fhmm_launch_list.sh <tag>   -> gpurun_out/r02/launches_fhmm_<tag>.csv + a summary on stdout
TAG=${1:-a}
OUT=gpurun_out/r02
mkdir -p $OUT
CMD="python tools/bench_fhmm_mma.py --only restricted_mma_overlap --reps 2"
$CMD > $OUT/fhmm_list_plain_$TAG.log 2>&1 || { tail -5 $OUT/fhmm_list_plain_$TAG.log; exit 1; }
ncu --metrics gpu__time_duration.sum --clock-control none -c 400 --csv --log-file $OUT/launches_fhmm_$TAG.csv $CMD > $OUT/fhmm_list_ncu_$TAG.log 2>&1
python - "$OUT/launches_fhmm_$TAG.csv" <<'PY'
import csv, re, sys
lines = [l for l in open(sys.argv[1]) if not l.startswith("==")]
seq = []
for row in csv.DictReader(lines):
    try:
        v = float(row["Metric Value"].replace(",", ""))
    except Exception:
        continue
    us = v / 1000 if row["Metric Unit"] in ("ns", "nsecond") else v
    seq.append((re.sub(r"\(.*", "", row["Kernel Name"])[:44], us, row.get("Grid Size"), row.get("Block Size")))
# the last call: everything after the last emission_softmax_partials
last = max(i for i, s in enumerate(seq) if "emission_softmax_partials" in s[0])
tot = 0.0
for s in seq[last:]:
    tot += s[1]
    print(f"{s[0]:44s} {s[1]:8.1f} us  grid {s[2]} block {s[3]}")
print("sum of the call's kernels: %.1f us" % tot)
PY
