#!/usr/bin/env bash
# C4 spot-check details; C4-shaped e2e (sparse tiles of 100,000-sample rows): timing and the launch list
set -u
TAG="${1:-r2s}"
O=gpurun_out/$TAG; mkdir -p "$O"
timeout 900 python tools/full_configs.py --config C4 --shard 3/8 --spot 2000 --max-sites 94720 > "$O/c4_spot.json" 2> "$O/c4_spot.err"; tail -c 3000 "$O/c4_spot.json"
timeout 600 python tools/e2e_sweep.py --config C4 --sites 18944 --u16 --tiles 9472 --slots 4 --reps 5 2>&1 | tee "$O/e2e_c4.log"
timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none -c 80 --csv --log-file "$O/launches_c4_e2e.csv" \
    python tools/e2e_sweep.py --config C4 --sites 18944 --u16 --tiles 9472 --slots 4 --reps 1 > "$O/e2e_c4_ncu.log" 2>&1
python - <<PY
import csv
rows=[r for r in csv.reader(open("$O/launches_c4_e2e.csv")) if len(r)>5]
hdr=rows[0]; ki=hdr.index("Kernel Name"); vi=hdr.index("Metric Value")
for r in rows[1:]:
    print(r[ki][:50], r[vi])
PY
