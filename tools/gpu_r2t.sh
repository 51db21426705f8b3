#!/usr/bin/env bash
# parity suite; C4 spot-check details with this build, the round's first commit (head) and the exact Fisher division (ffdiv);
# C4 / C3 / C2-shaped e2e through sparse tiles, launch list of the C4 one
set -u
TAG="${1:-r2t}"
O=gpurun_out/$TAG; mkdir -p "$O"
timeout 1200 python -m pytest tests -m gpu -x -q > "$O/pytest_gpu.log" 2>&1; echo "pytest rc=$?" >> "$O/pytest_gpu.log"; tail -5 "$O/pytest_gpu.log"
for v in "" head ffdiv; do
  L=""; [ -n "$v" ] && L="$PWD/basevar_b200/variants/libbv_$v.so"
  BASEVAR_B200_LIB=$L timeout 900 python tools/full_configs.py --config C4 --shard 3/8 --spot 2000 --max-sites 94720 > "$O/c4_spot_${v:-new}.json" 2> "$O/c4_spot_${v:-new}.err"
  python - <<PY
import json
r=json.load(open("$O/c4_spot_${v:-new}.json"))
print("${v:-new}", {k:r[k] for k in r if k.startswith("spot_") and k!="spot_details"}, [(d["site"], d["kind"], d["cuda"]["chi2"], d["oracle"]["chi2"]) for d in r.get("spot_details",[])])
PY
done
for cfg in "C4 18944 9472" "C3 262144 32768" "C2 1000000 131072"; do
  set -- $cfg
  timeout 600 python tools/e2e_sweep.py --config $1 --sites $2 --u16 --tiles $3 --slots 4 --reps 5 2>&1 | tee -a "$O/e2e.log"
  BASEVAR_B200_LIB=$PWD/basevar_b200/variants/libbv_head.so timeout 600 python tools/e2e_sweep.py --config $1 --sites $2 --u16 --tiles $3 --slots 4 --reps 5 2>&1 | sed 's/^/head: /' | tee -a "$O/e2e.log"
done
timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none -c 40 --csv --log-file "$O/launches_c4_e2e.csv" \
    python tools/e2e_sweep.py --config C4 --sites 18944 --u16 --tiles 9472 --slots 4 --reps 1 > "$O/e2e_c4_ncu.log" 2>&1
grep -c . "$O/launches_c4_e2e.csv"; grep expand "$O/launches_c4_e2e.csv" | tail -3
timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none -c 60 --csv --log-file "$O/launches_c3_e2e.csv" \
    python tools/e2e_sweep.py --config C3 --sites 131072 --u16 --tiles 32768 --slots 4 --reps 1 > "$O/e2e_c3_ncu.log" 2>&1
grep expand "$O/launches_c3_e2e.csv" | tail -2
