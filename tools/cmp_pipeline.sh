for cfg in "1 1 1 0" "4 2 2 0" "6 2 2 0" "4 2 2 1" "8 2 2 0"; do set -- $cfg; python bench.py --steps 2 --warmup 1 --no-cpu-baseline --sub-batches $1 --window $2 --fit-merge $3 --fit-server $4 > gpurun_out/b9_$1_$2_$3_$4.log 2>&1; python - <<PY
import json
l=[x for x in open("gpurun_out/b9_$1_$2_$3_$4.log") if x.startswith("{")]
d=json.loads(l[-1]) if l else None
print("sub/win/merge/server $cfg:", d and (round(d["value"]),round(d["e2e"]["value"]),round(d["ms_per_step"]),round(d["roofline"]["frac"],3),d["host_ms_last_step"]))
PY
done
