for cfg in "4 2 2" "2 2 2" "2 2 1" "3 3 3" "6 3 3"; do set -- $cfg; python bench.py --steps 3 --warmup 3 --no-cpu-baseline --sub-batches $1 --window $2 --fit-merge $3 > gpurun_out/b22_$1_$2_$3.log 2>&1; python - <<PY
import json
l=[x for x in open("gpurun_out/b22_$1_$2_$3.log") if x.startswith("{")]
d=json.loads(l[-1]) if l else None
print("sub/win/merge $cfg:", d and (round(d["value"]),round(d["e2e"]["value"]),round(d["ms_per_step"]),round(d["e2e"]["ms_per_step"]),d["host_ms_last_step"]["fit_rounds"]))
PY
done
