for g in 2 3 4; do GPET_FIT_GROUPS=$g python bench.py --steps 3 --warmup 2 --no-cpu-baseline > gpurun_out/b20_$g.log 2>&1; python - <<PY
import json
d=json.loads([x for x in open("gpurun_out/b20_$g.log") if x.startswith("{")][-1])
print("groups $g:", round(d["value"]),round(d["e2e"]["value"]),round(d["ms_per_step"]),round(d["e2e"]["ms_per_step"]),d["host_ms_last_step"],d["final_fit"])
PY
done
