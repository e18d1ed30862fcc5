run() { name=$1; shift; env "$@" timeout 400 python bench.py --steps 2 --warmup 3 --no-e2e --no-cpu-baseline > gpurun_out/exp16_$name.json 2> gpurun_out/exp16_$name.err; python - <<PY
import json
try:
    d=json.loads(open("gpurun_out/exp16_$name.json").read().strip().splitlines()[-1])
    print("$name", "ms/step", round(d["ms_per_step"],1), "sweep", round(d["roofline"]["sweep_ms_per_step"],1), "frac", round(d["roofline"]["frac"],4), "launches", d["gpu_launches"])
except Exception as e:
    print("$name failed", e)
PY
}
run dirs1 VRT_MAX_DIRS=1
run runlen16 VRT_RUN_LEN=16
