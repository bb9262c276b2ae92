#!/bin/bash
# End-of-round measurement set on one B200: GPU tests, bench (both arms), ncu launch list of the bench step
O=gpurun_out/r02c; mkdir -p $O
timeout 2400 python -m pytest tests -x -q -m gpu > $O/gputests.log 2>&1; echo "tests rc=$?"; tail -2 $O/gputests.log
python bench.py > $O/bench_1gpu.json 2> $O/bench_1gpu.err; echo "bench rc=$?"
python bench.py --impl reference --steps 3 --warmup 1 > $O/bench_reference.json 2> $O/bench_reference.err; echo "ref rc=$?"
ncu --metrics gpu__time_duration.sum --clock-control none -c 400 --csv --log-file $O/launches_bench.csv python bench.py --steps 3 --warmup 3 --configs none --msa-leaves 0 --no-cpu-baseline > $O/bench_under_ncu.json 2> /dev/null; echo "ncu rc=$?"
python - <<PY
import json
d=json.load(open("$O/bench_1gpu.json"))
print({k:d[k] for k in ("metric","value","ms_per_step","gpu_launches")}, "e2e", round(d["e2e"]["value"],1), "frac", round(d["roofline"]["frac"],4), d["clocks"])
for k,v in d.get("configs",{}).items():
    print(k, {x:(round(v[x],2) if isinstance(v[x],float) else v[x]) for x in ("wall_s","device_ms","gcups_device","gcups_dp_phase","gcups_e2e","level_calls_wall_ms","byte_identical_to_reference") if x in v})
r=json.load(open("$O/bench_reference.json")); print("reference", r.get("value"), r.get("unit"), r.get("cpu_baseline",{}).get("cores"))
PY
