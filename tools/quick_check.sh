#!/bin/bash
# Quick iteration run on one B200 (under gpurun): GPU parity tests, a short bench of the ensemble kernel and of the single
# trial, and the in-kernel stage timers.  usage: tools/quick_check.sh TAG [pytest -k expression]
TAG=${1:-q}
O=gpurun_out
mkdir -p $O
python -m pytest tests -m gpu -x -q ${2:+-k "$2"} > $O/pytest_$TAG.log 2>&1; echo "pytest rc $?" ; tail -3 $O/pytest_$TAG.log
python bench.py --steps 6 --warmup 3 --no-cpu --no-configs --no-peak > $O/bench_$TAG.json 2> $O/bench_$TAG.err; echo "bench rc $?"
python - <<PY
import json
d = json.loads(open("$O/bench_$TAG.json").read().strip().splitlines()[-1])
print("ms_per_step", d["ms_per_step"], "value", d["value"], "single", d.get("single_trial"))
PY
python tools/stage_profile.py 148 3 > $O/stage_profile_$TAG.txt 2>&1
RAAE_CTAS_PER_TRIAL=8 python tools/stage_profile.py 1 3 > $O/stage_profile_${TAG}_c8.txt 2>&1
head -18 $O/stage_profile_$TAG.txt | cut -c1-90
tail -1 $O/stage_profile_$TAG.txt
head -1 $O/stage_profile_${TAG}_c8.txt
