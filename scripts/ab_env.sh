#!/bin/bash
# Developer A/B: bench.py --quick under several environment settings.  usage: scripts/ab_env.sh TAG "ENV1=.. ENV2=.." "ENV..." ...
tag=$1; shift
i=0
for e in "$@"; do
  i=$((i+1))
  env $e python bench.py --quick --steps 10 --warmup 3 $AB_ARGS > gpurun_out/${tag}_$i.json 2> gpurun_out/${tag}_$i.err
  python - "$e" gpurun_out/${tag}_$i.json <<'PY'
import json, sys
try:
    d = json.loads(open(sys.argv[2]).read().strip().splitlines()[-1])
    print(f"{sys.argv[1]!r:50s} step {d['ms_per_step']:.3f} ms stages {[round(v,3) for v in d['stages_ms']]} launches {d['launches_per_step']}")
except Exception as ex:
    print(sys.argv[1], "FAILED", ex)
PY
done
