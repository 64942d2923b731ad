#!/bin/bash
# e2e sensitivity to the number of chunks per step (C2 default workload)
mkdir -p gpurun_out
for c in ${CHUNKS:-1 2}; do
  python bench.py --steps ${STEPS:-4} --warmup 3 --no-cpu-baseline --no-extra --chunks $c > gpurun_out/e2e_c$c.json 2> gpurun_out/e2e_c$c.err
  python - <<PY
import json
d = json.loads(open("gpurun_out/e2e_c$c.json").read().strip().splitlines()[-1])
print("chunks $c: value %.0f e2e %.0f ratio %.3f" % (d["value"], d["e2e"]["value"], d["e2e"]["value"] / d["value"]))
PY
done
