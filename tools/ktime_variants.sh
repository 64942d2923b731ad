#!/bin/bash
# Per-iteration kernel-class times (RVE_KTIME) of librve_b200.so and of every variant library under
# deepbnd_b200/lib/variants/ on a short C2 run; used for kernel-variant experiments on the GPU box.
NS=${NS:-1000}
for lib in deepbnd_b200/lib/librve_b200.so deepbnd_b200/lib/variants/*.so; do
  [ -f "$lib" ] || continue
  DEEPBND_RVE_LIB=$PWD/$lib RVE_KTIME=1 python bench.py --steps 1 --warmup 3 --ns $NS --chunks 1 --no-cpu-baseline --unique 64 $EXTRA > gpurun_out/kv.json 2> gpurun_out/kv.err
  echo "$(basename $lib): $(grep rve_solve gpurun_out/kv.err | tail -1 | sed 's/.*per iteration (us)//')" $(python - <<'PY'
import json
d = json.loads(open("gpurun_out/kv.json").read().strip().splitlines()[-1])
print("value", round(d["value"], 1), "its", d["config"]["mean_pcg_iters"])
PY
)
done
