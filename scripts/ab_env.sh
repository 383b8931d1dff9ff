#!/bin/bash
# scripts/ab_env.sh "ENV=1 ENV2=2" ... — A/B timing of run-time switches on the GPU box: one default bench (no CPU baseline) per argument
for e in "$@"; do
  tag=$(echo "$e" | tr ' =' '__')
  env $e python bench.py --no-cpu-baseline --repeats 3 $AB_ARGS > gpurun_out/abenv_$tag.json 2> gpurun_out/abenv_$tag.err || { echo "$e FAILED"; tail -3 gpurun_out/abenv_$tag.err; continue; }
  python - "$e" "gpurun_out/abenv_$tag.json" <<'PY'
import json, sys
d = json.loads(open(sys.argv[2]).read().strip().splitlines()[-1])
k = d["kernel_ms"]
keys = ["trace.camera0", "trace.photon0", "trace.photon1", "trace.photon2", "trace.photon3", "trace.shadow", "ppm.camera_tail", "ppm.photon_tail", "ppm.photon_late", "ppm.gather", "vcm.light", "vcm.camera"]
print("%-40s ms/step %.3f  Mrays/s %.0f  e2e %.0f  early %s | %s" % (sys.argv[1], d["ms_per_step"], d["value"], d["e2e"]["value"], ("%.3f" % d["early_iterations"]["ms_per_step"]) if d.get("early_iterations") else "-",
      "  ".join("%s %.3f" % (x.split(".")[-1] if x.startswith("trace") else x, k[x]) for x in keys if x in k)))
PY
done
