#!/bin/bash
# scripts/ab_bench.sh NAME... — A/B timing on the GPU box: runs the default bench (no CPU baseline) against the in-tree library ("tree")
# and against every build/NAME variant (scripts/build_variant.sh), and prints ms/step plus the per-kernel times that differ.
for v in "$@"; do
  if [ "$v" = tree ]; then lib=""; else lib="build/$v/csrc/liboppositerenderer_b200.so"; fi
  OPR_LIBRARY=$lib python bench.py --no-cpu-baseline --repeats 3 $AB_ARGS > gpurun_out/ab_$v.json 2> gpurun_out/ab_$v.err || { echo "$v FAILED"; tail -3 gpurun_out/ab_$v.err; continue; }
  python - "$v" <<'PY'
import json, sys
v = sys.argv[1]
d = json.loads(open("gpurun_out/ab_%s.json" % v).read().strip().splitlines()[-1])
k = d["kernel_ms"]
keys = ["trace.camera0", "trace.photon0", "trace.photon1", "trace.photon2", "trace.photon3", "trace.shadow", "ppm.camera_tail", "ppm.photon_tail", "ppm.photon_late", "ppm.gather", "ppm.photon", "vcm.light", "vcm.camera"]
print("%-10s ms/step %.3f  Mrays/s %.0f  e2e %.0f  early %s | %s" % (v, d["ms_per_step"], d["value"], d["e2e"]["value"], ("%.3f" % d["early_iterations"]["ms_per_step"]) if d.get("early_iterations") else "-",
      "  ".join("%s %.3f" % (x.split(".")[-1] if x.startswith("trace") else x, k[x]) for x in keys if x in k)))
PY
done
