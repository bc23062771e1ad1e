#!/bin/bash
# A/B of whole-pass kernels: python bench.py on a reduced volume for the product library and every variant under labso/;
# prints ms per step, one kernel's time / fraction and the result hash (must be equal across variants).
# usage: bash tools/gpu_ab_bench.sh KERNEL_KEY [bench args...]
key=$1; shift
mkdir -p gpurun_out/ab
for v in product $(ls labso 2>/dev/null | sed 's/\.so$//'); do
  if [ $v = product ]; then unset FDN_LIB_PATH; else export FDN_LIB_PATH=labso/$v.so; fi
  python bench.py "$@" --skip-cpu-baseline --skip-e2e --skip-parity > gpurun_out/ab/$v.json 2> gpurun_out/ab/$v.err
  python - $v $key <<'PY'
import json, sys
v, key = sys.argv[1:3]
try:
    d = json.loads(open(f"gpurun_out/ab/{v}.json").read().strip().splitlines()[-1])
    r = d["roofline"]
    print(v, "ms/step", round(d["ms_per_step"], 1), key, r["kernel_ms_per_step"].get(key), r["kernel_frac_of_peak"].get(key),
          "hash", d["identity"]["result_sha256"][:12])
except Exception as e:
    print(v, "ERR", e)
PY
done
