#!/bin/bash
# A/B session: lab timing + short bench for every variant library under labso/
mkdir -p gpurun_out/s
for v in product $(ls labso | sed 's/\.so$//'); do
  if [ $v = product ]; then unset FDN_LIB_PATH; else export FDN_LIB_PATH=labso/$v.so; fi
  timeout 120 python tools/flow_iter_lab.py --n 256 --reps 3 2>&1 | tail -1
  timeout 200 python bench.py --steps 1 --warmup 1 --skip-cpu-baseline --skip-e2e 2>/dev/null | python -c "
import json,sys; d=json.loads(sys.stdin.read()); r=d['roofline']; print('   bench', round(d['ms_per_step'],1), 'ms; flow_iter', r['kernel_ms_per_step']['k_flow_iter'], 'frac', round(r['frac'],3), {k:v['frac_of_peak'] for k,v in r['flow_iter_by_level'].items()})"
done 2>&1 | tee gpurun_out/s/ab.log
