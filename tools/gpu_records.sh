#!/bin/bash
# Round records on one B200: the GPU test suite, then one bench line per BASELINE.json configuration (N = 1).
mkdir -p gpurun_out/rec
(time python -m pytest tests -m gpu -x -q) > gpurun_out/rec/pytest_gpu.log 2>&1; tail -4 gpurun_out/rec/pytest_gpu.log
python bench.py > gpurun_out/rec/bench_cfg2_n1.json 2> gpurun_out/rec/bench_cfg2_n1.err
python bench.py --config cfg1 > gpurun_out/rec/bench_cfg1_n1.json 2> gpurun_out/rec/bench_cfg1_n1.err
python bench.py --config cfg3 > gpurun_out/rec/bench_cfg3_noof_exact.json 2> gpurun_out/rec/bench_cfg3.err
python bench.py --config cfg3 --fast-noof > gpurun_out/rec/bench_cfg3_noof_fast.json 2>> gpurun_out/rec/bench_cfg3.err
python bench.py --config cfg4 --steps 2 --warmup 1 > gpurun_out/rec/bench_cfg4_n1.json 2> gpurun_out/rec/bench_cfg4_n1.err
python bench.py --recompute-flow --steps 1 --warmup 1 --cpu-slices 4 > gpurun_out/rec/bench_cfg2_recompute_n1.json 2> gpurun_out/rec/bench_cfg2_recompute_n1.err
python -c "import __graft_entry__ as g; g.smoke()" > gpurun_out/rec/smoke.log 2>&1; tail -2 gpurun_out/rec/smoke.log
for f in gpurun_out/rec/*.json; do python - "$f" <<'PY'
import json,sys
try:
    d=json.loads(open(sys.argv[1]).read().strip().splitlines()[-1])
    print(sys.argv[1].split('/')[-1], round(d['value'],1), d['unit'], 'ms', round(d['ms_per_step'],2), 'e2e', round(d['e2e']['value'],1), 'frac', round(d['roofline']['frac'],3), 'parity', (d.get('parity') or {}).get('bit_equal'), 'hash_ok', (d.get('identity') or {}).get('equals_n1'))
except Exception as e:
    print(sys.argv[1], 'ERR', e)
PY
done
