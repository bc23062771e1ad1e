#!/bin/bash
# Round evidence: ncu --set full of every kernel the OF / no-OF paths launch (one small volume), summarised on the box
# (the reports themselves are too large to bring back), plus the launch list. Run only after the same commands have
# exited 0 without ncu.
mkdir -p gpurun_out/ev /tmp/ev
B="python bench.py --shape 20 1024 1024 --steps 1 --warmup 0 --skip-cpu-baseline --skip-e2e --skip-parity"
$B > gpurun_out/ev/plain_of.json 2> gpurun_out/ev/plain_of.err || exit 1
$B --no-of > gpurun_out/ev/plain_noof.json 2>> gpurun_out/ev/plain_of.err || exit 1
timeout 600 ncu --set full --clock-control none -k regex:k_ -c 60 -o /tmp/ev/of_kernels $B > gpurun_out/ev/ncu_of.log 2>&1
python tools/ncu_kernels.py /tmp/ev/of_kernels.ncu-rep > gpurun_out/ev/kernels_of.txt 2>&1
timeout 200 ncu --set full --clock-control none -k regex:k_ -c 3 -o /tmp/ev/noof_exact $B --no-of > gpurun_out/ev/ncu_noof.log 2>&1
python tools/ncu_kernels.py /tmp/ev/noof_exact.ncu-rep > gpurun_out/ev/kernels_noof_exact.txt 2>&1
timeout 200 ncu --set full --clock-control none -k regex:k_ -c 3 -o /tmp/ev/noof_fast $B --no-of --fast-noof >> gpurun_out/ev/ncu_noof.log 2>&1
python tools/ncu_kernels.py /tmp/ev/noof_fast.ncu-rep > gpurun_out/ev/kernels_noof_fast.txt 2>&1
timeout 300 ncu --metrics gpu__time_duration.sum --clock-control none -c 4000 --csv --log-file gpurun_out/ev/launches_of.csv $B > gpurun_out/ev/ncu_launches.log 2>&1
ls -la gpurun_out/ev /tmp/ev; du -sh gpurun_out
