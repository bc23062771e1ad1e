#!/bin/bash
mkdir -p gpurun_out/s2
for v in new l2pf1 l2pf2 midfence l2pf1mid l1pfs; do
  FDN_LIB_PATH=labso/$v.so timeout 300 python tools/flow_iter_lab.py --n 128 --check >> gpurun_out/s2/lab.log 2>&1
done
cat gpurun_out/s2/lab.log
