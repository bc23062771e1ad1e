#!/bin/bash
# A/B session: times tools/flow_iter_lab.py for the product library and every variant under labso/
mkdir -p gpurun_out/s
for v in product $(ls labso | sed 's/\.so$//'); do
  if [ $v = product ]; then unset FDN_LIB_PATH; else export FDN_LIB_PATH=labso/$v.so; fi
  for args in "$@"; do
    timeout 120 python tools/flow_iter_lab.py $args --check 2>&1 | tail -2 | tr '\n' ' '; echo
  done
done 2>&1 | tee gpurun_out/s/ab.log
