#!/bin/bash
# A/B session: times tools/noof_lab.py for the product library and every variant under labso/
mkdir -p gpurun_out/s
for v in product $(ls labso | sed 's/\.so$//'); do
  if [ $v = product ]; then unset FDN_LIB_PATH; else export FDN_LIB_PATH=labso/$v.so; fi
  timeout 120 python tools/noof_lab.py 2>&1 | tail -1
done 2>&1 | tee gpurun_out/s/noof_ab.log
