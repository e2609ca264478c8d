#!/bin/bash
# ncu --set full of one launch per hot kernel (1 GPU).  The raw metric pages come back as CSV; two .ncu-rep files are kept.
set -x
export TD_NO_AUTOBUILD=1 PYTHONPATH=$PWD:$PYTHONPATH
mkdir -p gpurun_out/ncu
run() {  # name target kernel-regex [keep-rep]
  timeout 240 ncu --set full --clock-control none --import-source on -k "regex:$3" -s 3 -c 1 -f -o gpurun_out/ncu/$1 python scripts/ncu_targets.py $2 > gpurun_out/ncu/$1.log 2>&1
  echo "$1 rc=$?"
  ncu -i gpurun_out/ncu/$1.ncu-rep --page raw --csv > gpurun_out/ncu/$1.raw.csv 2>/dev/null
  ncu -i gpurun_out/ncu/$1.ncu-rep --page details --csv > gpurun_out/ncu/$1.details.csv 2>/dev/null
  if [ "$4" != "keep" ]; then rm -f gpurun_out/ncu/$1.ncu-rep; fi
}
run gemm_k49152 gemm_k49152 gemm_kernel keep
run gemm_4096 gemm_4096 gemm_kernel
run gemm_rs_shape gemm_rs_shape gemm_kernel
run mxfp8 mxfp8 gemm_kernel keep
run flash flash flash_fwd
run grouped grouped gemm_kernel
run decode decode decode_splitkv
run gemv gemv gemv
run mega mega mega_kernel
ls -la gpurun_out/ncu
