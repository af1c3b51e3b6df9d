#!/bin/bash
# Timing experiment (profiles/r01b_summary.md): engine 2 with the split arithmetic compiled out (-DHPL_EXP_NOSPLIT) or the
# gathered loads compiled out (-DHPL_EXP_NOLOAD).  Numbers produced by these builds are wrong by design.
#   here (no GPU):  for v in NOSPLIT NOLOAD; do nvcc -O3 -std=c++17 -gencode arch=compute_100a,code=sm_100a -DHPL_SM_ARCH=100 \
#                     -DHPL_EXP_$v -Xcompiler -fPIC -shared -I include -I hplflownet_b200/csrc -o tools/_exp/libexp_$v.so hplflownet_b200/csrc/*.cu; done
#   on the GPU box: bash tools/_exp/run.sh
for v in BASE NOSPLIT NOLOAD; do
  if [ $v = BASE ]; then unset HPL_LIB_PATH; else export HPL_LIB_PATH=$PWD/tools/_exp/libexp_$v.so; fi
  echo "== $v"
  timeout 120 python tools/try_tma.py time 2>&1 | grep "engine=2"
done
