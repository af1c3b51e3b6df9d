#!/bin/bash
# timing experiment: engine 2 with the split arithmetic removed / the gathered loads removed (numbers are wrong by design)
for v in BASE NOSPLIT NOLOAD; do
  if [ $v = BASE ]; then unset HPL_LIB_PATH; else export HPL_LIB_PATH=$PWD/tools/_exp/libexp_$v.so; fi
  echo "== $v"
  timeout 120 python tools/try_tma.py time 2>&1 | grep "engine=2"
done
