#!/bin/bash
# A/B of library builds on the GPU box:  tools/sweep_variants.sh "C2 C2m C5" tools/bin/libslf_x.so ...
# (variants are linked by hand from objects compiled with -DSLF_MINB_*=N etc.; SLF_B200_LIB selects one;
#  tools/bin/ travels to the box, gpurun_out/ comes back).  The default build runs last.
configs="$1"; shift
mkdir -p gpurun_out; out=gpurun_out/sweep_variants.txt; : > $out
for lib in "$@" ""; do
  echo "== ${lib:-default build}" >> $out
  SLF_B200_LIB=${lib:+$PWD/$lib} timeout 300 python tools/bench_configs.py $configs >> $out 2>&1
done
cat $out
