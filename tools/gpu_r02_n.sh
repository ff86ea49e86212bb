#!/bin/bash
# round-2 call N (1 GPU): one-block-ahead loads in the filter / delta-filter kernels (parity + A/B), and an old-vs-new
# library A/B of the bandwidth ops that read lower in bench_r02_m than in bench_r02_h
mkdir -p gpurun_out
for mode in "FLB_FILTER_PIPE=1" "FLB_FILTER_PIPE=8" "FLB_DELTA_FILTER_NB=4" "FLB_DELTA_FILTER_NB=8"; do
  env $mode timeout 600 python -m pytest tests/test_gpu_scan.py -x -q -m gpu -k "filter" > gpurun_out/pytest_gpu_n.log 2>&1; echo "pytest [$mode] exit $?"; tail -1 gpurun_out/pytest_gpu_n.log
done
for mode in "X=0" "FLB_FILTER_PIPE=1" "FLB_FILTER_PIPE=8"; do echo "== unpack_filter [$mode]"; env $mode timeout 300 python tools/opbench.py unpack_filter 2>&1 | tee gpurun_out/opbench_filter_n_${mode#*=}.txt; done
for mode in "FLB_DELTA_FILTER_NB=1" "FLB_DELTA_FILTER_NB=4" "FLB_DELTA_FILTER_NB=8"; do echo "== undelta_pack_filter [$mode]"; env $mode timeout 300 python tools/opbench.py undelta_pack_filter 2>&1 | tee gpurun_out/opbench_dfilter_n_${mode#*=}.txt; done
OPS=undelta_pack_untranspose,transpose,for_pack,undelta_pack,pack
echo "== new lib"; timeout 300 python tools/opbench.py $OPS --types 16,32,64 2>&1 | grep -E "W=1 |W=0 " | tee gpurun_out/opbench_ab_new.txt
cp fastlanes_b200/lib/libfastlanes_b200.so /tmp/new.so; cp build/lib_old/libfastlanes_b200.so fastlanes_b200/lib/libfastlanes_b200.so
echo "== old lib (5c75625)"; timeout 300 python tools/opbench.py $OPS --types 16,32,64 2>&1 | grep -E "W=1 |W=0 " | tee gpurun_out/opbench_ab_old.txt
cp /tmp/new.so fastlanes_b200/lib/libfastlanes_b200.so
echo "== new lib again"; timeout 300 python tools/opbench.py $OPS --types 16,32,64 2>&1 | grep -E "W=1 |W=0 " | tee gpurun_out/opbench_ab_new2.txt
