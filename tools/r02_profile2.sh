#!/bin/bash
# Round-2 profiling, second pass: ncu --set full of the conv kernels by DEMANGLED name (template arguments), the grid-wide decoder after the
# stencil rewrite, and the bf16 GEMM; summaries only.
set -x
O=gpurun_out
NCU="ncu --clock-control none --kernel-name-base demangled"
$NCU --set full -k 'regex:tc_gemm_kernel<.int.96, .int.0, .bool.0, .bool.1>' -s 40 -c 6 -o $O/r02_bf16_c1 python bench.py --workload cfg2b16 --precision bf16 --steps 1 --warmup 1 --profile-run --no-graphs > $O/r02_prof_e.log 2>&1
$NCU --set full -k 'regex:tc_gemm_kernel<.int.96, .int.2, .bool.0, .bool.1>' -s 40 -c 6 -o $O/r02_bf16_c3 python bench.py --workload cfg2b16 --precision bf16 --steps 1 --warmup 1 --profile-run --no-graphs > $O/r02_prof_f.log 2>&1
$NCU --set full -k 'regex:tc_gemm_kernel<.int.96, .int.0, .bool.1, .bool.0>' -s 40 -c 6 -o $O/r02_x3_c1 python bench.py --workload cfg2b16 --steps 1 --warmup 1 --profile-run --no-graphs > $O/r02_prof_h.log 2>&1
$NCU --set full -k 'regex:tc_gemm_kernel<.int.96, .int.2, .bool.1, .bool.0>' -s 40 -c 6 -o $O/r02_x3_c3 python bench.py --workload cfg2b16 --steps 1 --warmup 1 --profile-run --no-graphs > $O/r02_prof_j.log 2>&1
$NCU --set full -k 'regex:tc_gemm_kernel<.int.128, .int.0, .bool.1, .bool.0>' -s 60 -c 6 -o $O/r02_x3_dg python bench.py --steps 1 --warmup 1 --profile-run --no-graphs > $O/r02_prof_i.log 2>&1
$NCU --set full -k regex:greedy_persistent_kernel -c 1 -o $O/r02_persist python tools/prof_decode.py > $O/r02_prof_d.log 2>&1
for f in bf16_c1 bf16_c3 x3_c1 x3_c3 x3_dg persist; do
  [ -f $O/r02_$f.ncu-rep ] && python profiles/ncu_summary.py $O/r02_$f.ncu-rep > $O/r02_ncu_$f.txt 2>&1
done
rm -f $O/r02_bf16_c1.ncu-rep $O/r02_bf16_c3.ncu-rep $O/r02_x3_c1.ncu-rep $O/r02_x3_c3.ncu-rep $O/r02_x3_dg.ncu-rep $O/r02_persist.ncu-rep $O/r02_cluster.ncu-rep
grep -h "Kernel Name" $O/r02_ncu_*.txt | sort | uniq -c
