#!/bin/bash
# Round-2 profiling pass (run through gpurun): launch lists + ncu --set full captures, summarised on the box into gpurun_out/.
set -x
O=gpurun_out
NCU="ncu --clock-control none"
# 1) launch lists of ONE training step (stream launches so that every kernel is its own launch), tf32x3 and bf16
$NCU --metrics gpu__time_duration.sum --csv --log-file $O/r02_launches_tf32x3.csv python bench.py --steps 1 --warmup 1 --profile-run --no-graphs > $O/r02_prof_a.log 2>&1
$NCU --metrics gpu__time_duration.sum --csv --log-file $O/r02_launches_bf16.csv python bench.py --precision bf16 --steps 1 --warmup 1 --profile-run --no-graphs > $O/r02_prof_b.log 2>&1
python profiles/summarize_launches.py $O/r02_launches_tf32x3.csv 0.5 > $O/r02_launches_tf32x3.txt
python profiles/summarize_launches.py $O/r02_launches_bf16.csv 0.5 > $O/r02_launches_bf16.txt
# 2) ncu --set full: cluster greedy kernel, grid-wide persistent kernel (B=8), beam kernel
$NCU --set full --import-source on -k regex:greedy_cluster_kernel -c 1 -o $O/r02_cluster python tools/prof_cluster.py 1 > $O/r02_prof_c.log 2>&1
$NCU --set full -k regex:greedy_persistent_kernel -c 1 -o $O/r02_persist python tools/prof_decode.py > $O/r02_prof_d.log 2>&1
# 3) bf16 and 3xTF32 conv kernels of a B=16 training step (steady state: skip the first launches of each)
$NCU --set full -k 'regex:tc_gemm_kernel<96, 0, 0, 1>' -s 40 -c 6 -o $O/r02_bf16_c1 python bench.py --workload cfg2b16 --precision bf16 --steps 1 --warmup 1 --profile-run --no-graphs > $O/r02_prof_e.log 2>&1
$NCU --set full -k 'regex:tc_gemm_kernel<96, 2, 0, 1>' -s 40 -c 6 -o $O/r02_bf16_c3 python bench.py --workload cfg2b16 --precision bf16 --steps 1 --warmup 1 --profile-run --no-graphs > $O/r02_prof_f.log 2>&1
$NCU --set full -k 'regex:tc_wgrad_bf16_kernel' -s 60 -c 8 -o $O/r02_bf16_wg python bench.py --workload cfg2b16 --precision bf16 --steps 1 --warmup 1 --profile-run --no-graphs > $O/r02_prof_g.log 2>&1
$NCU --set full -k 'regex:tc_gemm_kernel<96, 0, 1, 0>' -s 40 -c 6 -o $O/r02_x3_c1 python bench.py --workload cfg2b16 --steps 1 --warmup 1 --profile-run --no-graphs > $O/r02_prof_h.log 2>&1
$NCU --set full -k 'regex:tc_gemm_kernel<128, 0, 1, 0>' -s 60 -c 6 -o $O/r02_x3_dg python bench.py --steps 1 --warmup 1 --profile-run --no-graphs > $O/r02_prof_i.log 2>&1
for f in cluster persist bf16_c1 bf16_c3 bf16_wg x3_c1 x3_dg; do
  [ -f $O/r02_$f.ncu-rep ] && python profiles/ncu_summary.py $O/r02_$f.ncu-rep > $O/r02_ncu_$f.txt 2>&1
done
ncu -i $O/r02_cluster.ncu-rep --page source --csv > $O/r02_cluster_source.csv 2>/dev/null
ls -la $O/*.ncu-rep
rm -f $O/r02_bf16_c1.ncu-rep $O/r02_bf16_c3.ncu-rep $O/r02_bf16_wg.ncu-rep $O/r02_x3_c1.ncu-rep $O/r02_x3_dg.ncu-rep $O/r02_persist.ncu-rep
du -sh $O
