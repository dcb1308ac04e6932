#!/bin/bash
# Round-2 profiling, third pass (after the scan / epilogue work): launch lists of one training step in both modes and ncu --set full of the kernels
# that changed: the 1x1 forward conv with the statistics epilogue, the split-K coverage-backward GEMM, the fused attention-backward kernel.
set -x
O=gpurun_out
NCU="ncu --clock-control none --kernel-name-base demangled"
ncu --clock-control none --metrics gpu__time_duration.sum --csv --log-file $O/r02_launches_tf32x3.csv python bench.py --steps 1 --warmup 1 --profile-run --no-graphs > $O/r02_prof_a.log 2>&1
ncu --clock-control none --metrics gpu__time_duration.sum --csv --log-file $O/r02_launches_bf16.csv python bench.py --precision bf16 --steps 1 --warmup 1 --profile-run --no-graphs > $O/r02_prof_b.log 2>&1
python profiles/summarize_launches.py $O/r02_launches_tf32x3.csv 0.5 > $O/r02_launches_tf32x3.txt
python profiles/summarize_launches.py $O/r02_launches_bf16.csv 0.5 > $O/r02_launches_bf16.txt
$NCU --set full -k 'regex:tc_gemm_kernel<.int.96, .int.0, .bool.1, .bool.0, .int.1>' -s 70 -c 4 -o $O/r02_x3_c1e python bench.py --steps 1 --warmup 1 --profile-run --no-graphs > $O/r02_prof_k.log 2>&1
$NCU --set full -k 'regex:tc_gemm_kernel<.int.128, .int.0, .bool.1, .bool.0, .int.2>' -s 60 -c 3 -o $O/r02_x3_splitk python bench.py --steps 1 --warmup 1 --profile-run --no-graphs > $O/r02_prof_l.log 2>&1
$NCU --set full -k 'regex:attn_bwd_kernel' -s 50 -c 2 -o $O/r02_attn_bwd python bench.py --steps 1 --warmup 1 --profile-run --no-graphs > $O/r02_prof_m.log 2>&1
for f in x3_c1e x3_splitk attn_bwd; do
  [ -f $O/r02_$f.ncu-rep ] && python profiles/ncu_summary.py $O/r02_$f.ncu-rep > $O/r02_ncu_$f.txt 2>&1
done
rm -f $O/r02_x3_c1e.ncu-rep $O/r02_x3_splitk.ncu-rep $O/r02_attn_bwd.ncu-rep
grep -h "Kernel Name\|gpu__time_duration\|dram__bytes_read" $O/r02_ncu_x3_c1e.txt $O/r02_ncu_x3_splitk.txt $O/r02_ncu_attn_bwd.txt | head -40
