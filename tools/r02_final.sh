#!/bin/bash
# Round-2 final measurement pass (1 GPU): whole GPU suite, then every bench workload.
O=gpurun_out
python -m pytest tests -m gpu -q 2>&1 | tail -15 > $O/r02_pytest_gpu_final.log
tail -3 $O/r02_pytest_gpu_final.log
python bench.py --steps 20 --warmup 5 > $O/r02_bench_n1_tf32x3.json 2> $O/r02_bench_n1_tf32x3.err
python bench.py --impl reference --steps 3 --warmup 1 > $O/r02_bench_reference.json 2> $O/r02_bench_reference.err
python bench.py --precision bf16 --steps 20 --warmup 5 --no-cpu-baseline > $O/r02_bench_n1_bf16.json 2> $O/r02_bench_n1_bf16.err
python bench.py --workload cfg2b16 --steps 20 --warmup 5 --no-cpu-baseline > $O/r02_bench_cfg2b16_tf32x3.json 2> $O/r02_bench_cfg2b16_tf32x3.err
python bench.py --workload cfg2b16 --precision bf16 --steps 20 --warmup 5 --no-cpu-baseline > $O/r02_bench_cfg2b16_bf16.json 2> $O/r02_bench_cfg2b16_bf16.err
python bench.py --workload cfg3 --steps 20 --warmup 3 --no-cpu-baseline > $O/r02_bench_cfg3_bf16.json 2> $O/r02_bench_cfg3_bf16.err
python bench.py --workload cfg3 --precision tf32x3 --steps 20 --warmup 3 --no-cpu-baseline > $O/r02_bench_cfg3_tf32x3.json 2> $O/r02_bench_cfg3_tf32x3.err
python bench.py --workload cfg5 --images 1000 --oracle-images 10 > $O/r02_bench_cfg5.json 2> $O/r02_bench_cfg5.err
python tools/prof_cluster.py 1 > $O/r02_cluster_stages.txt 2>&1
python tools/prof_decode.py 8 > $O/r02_persist_phases.txt 2>&1
python tools/seg_time.py > $O/r02_segments.txt 2>&1
python tools/seg_time.py bf16 >> $O/r02_segments.txt 2>&1
for f in n1_tf32x3 n1_bf16 cfg2b16_tf32x3 cfg2b16_bf16 cfg3_bf16 cfg3_tf32x3 cfg5; do python -c "
import json; d=json.load(open('$O/r02_bench_$f.json')); print('$f', round(d['value'],1), d['unit'], round(d['ms_per_step'],3), 'e2e', round(d['e2e']['value'],1))"; done
