#!/bin/bash
mkdir -p gpurun_out
echo "== bench default"; timeout 900 python bench.py > gpurun_out/bench_r1c.json 2> gpurun_out/bench_r1c.err; echo rc=$?; tail -2 gpurun_out/bench_r1c.err
timeout 1200 python bench.py --impl reference --steps 3 --warmup 1 > gpurun_out/bench_r1c_ref.json 2> gpurun_out/bench_ref.err; echo ref rc=$?
B="python bench.py --steps 2 --warmup 1 --no-cpu-baseline --e2e-format csr"
echo "== launch list"; timeout 900 ncu --metrics gpu__time_duration.sum --clock-control none -c 1200 --csv --log-file gpurun_out/launches_r1c.csv $B > gpurun_out/launches_r1c.log 2>&1; echo rc=$?
for k in k_bits_pull_mid k_bits_fill_rows k_bits_pull_small; do
  echo "== full: $k"; timeout 900 ncu --set full --clock-control none --import-source on -k "regex:^$k\$" -s 1 -c 1 -f -o gpurun_out/prof_r1c_$k $B > gpurun_out/p_$k.log 2>&1; echo rc=$?
done
for k in k_bits_count k_bits_pull_long k_bits_push; do
  echo "== summary metrics: $k"; timeout 900 ncu --clock-control none -k "regex:^$k\$" -s 1 -c 1 --metrics gpu__time_duration.sum,dram__bytes_read.sum,dram__bytes_write.sum,gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed,smsp__inst_executed.sum,smsp__issue_active.avg.pct_of_peak_sustained_active,sm__warps_active.avg.pct_of_peak_sustained_active,l1tex__throughput.avg.pct_of_peak_sustained_elapsed,lts__throughput.avg.pct_of_peak_sustained_elapsed,lts__t_sector_hit_rate.pct,launch__registers_per_thread,launch__grid_size $B 2>&1 | grep -E "k_bits|gpu__|dram__|smsp__|sm__|l1tex__|lts__|launch__" > gpurun_out/sum_$k.txt; cat gpurun_out/sum_$k.txt | head -20
done
du -sh gpurun_out
