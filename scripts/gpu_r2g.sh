#!/bin/bash
# round 2, call G: the evidence run -- final bench lines, launch lists, one ncu --set full capture per kernel (raw CSV pages only:
# the .ncu-rep files stay on the box, gpurun_out/ is capped at 64 MiB)
set -x
mkdir -p gpurun_out
make -C falkordb_b200/csrc -j16 -s 2>&1 | tail -3; make -C oracle -s
timeout 900 python bench.py --steps 10 --warmup 3 > gpurun_out/r2g_bench_chain_n1.json 2> gpurun_out/r2g_bench_chain_n1.err; tail -c 600 gpurun_out/r2g_bench_chain_n1.json
timeout 900 python bench.py --impl reference --steps 5 --warmup 1 > gpurun_out/r2g_bench_reference.json 2> gpurun_out/r2g_bench_reference.err; tail -c 400 gpurun_out/r2g_bench_reference.json
BQ="python bench.py --steps 2 --warmup 1 --no-cpu-baseline --e2e-format csr"
timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none -c 1500 --csv --log-file gpurun_out/r2g_launches.csv $BQ > gpurun_out/r2g_launches_bench.log 2>&1
cap() {  # cap <kernel-regex> <tag> <skip> <cmd...>
  k=$1; tag=$2; skip=$3; shift 3
  timeout 500 ncu --set full --clock-control none --import-source on -k regex:"$k" -s $skip -c 1 -o /tmp/r2g_$tag -f "$@" > gpurun_out/r2g_ncu_$tag.log 2>&1
  ncu -i /tmp/r2g_$tag.ncu-rep --page raw --csv > gpurun_out/r2g_raw_$tag.csv 2>/dev/null
  ncu -i /tmp/r2g_$tag.ncu-rep --page source --csv 2>/dev/null | cut -d, -f1-12 | head -2500 > gpurun_out/r2g_src_$tag.csv
  rm -f /tmp/r2g_$tag.ncu-rep
}
cap "^k_pull_seg" pull_seg 3 $BQ
cap "^k_pull_small" pull_small 3 $BQ
cap "^k_bits_fill_v3" fill_v3 3 $BQ
cap "^k_bits_count_csa" count_csa 3 $BQ
cap "^k_csr_push" csr_push 3 $BQ
cap "^k_ordered_flops" ordered_flops 2 $BQ
cap "^k_union_fill" union_fill 1 python bench.py --workload delta --steps 1 --warmup 1
cap "^k_rowfilter_fill" rowfilter_fill 2 python bench.py --workload delta --steps 1 --warmup 1
cap "^k_masked_pairs" masked_pairs 1 python bench.py --workload triangles --scale 22 --steps 1 --warmup 1 --tri-parity 0
cap "^k_mxv_fp64" mxv_fp64 5 python bench.py --workload pagerank --steps 1 --warmup 1
cap "^k_do_pull" bfs_pull 1 python bench.py --workload bfs --scale 24 --bfs-sources 1 --warmup 1 --bfs-parity 0
cap "^k_do_expand" bfs_expand 6 python bench.py --workload bfs --scale 24 --bfs-sources 1 --warmup 1 --bfs-parity 0
timeout 400 ncu --metrics gpu__time_duration.sum --clock-control none -c 3000 --csv --log-file gpurun_out/r2g_bfs_launches.csv python bench.py --workload bfs --scale 26 --bfs-sources 2 --warmup 1 --bfs-parity 0 > gpurun_out/r2g_bfs_launches.log 2>&1
timeout 400 python bench.py --workload delta --steps 3 --warmup 1 > gpurun_out/r2g_delta.json 2> gpurun_out/r2g_delta.err
timeout 400 python bench.py --workload pagerank --steps 3 --warmup 1 > gpurun_out/r2g_pagerank.json 2> gpurun_out/r2g_pagerank.err
timeout 600 python bench.py --workload triangles --scale 24 --steps 3 --warmup 1 > gpurun_out/r2g_tri_s24_n1.json 2> gpurun_out/r2g_tri_s24_n1.err; tail -c 800 gpurun_out/r2g_tri_s24_n1.json
timeout 400 python bench.py --workload bfs --scale 26 --bfs-sources 16 --warmup 2 --bfs-parity 2 > gpurun_out/r2g_bfs_s26_n1.json 2> gpurun_out/r2g_bfs_s26_n1.err; tail -c 800 gpurun_out/r2g_bfs_s26_n1.json
du -sh gpurun_out; ls -la gpurun_out | grep r2g | head -60
