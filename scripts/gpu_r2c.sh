#!/bin/bash
# round 2, call C: full-size parity, one test per process, bounded in time and host memory
set -x
mkdir -p gpurun_out
make -C falkordb_b200/csrc -j16 -s 2>&1 | tail -3; make -C oracle -s
timeout 300 python -m pytest tests/test_gpu_parity.py tests/test_fp64.py -m gpu -x -q -k "hot_set_order or connected_components" > gpurun_out/r2c_pytest.log 2>&1; tail -3 gpurun_out/r2c_pytest.log
: > gpurun_out/r2c_fullsize.log
for t in "test_chain_rmat24_all_rows" "test_config3_ldbc_sf10_shaped_chain" "test_config4_masked_triangles_rmat24" "test_config2_single_mxm_rmat22" "test_config5_bfs_rmat26_levels_and_parents"; do
  ( time timeout 600 python -m pytest tests/test_zz_full_size.py -m gpu -x -q -k "$t" --durations=0 ) >> gpurun_out/r2c_fullsize.log 2>&1
  tail -4 gpurun_out/r2c_fullsize.log
  free -g | head -2 | tail -1
done
