#!/bin/bash
# block reductions (ordered + order-free): tests, throughput, ncu of the one-pass kernel; elastic predict capture;
# launch list without look-ahead (serialisable under ncu)
timeout 900 python -m pytest tests/test_gpu_blocks.py -q -p no:cacheprovider -x > gpurun_out/r2p_pytest_blocks.log 2>&1; echo "pytest rc=$?"; tail -4 gpurun_out/r2p_pytest_blocks.log
timeout 600 python tools/run_blocks.py > gpurun_out/r2p_blocks.jsonl 2> gpurun_out/r2p_blocks.err; echo "blocks rc=$?"; cut -c1-900 gpurun_out/r2p_blocks.jsonl; tail -3 gpurun_out/r2p_blocks.err
timeout 600 ncu --set full --clock-control none --import-source on -k regex:vb_block_stream_kernel -s 8 -c 1 -o gpurun_out/r2p_block_stream python tools/run_blocks.py > gpurun_out/r2p_ncu_stream.log 2>&1; echo "ncu stream rc=$?"
timeout 900 ncu --set full --clock-control none --import-source on -k regex:vb_predict2d -s 4 -c 1 -o gpurun_out/r2p_elastic_predict python tools/run_configs.py 4 > gpurun_out/r2p_ncu_elastic_predict.log 2>&1; echo "ncu elastic predict rc=$?"
CMD="python bench.py --steps 1 --warmup 1 --no-e2e --no-cpu-baseline --option chol_lookahead=0"
$CMD > gpurun_out/r2p_bench_nolookahead.json 2> gpurun_out/r2p_bench_nolookahead.err; echo "bench rc=$?"
timeout 900 ncu --metrics gpu__time_duration.sum --clock-control none -c 2000 --csv --log-file gpurun_out/r2p_launches_nolookahead.csv $CMD > gpurun_out/r2p_ncu_list.log 2>&1; echo "launch list rc=$?"
ls -la gpurun_out | grep r2p
