#!/bin/bash
# two-GPU call: native NCCL gather vs torch gather, bench at N = 2 (TV-L1 headline only to keep it short), multi-device test
mkdir -p gpurun_out
timeout 300 python -m pytest tests/test_adjacent_gpu.py -x -q -m gpu -k "device" > gpurun_out/r2n2_pytest.log 2>&1; tail -3 gpurun_out/r2n2_pytest.log
for g in native torch; do
  timeout 900 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29517 bench.py --gpus 2 --steps 5 --warmup 3 --workload tvl1 --gather $g --no-cpu --no-extras > gpurun_out/r2n2_bench_$g.json 2> gpurun_out/r2n2_bench_$g.err
  echo "gather=$g rc=$?"; tail -c 900 gpurun_out/r2n2_bench_$g.json; tail -3 gpurun_out/r2n2_bench_$g.err
done
timeout 900 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29518 bench.py --gpus 2 --steps 3 --warmup 3 > gpurun_out/r2n2_bench_default.json 2> gpurun_out/r2n2_bench_default.err
echo "default rc=$?"; tail -c 1200 gpurun_out/r2n2_bench_default.json; tail -3 gpurun_out/r2n2_bench_default.err
