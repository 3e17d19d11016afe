#!/bin/bash
mkdir -p gpurun_out
python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29511 bench.py --gpus 2 --steps 5 --warmup 3 --no-cpu-baseline > gpurun_out/bench_n2.json 2> gpurun_out/bench_n2.err; echo rc=$?; tail -3 gpurun_out/bench_n2.err | cut -c1-300; cut -c1-700 gpurun_out/bench_n2.json
python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29512 tools/ddp_parity.py > gpurun_out/ddp_parity.log 2>&1; echo rc=$?; tail -5 gpurun_out/ddp_parity.log
