#!/bin/bash
# (2 GPUs) graph replay of the multi-GPU step: NVLink SyncBN exchange + captured NCCL gradient all-reduce
mkdir -p gpurun_out
python -c "import __graft_entry__ as g; g.build()" > gpurun_out/e_build.log 2>&1
TR="python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1"
timeout 300 $TR --master-port 29511 tools/ddp_parity.py > gpurun_out/e_parity_p2p.log 2>&1; echo "rc=$?" >> gpurun_out/e_parity_p2p.log; tail -3 gpurun_out/e_parity_p2p.log
timeout 300 $TR --master-port 29514 bench.py --gpus 2 --steps 10 --warmup 3 > gpurun_out/e_bench2_graph.json 2> gpurun_out/e_bench2_graph.err; echo "rc=$?"; grep "^{" gpurun_out/e_bench2_graph.json | head -c 1000; echo; grep -v "^\*\|OMP" gpurun_out/e_bench2_graph.err | tail -5
