#!/bin/bash
mkdir -p gpurun_out
python -c "import __graft_entry__ as g; g.build()" > gpurun_out/e_build.log 2>&1
TR="python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1"
timeout 120 $TR --master-port 29521 tools/diag_graph_ddp.py > gpurun_out/e_diag_ddp.log 2>&1; echo "rc=$?" >> gpurun_out/e_diag_ddp.log; grep "rank 0\|rc=\|Error" gpurun_out/e_diag_ddp.log | tail -6
timeout 240 $TR --master-port 29514 bench.py --gpus 2 --steps 20 --warmup 5 > gpurun_out/e_bench2_graph.json 2> gpurun_out/e_bench2_graph.err; echo "rc=$?"; grep "^{" gpurun_out/e_bench2_graph.json | head -c 1300; echo; grep -v "^\*\|OMP" gpurun_out/e_bench2_graph.err | tail -4
timeout 120 python -m pytest tests/test_gpu_ddp.py -q > gpurun_out/e_ddp_test.log 2>&1; tail -2 gpurun_out/e_ddp_test.log
