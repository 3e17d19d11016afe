#!/bin/bash
mkdir -p gpurun_out
python -c "import __graft_entry__ as g; g.build()" > gpurun_out/e_build.log 2>&1
TR="python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1"
timeout 100 $TR --master-port 29521 tools/diag_graph_ddp.py > gpurun_out/e_diag_ddp.log 2>&1; echo "rc=$?" >> gpurun_out/e_diag_ddp.log; grep "rank\|rc=\|Error\|File" gpurun_out/e_diag_ddp.log | tail -30
B=16 SIZE=1024 timeout 150 $TR --master-port 29522 tools/diag_graph_ddp.py > gpurun_out/e_diag_ddp_big.log 2>&1; echo "rc=$?" >> gpurun_out/e_diag_ddp_big.log; grep "rank\|rc=\|Error\|File" gpurun_out/e_diag_ddp_big.log | tail -30
