#!/bin/bash
# round-2 GPU call E (2 GPUs): NVLink SyncBN exchange — parity, then NCCL vs P2P step time
mkdir -p gpurun_out
python -c "import __graft_entry__ as g; g.build()" > gpurun_out/e_build.log 2>&1
TR="python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1"
timeout 300 $TR --master-port 29511 tools/ddp_parity.py > gpurun_out/e_parity_p2p.log 2>&1; echo "rc=$?" >> gpurun_out/e_parity_p2p.log; tail -4 gpurun_out/e_parity_p2p.log
TSB_SYNCBN_P2P=0 timeout 300 $TR --master-port 29512 tools/ddp_parity.py > gpurun_out/e_parity_nccl.log 2>&1; echo "rc=$?" >> gpurun_out/e_parity_nccl.log; tail -3 gpurun_out/e_parity_nccl.log
TSB_SYNCBN_P2P=0 timeout 300 $TR --master-port 29513 bench.py --gpus 2 --steps 10 --warmup 3 > gpurun_out/e_bench2_nccl.json 2> gpurun_out/e_bench2_nccl.err; echo "rc=$?"; head -c 400 gpurun_out/e_bench2_nccl.json; echo
timeout 300 $TR --master-port 29514 bench.py --gpus 2 --steps 10 --warmup 3 > gpurun_out/e_bench2_p2p.json 2> gpurun_out/e_bench2_p2p.err; echo "rc=$?"; head -c 400 gpurun_out/e_bench2_p2p.json; echo; tail -5 gpurun_out/e_bench2_p2p.err
