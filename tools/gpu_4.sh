#!/bin/bash
mkdir -p gpurun_out
python -c "import __graft_entry__ as g; g.build()" > gpurun_out/n4_build.log 2>&1
TR="python -m torch.distributed.run --nnodes=1 --nproc-per-node 4 --master-addr 127.0.0.1"
timeout 240 $TR --master-port 29714 bench.py --gpus 4 --steps 20 --warmup 5 > gpurun_out/n4_bench.json 2> gpurun_out/n4_bench.err; echo "rc=$?"; grep "^{" gpurun_out/n4_bench.json | head -c 500; echo; grep -v "^\*\|OMP" gpurun_out/n4_bench.err | tail -3
