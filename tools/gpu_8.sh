#!/bin/bash
mkdir -p gpurun_out
python -c "import __graft_entry__ as g; g.build()" > gpurun_out/n8_build.log 2>&1
TR="python -m torch.distributed.run --nnodes=1 --nproc-per-node 8 --master-addr 127.0.0.1"
timeout 240 $TR --master-port 29614 bench.py --gpus 8 --steps 10 --warmup 3 > gpurun_out/n8_bench.json 2> gpurun_out/n8_bench.err; echo "rc=$?"; grep "^{" gpurun_out/n8_bench.json | head -c 1300; echo; grep -v "^\*\|OMP" gpurun_out/n8_bench.err | tail -6
