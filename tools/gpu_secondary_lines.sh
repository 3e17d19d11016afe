#!/bin/bash
mkdir -p gpurun_out
python -c "import __graft_entry__ as g; g.build()" > gpurun_out/m_build.log 2>&1
TR="python -m torch.distributed.run --nnodes=1 --nproc-per-node $NG --master-addr 127.0.0.1"
port=29700
for spec in $SPECS; do
  IFS=: read model size batch <<< "$spec"
  port=$((port+1))
  timeout 300 $TR --master-port $port bench.py --gpus $NG --model $model --size $size --batch $batch --steps 5 --warmup 3 > gpurun_out/m_bench_${model}_n$NG.json 2> gpurun_out/m_bench_${model}_n$NG.err; echo "$model rc=$?"
  grep "^{" gpurun_out/m_bench_${model}_n$NG.json | python -c "import sys,json; d=json.loads(sys.stdin.read()); print(d['metric'], 'N', d['n_gpus'], round(d['value'],1), round(d['ms_per_step'],2), d['config']['step_launch'], d['config']['sync_bn_exchange'], 'eager', round(d['config']['eager_ms_per_step'],2))" 2>/dev/null || (grep -v "^\*\|OMP" gpurun_out/m_bench_${model}_n$NG.err | tail -4)
done
