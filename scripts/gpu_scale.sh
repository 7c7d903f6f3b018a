#!/usr/bin/env bash
# bench.py at N GPUs (the driver's launch line); prints the parsed line
set -u
N=${1:-8}
mkdir -p gpurun_out/scale
python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29551 bench.py --gpus $N --steps 20 --warmup 5 > gpurun_out/scale/bench_n$N.json 2> gpurun_out/scale/bench_n$N.err
echo "exit $?"; tail -3 gpurun_out/scale/bench_n$N.err
tail -1 gpurun_out/scale/bench_n$N.json | python -c "
import json,sys; d=json.loads(sys.stdin.read()); print('N=%d' % d['n_gpus'], 'pairs/s %.4g' % d['value'], 'ms/step %.3f' % d['ms_per_step'], d['roofline']['phase_ms_per_step'], 'e2e ms %.3f' % d['e2e']['ms_per_step'], 'finite', d['config']['finite'])"
