#!/usr/bin/env bash
# NG-GPU evidence: bench lines of both workloads and a sweep subset under torch.distributed.run (parity check optional)
set -u
NG=$1
OUT=gpurun_out/multi$NG; mkdir -p $OUT
SKIP_CHECK=${SKIP_CHECK:-1} bash scripts/gpu_multi.sh $NG c3 c4
timeout 900 python -m torch.distributed.run --nnodes=1 --nproc-per-node $NG --master-addr 127.0.0.1 --master-port 29700 scripts/sweep.py --ns 65536,262144,1048576 --ds 8,64,256 --no-f64 --budget 30 > $OUT/sweep.jsonl 2> $OUT/sweep.err
echo "sweep exit $?"; grep -c ms_per_step $OUT/sweep.jsonl; grep -v "OMP_NUM\|\*\*\*" $OUT/sweep.err | tail -3
