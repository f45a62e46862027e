#!/bin/bash
# tools/r02_n8.sh N — the BASELINE configs sharded over N GPUs of one box (J2) + the headline bench with both host modes
N=${1:-8}
TR="python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1"
export NCCL_DEBUG=INFO NCCL_DEBUG_SUBSYS=INIT
for c in 3 5 4; do
  $TR --master-port 295$c bench_configs.py --config $c --gpus $N --steps 3 --no-cpu-baseline > gpurun_out/r02x_cfg${c}_n$N.json 2> gpurun_out/r02x_cfg${c}_n$N.err
  grep -m1 -o "NCCL INFO.*nranks [0-9]*" gpurun_out/r02x_cfg${c}_n$N.err > gpurun_out/r02x_cfg${c}_n$N.nccl
done
unset NCCL_DEBUG NCCL_DEBUG_SUBSYS
$TR --master-port 29520 bench.py --gpus $N --steps 10 --warmup 3 > gpurun_out/r02x_bench_n$N.json 2> gpurun_out/r02x_bench_n$N.err
$TR --master-port 29521 bench.py --gpus $N --steps 10 --warmup 3 --e2e-mode zero_copy > gpurun_out/r02x_bench_n${N}_zc.json 2> gpurun_out/r02x_bench_n${N}_zc.err
ls -la gpurun_out/r02x*
