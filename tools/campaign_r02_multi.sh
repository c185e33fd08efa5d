# Multi-GPU part of the round-2 evidence: gpurun --gpus N -- bash tools/campaign_r02_multi.sh N
N=$1
mkdir -p gpurun_out
python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29531 bench.py --gpus $N --steps 2000 --warmup 5 > gpurun_out/f_bench_n$N.json 2> gpurun_out/f_bench_n$N.err; echo "bench rc $?"
python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29532 tools/dist_check.py > gpurun_out/f_dist_check_n$N.log 2>&1; echo "dist_check rc $?"; tail -3 gpurun_out/f_dist_check_n$N.log
