set -u
OUT=gpurun_out; mkdir -p $OUT
N=${1:-4}
timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29543 bench.py --gpus $N --steps 10 --warmup 3 > $OUT/bench_n${N}_v2.json 2> $OUT/bench_n${N}_v2.err
echo rc=$?; wc -l $OUT/bench_n${N}_v2.json; cut -c1-300 $OUT/bench_n${N}_v2.json
