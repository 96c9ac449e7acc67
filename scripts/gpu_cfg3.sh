# BASELINE.json configs[2] (1,000 views, 499,500 pairs, 2,000 correspondences/pair, 40 % outliers; 32 GB of FP64 rows) on N GPUs
cd $GRAFT_REPO_ROOT
mkdir -p gpurun_out
free -g | head -2; nproc
AVAIL=$(free -g | awk '/Mem:/ {print $7}')
if [ "$AVAIL" -lt 48 ]; then echo "not enough host memory ($AVAIL GB)"; exit 0; fi
make -C oracle -s 2>&1 | tail -3
N=${N:-1}
if [ "$N" = "1" ]; then
  timeout 1700 python bench.py --config cfg3_1000v --steps 1 --warmup 1 --cpu-sample 1024 > gpurun_out/bench_cfg3_n$N.json 2> gpurun_out/bench_cfg3_n$N.err
else
  timeout 1700 python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29517 bench.py --gpus $N --config cfg3_1000v --steps 1 --warmup 1 --cpu-sample 1024 > gpurun_out/bench_cfg3_n$N.json 2> gpurun_out/bench_cfg3_n$N.err
fi
echo rc=$?; tail -5 gpurun_out/bench_cfg3_n$N.err; cat gpurun_out/bench_cfg3_n$N.json | cut -c1-3000
