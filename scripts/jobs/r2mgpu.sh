mkdir -p gpurun_out
nvidia-smi --query-gpu=index,name --format=csv,noheader | head -8
N=${NGPU:-2}
timeout 900 python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29517 bench.py --gpus $N --steps 8 --warmup 3 > gpurun_out/r2_bench_N800_${N}gpu.json 2> gpurun_out/r2_bench_N800_${N}gpu.err; echo "bench rc=$?"; cut -c1-1500 gpurun_out/r2_bench_N800_${N}gpu.json; tail -5 gpurun_out/r2_bench_N800_${N}gpu.err | cut -c1-300
