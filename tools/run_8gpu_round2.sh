set -x
T="python -m torch.distributed.run --nnodes=1 --master-addr 127.0.0.1"
for N in 2 4 8; do
  NCCL_DEBUG=INFO NCCL_DEBUG_FILE=gpurun_out/r2k_nccl_cfg5_${N}gpu.%h.%p.log $T --nproc-per-node $N --master-port $((29500+N)) bench.py --config 5 --gpus $N > gpurun_out/r2k_bench_cfg5_${N}gpu.json 2> gpurun_out/r2k_cfg5_${N}gpu.err
done
$T --nproc-per-node 8 --master-port 29611 bench.py --gpus 8 > gpurun_out/r2k_bench_8gpu.json 2> gpurun_out/r2k_bench_8gpu.err
$T --nproc-per-node 8 --master-port 29612 bench.py --config 4 --gpus 8 > gpurun_out/r2k_bench_cfg4_8gpu.json 2> gpurun_out/r2k_cfg4_8gpu.err
# keep one NCCL log per N (rank 0's), drop the rest
for N in 2 4 8; do f=$(ls gpurun_out/r2k_nccl_cfg5_${N}gpu.*.log | head -1); grep -E "NCCL version|NVLS|Channel|Connected|via P2P|nranks|comm 0x.*rank 0" $f | head -40 > gpurun_out/r2k_nccl_cfg5_${N}gpu.txt; rm -f gpurun_out/r2k_nccl_cfg5_${N}gpu.*.log; done
tail -n 2 gpurun_out/r2k_*.err
