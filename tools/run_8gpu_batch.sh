set -x
T="python -m torch.distributed.run --nnodes=1 --nproc-per-node 8 --master-addr 127.0.0.1 --master-port 29511"
$T bench.py --gpus 8 --steps 3 --warmup 3 --dump-launches gpurun_out/r02_l8c > gpurun_out/r02_bench_c5_8gpu.json 2> gpurun_out/r02_bench_c5_8gpu.err
GPX_MG_GROUP_K=2048 $T bench.py --gpus 8 --steps 2 --warmup 2 --no-e2e --parity-n 0 > gpurun_out/r02_bench_c5_8gpu_k2048.json 2> gpurun_out/r02_bench_c5_8gpu_k2048.err
$T bench.py --config c3 --gpus 8 --steps 5 --warmup 3 > gpurun_out/r02_bench_c3_8gpu.json 2> gpurun_out/r02_bench_c3_8gpu.err
$T bench.py --config c4 --gpus 8 --steps 5 --warmup 3 > gpurun_out/r02_bench_c4_8gpu.json 2> gpurun_out/r02_bench_c4_8gpu.err
$T bench.py --config c2p --gpus 8 --steps 10 --warmup 3 > gpurun_out/r02_bench_c2p_8gpu.json 2> gpurun_out/r02_bench_c2p_8gpu.err
tail -c 300 gpurun_out/r02_bench_c*_8gpu*.err
