# usage: bash tools/run_ngpu_batch.sh N   -- the bench lines of one GPU count (profiles/r02_bench_*_{N}gpu.json)
N=$1
T="python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29511"
$T bench.py --gpus $N --steps 3 --warmup 3 > gpurun_out/r02_bench_c5_${N}gpu.json 2> gpurun_out/r02_bench_c5_${N}gpu.err
$T bench.py --config c3 --gpus $N --steps 5 --warmup 3 > gpurun_out/r02_bench_c3_${N}gpu.json 2> gpurun_out/r02_bench_c3_${N}gpu.err
$T bench.py --config c4 --gpus $N --steps 5 --warmup 3 > gpurun_out/r02_bench_c4_${N}gpu.json 2> gpurun_out/r02_bench_c4_${N}gpu.err
$T bench.py --config c2p --gpus $N --steps 10 --warmup 3 > gpurun_out/r02_bench_c2p_${N}gpu.json 2> gpurun_out/r02_bench_c2p_${N}gpu.err
tail -c 200 gpurun_out/r02_bench_c*_${N}gpu.err
