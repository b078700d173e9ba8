T="python -m torch.distributed.run --nnodes=1 --nproc-per-node 8 --master-addr 127.0.0.1 --master-port 29511"
for v in 0 1; do
GPX_MG_PANEL_SUB=$v $T bench.py --gpus 8 --steps 3 --warmup 2 --no-e2e --parity-n 2048 > gpurun_out/r02_ab_c5_8gpu_sub$v.json 2> gpurun_out/r02_ab_c5_8gpu_sub$v.err
GPX_MG_PANEL_SUB=$v $T bench.py --config c3 --gpus 8 --steps 5 --warmup 3 > gpurun_out/r02_ab_c3_8gpu_sub$v.json 2> gpurun_out/r02_ab_c3_8gpu_sub$v.err
done
tail -c 200 gpurun_out/r02_ab_*.err
