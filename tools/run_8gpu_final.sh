T="python -m torch.distributed.run --nnodes=1 --nproc-per-node 8 --master-addr 127.0.0.1 --master-port 29511"
$T bench.py --gpus 8 --steps 3 --warmup 2 > gpurun_out/r02_final_c5_8gpu.json 2> gpurun_out/r02_final_c5_8gpu.err
GPX_MG_PANEL_SUB_ROWS=0 $T bench.py --gpus 8 --steps 3 --warmup 2 --no-e2e --parity-n 0 > gpurun_out/r02_final_c5_8gpu_nosub.json 2> gpurun_out/r02_final_c5_8gpu_nosub.err
$T bench.py --config c4 --gpus 8 --steps 5 --warmup 3 > gpurun_out/r02_final_c4_8gpu.json 2> gpurun_out/r02_final_c4_8gpu.err
tail -c 200 gpurun_out/r02_final_*.err
