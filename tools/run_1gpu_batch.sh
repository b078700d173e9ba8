# the single-GPU bench lines of every config (profiles/r02_bench_*_1gpu.json) + the ncu launch list of the bench entry point
for c in c1 c2 c2p; do python bench.py --config $c --steps 20 --warmup 3 > gpurun_out/r02_bench_${c}_1gpu.json 2> gpurun_out/r02_bench_${c}_1gpu.err; done
for c in c3 c4; do python bench.py --config $c --steps 6 --warmup 3 > gpurun_out/r02_bench_${c}_1gpu.json 2> gpurun_out/r02_bench_${c}_1gpu.err; done
python bench.py --steps 3 --warmup 3 > gpurun_out/r02_bench_c5_1gpu.json 2> gpurun_out/r02_bench_c5_1gpu.err
python bench.py --impl reference --steps 2 --warmup 1 > gpurun_out/r02_bench_c5_reference_arm.json 2> gpurun_out/r02_bench_c5_reference_arm.err
python bench.py --npoints 32768 --steps 1 --warmup 1 --no-e2e --no-cpu-baseline > gpurun_out/r02_plain_32k.log 2>&1 && ncu --metrics gpu__time_duration.sum --clock-control none --csv --log-file gpurun_out/r02_launches_32k.csv python bench.py --npoints 32768 --steps 1 --warmup 1 --no-e2e --no-cpu-baseline > gpurun_out/r02_ncu_32k.log 2>&1
tail -c 200 gpurun_out/r02_bench_c*_1gpu.err
python tools/cusolver_ref.py 8192 16384 32768 > gpurun_out/r02_cusolver_reference_point.txt 2>&1
tail -5 gpurun_out/r02_cusolver_reference_point.txt
