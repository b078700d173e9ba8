# the single-GPU bench lines of every config (profiles/r02_bench_*_1gpu.json)
for c in c1 c2 c2p; do python bench.py --config $c --steps 20 --warmup 3 > gpurun_out/r02_bench_${c}_1gpu.json 2> gpurun_out/r02_bench_${c}_1gpu.err; done
for c in c3 c4; do python bench.py --config $c --steps 6 --warmup 3 > gpurun_out/r02_bench_${c}_1gpu.json 2> gpurun_out/r02_bench_${c}_1gpu.err; done
python bench.py --steps 3 --warmup 3 > gpurun_out/r02_bench_c5_1gpu.json 2> gpurun_out/r02_bench_c5_1gpu.err
tail -c 200 gpurun_out/r02_bench_c*_1gpu.err
python tools/cusolver_ref.py 8192 16384 > gpurun_out/r02_cusolver_reference_point.txt 2>&1
