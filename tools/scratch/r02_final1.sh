python bench.py > gpurun_out/r02_bench_final.out 2> gpurun_out/r02_bench_final.err; grep '^{' gpurun_out/r02_bench_final.out > gpurun_out/r02_bench_final.json
python bench.py --impl reference --steps 3 --warmup 1 > gpurun_out/r02_bench_final_refarm.json 2>> gpurun_out/r02_bench_final.err
python bench.py --impl reference --mode reference --steps 2 --warmup 1 > gpurun_out/r02_bench_final_refmode_refarm.json 2>> gpurun_out/r02_bench_final.err
tail -c 300 gpurun_out/r02_bench_final.err
