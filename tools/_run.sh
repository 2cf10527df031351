python -m pytest tests -m gpu -x -q > gpurun_out/t35.log 2>&1; tail -3 gpurun_out/t35.log
python tools/bench_kernels.py 1250 score1 > gpurun_out/k35_plain.log 2>&1 && cat gpurun_out/k35_plain.log && \
ncu --set full --clock-control none --import-source on -k regex:score_streamN -s 3 -c 1 -f -o gpurun_out/prof_r1_score6 python tools/bench_kernels.py 1250 score1 > gpurun_out/k35_ncu.log 2>&1; tail -2 gpurun_out/k35_ncu.log
python bench.py --steps 3 --warmup 3 --no-cpu-baseline > gpurun_out/b35.log 2> gpurun_out/b35.err; tail -c 1500 gpurun_out/b35.log
