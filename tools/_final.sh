set -x
python -m pytest tests -m gpu -q 2>&1 | tail -3 > gpurun_out/final_tests.log
python bench.py > gpurun_out/bench_default.json 2> gpurun_out/bench_default.err
python bench.py --impl reference --steps 2 --warmup 1 > gpurun_out/bench_reference.json 2> gpurun_out/bench_reference.err
# launch list (cold-cache, serialised per-launch times) of a steady-state stretch of the default bench command
ncu --metrics gpu__time_duration.sum --clock-control none --launch-skip 40000 -c 1200 --csv --log-file gpurun_out/launches.csv \
    python bench.py --steps 1 --warmup 3 --no-cpu > gpurun_out/ncu_list.log 2>&1
# full capture of three steady-state rounds, one group, no graph (kernels serialised)
DRMLT_GROUPS=1 DRMLT_NO_GRAPH=1 ncu --set full --clock-control none --import-source on -k regex:"k_trace|k_walk|k_connect|k_chain|k_begin" \
    --launch-skip 9000 -c 15 -o gpurun_out/r01_round python bench.py --steps 1 --warmup 3 --no-cpu --chains 1048576 --mutations 32 --e2e-spp 1 > gpurun_out/ncu_full.log 2>&1
ls -la gpurun_out
tail -3 gpurun_out/final_tests.log; cat gpurun_out/bench_default.json | cut -c1-600; cat gpurun_out/bench_reference.json | cut -c1-400
