# round-2 neck captures (one GPU): full GPU test suite, launch list of the neck bench, full-set capture of k_neck_conv / proj / out
mkdir -p gpurun_out
timeout 1500 python -m pytest tests -m gpu -x -q 2>&1 | tail -6 > gpurun_out/n1_pytest.txt; cat gpurun_out/n1_pytest.txt
timeout 200 python tools/neck_bench.py --no-torch > gpurun_out/n1_bench.json 2> gpurun_out/n1_bench.err; cat gpurun_out/n1_bench.json
timeout 200 python tools/neck_bench.py --no-torch --size 52 --pairs 16 > gpurun_out/n1_bench_840.json 2>> gpurun_out/n1_bench.err; cat gpurun_out/n1_bench_840.json
timeout 300 ncu --metrics gpu__time_duration.sum --clock-control none -c 40 --csv --log-file gpurun_out/r02_neck_launches_raw.csv \
    python tools/neck_bench.py --no-torch --steps 4 > gpurun_out/n1_ncu_launches.log 2>&1
timeout 400 ncu --set full --clock-control none --import-source on -k regex:k_neck -s 9 -c 3 -o gpurun_out/r02_neck -f \
    python tools/neck_bench.py --no-torch --steps 2 > gpurun_out/n1_ncu_full.log 2>&1
ls -la gpurun_out | tail -8
