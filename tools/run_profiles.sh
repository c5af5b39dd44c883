# round-2 profile captures (one GPU): launch list of the bench command, full-set capture of one steady-state k_enc launch
mkdir -p gpurun_out
ncu --metrics gpu__time_duration.sum --clock-control none -s 800 -c 320 --csv --log-file gpurun_out/r02_launches_raw.csv \
    python bench.py --steps 2 --warmup 3 --regions 1 > gpurun_out/r02_launches_bench.log 2>&1
ncu --set full --clock-control none --import-source on -k regex:k_enc -s 23 -c 1 -o gpurun_out/r02_kenc -f \
    python tools/stage_cycles.py --in-flight 1 --chunk-pairs 8 --steps 2 > gpurun_out/r02_kenc_ncu.log 2>&1
ncu --set full --clock-control none -k regex:k_conv -s 2 -c 1 -o gpurun_out/r02_kconv -f \
    python tools/stage_cycles.py --in-flight 1 --chunk-pairs 8 --steps 1 > gpurun_out/r02_kconv_ncu.log 2>&1
ncu --set full --clock-control none -k regex:k_decoder -s 2 -c 1 -o gpurun_out/r02_kdec -f \
    python tools/stage_cycles.py --in-flight 1 --chunk-pairs 8 --steps 1 > gpurun_out/r02_kdec_ncu.log 2>&1
ls -la gpurun_out | tail -12
