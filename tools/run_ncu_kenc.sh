# one full-set ncu capture of a steady-state k_enc launch (query + source phase), whole batch on one stream
mkdir -p gpurun_out
ncu --set full --clock-control none --import-source on -k regex:k_enc -s 13 -c 1 -o gpurun_out/${1:-kenc} -f \
    python tools/stage_cycles.py --in-flight 1 --chunk-pairs 0 --steps 2 > gpurun_out/${1:-kenc}_ncu.log 2>&1
tail -3 gpurun_out/${1:-kenc}_ncu.log
ls -la gpurun_out/
