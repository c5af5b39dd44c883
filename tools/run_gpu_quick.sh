mkdir -p gpurun_out
timeout 900 python -m pytest tests/test_hotpath_gpu.py -m gpu -x -q -k "stage_parity or edge or full_size" 2>&1 | tail -4
OETR_TIMING=1 timeout 300 python tools/stage_cycles.py > gpurun_out/tq_stage.txt 2>&1; cat gpurun_out/tq_stage.txt
