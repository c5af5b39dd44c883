mkdir -p gpurun_out
python - > gpurun_out/t1_selftest.txt 2>&1 <<'PY'
import ctypes, sys
sys.path.insert(0, '.')
from oetr_b200 import cabi
lib = cabi.load_library()
e = (ctypes.c_float * 8)()
rc = lib.oetr_selftest_tcgen05(e, 8)
print("selftest rc", rc, list(e), lib.oetr_last_error())
PY
cat gpurun_out/t1_selftest.txt
timeout 600 python -m pytest tests -m gpu -x -q 2>&1 | tail -15 > gpurun_out/t1_pytest.txt; cat gpurun_out/t1_pytest.txt
OETR_TIMING=1 timeout 300 python tools/stage_cycles.py > gpurun_out/t1_stage.txt 2>&1; cat gpurun_out/t1_stage.txt
timeout 600 python bench.py --steps 20 --warmup 3 > gpurun_out/t1_bench.json 2> gpurun_out/t1_bench.err; cut -c1-300 gpurun_out/t1_bench.json
