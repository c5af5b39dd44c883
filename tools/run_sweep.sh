for inf in 2 3 4; do for cp in -1 4 6 11 16; do
  OETR_TIMING= python tools/stage_cycles.py --in-flight $inf --chunk-pairs $cp --steps 30 2>&1 | head -1 | sed "s/^/in-flight $inf chunk $cp: /"
done; done
