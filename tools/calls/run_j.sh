timeout 900 python -m pytest tests -m gpu -x -q 2>&1 | tail -4
for cfg in "gpu_builder=lbvh" "gpu_builder=ploc,ploc_radius=4" "gpu_builder=ploc,ploc_radius=8" "gpu_builder=ploc,ploc_radius=16" "gpu_builder=ploc,ploc_radius=32"; do
  for w in c2 c3; do
    echo "== $w $cfg" | tee -a gpurun_out/r01j_ploc.log
    timeout 600 python tools/profile_trace.py --workload $w --reps 2 --counters --cfg "$cfg" 2>&1 | grep -v "^$" | tail -5 | tee -a gpurun_out/r01j_ploc.log
  done
done
