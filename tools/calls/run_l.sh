bash tools/bench_ab.sh r01l "-" "lib=head" "gpu_builder=ploc,ploc_radius=8" "gpu_builder=ploc,ploc_radius=16"
for w in c2 c3; do
for srt in none octant chunkdir:4096:8 chunkdir:16384:16 chunkdir:65536:16 global:4:8 global:5:8 global:6:16; do
  echo "== $w sort=$srt" | tee -a gpurun_out/r01l_sort.log
  timeout 900 python tools/profile_trace.py --workload $w --reps 2 --bands 4 --sort $srt --cfg "gpu_builder=ploc" 2>&1 | tail -1 | tee -a gpurun_out/r01l_sort.log
done; done
