bash tools/bench_ab.sh r01h "-" "lib=cta9" "lib=cta10" "lib=cta12" "chunk_rays=262144" "chunk_rays=524288" "refill=24" "refill=28" "refill_occluded=2"
timeout 600 python tools/profile_trace.py --workload c3 --reps 2 --counters > gpurun_out/r01h_c3_counters.log 2>&1; tail -4 gpurun_out/r01h_c3_counters.log
timeout 900 ncu --set full --clock-control none --import-source on --profile-from-start off -k regex:k_trace -c 2 -f -o gpurun_out/r01h_c3_trace python tools/profile_trace.py --workload c3 --reps 1 > gpurun_out/r01h_ncu_c3.log 2>&1; tail -2 gpurun_out/r01h_ncu_c3.log
