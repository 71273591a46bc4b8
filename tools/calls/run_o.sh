#!/bin/bash
# call o: e2e A/B of the hit-download modes
bash tools/bench_ab.sh r01o "-" "d2h=1" "d2h=2" "d2h=1,chunk_rays=2097152" "d2h=2,chunk_rays=524288"
