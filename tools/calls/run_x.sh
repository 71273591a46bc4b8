#!/bin/bash
OUT=gpurun_out; mkdir -p $OUT
timeout 1500 python -m pytest tests -m gpu -q -x > $OUT/r02v_pytest_gpu.log 2>&1; echo "pytest exit $?" >> $OUT/r02v_pytest_gpu.log; tail -6 $OUT/r02v_pytest_gpu.log
python -c "import __graft_entry__ as g; g.smoke(); print('smoke ok')" > $OUT/r02v_smoke.log 2>&1; tail -2 $OUT/r02v_smoke.log
