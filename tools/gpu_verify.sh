#!/bin/bash
# Round-end verification on a GPU box: the GPU tests touched by the latest host changes first, then smoke(), then the
# rest of the suite; every step under its own timeout, logs under gpurun_out/.
mkdir -p gpurun_out
T1=${T1:-110}; T2=${T2:-180}
FIRST="tests/test_gpu_sharded.py tests/test_gpu_table.py tests/test_gpu_search.py tests/test_gpu_ingest.py"
timeout $T1 python -m pytest $FIRST -m gpu -q --durations=6 > gpurun_out/verify_first.log 2>&1; echo "rc=$?" >> gpurun_out/verify_first.log
tail -4 gpurun_out/verify_first.log
timeout 40 python -c "import __graft_entry__ as g; g.smoke()" > gpurun_out/verify_smoke.log 2>&1; echo "rc=$?" >> gpurun_out/verify_smoke.log
tail -2 gpurun_out/verify_smoke.log
IGN=""; for f in $FIRST; do IGN="$IGN --ignore=$f"; done
timeout $T2 python -m pytest tests -m gpu -q --durations=10 $IGN > gpurun_out/verify_rest.log 2>&1; echo "rc=$?" >> gpurun_out/verify_rest.log
tail -4 gpurun_out/verify_rest.log
