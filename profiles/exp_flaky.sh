#!/bin/bash
# repeat the GPU suite to catch run-to-run flakiness (atomics change the fp32 summation order between runs)
O=gpurun_out/${1:-exp_flaky}
mkdir -p $O
for i in 1 2 3; do
  timeout 600 python -m pytest tests -q -m gpu > $O/pytest_$i.log 2>&1; echo "run $i rc=$?"; grep -E "passed|failed|^FAILED|AssertionError: \(" $O/pytest_$i.log | head -6
done
python -c "import __graft_entry__ as g; g.smoke()" 2>&1 | tail -2
