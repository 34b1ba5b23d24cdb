#!/bin/bash
# Full GPU test suite + smoke on the final code of round 2.
mkdir -p gpurun_out
(time timeout 400 python -m pytest tests -m gpu -x -q) > gpurun_out/rp_pytest.log 2>&1; echo "pytest rc=$?" >> gpurun_out/rp_pytest.log
timeout 120 python -c "import __graft_entry__ as g; g.smoke()" > gpurun_out/rp_smoke.log 2>&1; echo "smoke rc=$?" >> gpurun_out/rp_smoke.log
tail -6 gpurun_out/rp_pytest.log; tail -4 gpurun_out/rp_smoke.log
