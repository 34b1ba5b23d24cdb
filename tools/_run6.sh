cd $GRAFT_REPO_ROOT
mkdir -p gpurun_out
export EZK_TRACE_CACHE=/tmp/ezk_traces
timeout 900 python -m pytest tests -m gpu -x -q > gpurun_out/r01t_pytest.log 2>&1; echo "pytest exit $?" >> gpurun_out/r01t_pytest.log
tail -2 gpurun_out/r01t_pytest.log
for i in 1 2; do python tools/profile_prove.py 20 2>&1 | tail -3; done > gpurun_out/r01t_new.log 2>&1
cat gpurun_out/r01t_new.log
