#!/bin/bash
# Ring slot size and launch-group cap of the host-trace pipeline (environment knobs, one box).
mkdir -p gpurun_out
run() { python - "$@" <<'P'
import json, os, sys, time
sys.path.insert(0, os.getcwd())
import encrypt_zkvm_b200 as ezk
label = sys.argv[1]
prog, ex = ezk.synthetic_case(2, 20)
trace, ph, outs = ex.trace(), prog.hash(), ex.outputs()
with ezk.ExecutionProver(ezk.ProofOptions(), ph, outs, ezk.ServerKey()) as p:
    for _ in range(3): p.prove(trace)
    w = []
    for _ in range(9):
        t0 = time.perf_counter(); p.prove(trace); w.append((time.perf_counter() - t0) * 1e3)
    w.sort()
    print(json.dumps({"setting": label, "median_ms": round(w[4], 3), "min_ms": round(w[0], 3), "upload_stage_ms": round(p.stage_times_ms()["upload"], 3)}), flush=True)
P
}
{
run default
EZK_STAGE_SLOT_KB=8192 run slot_8MiB
EZK_STAGE_SLOT_KB=4096 run slot_4MiB
EZK_STAGE_SLOT_KB=2048 run slot_2MiB
EZK_HOST_GROUP_CAP=14 run cap_14
EZK_HOST_GROUP_CAP=4 run cap_4
EZK_HOST_GROUP_CAP=2 run cap_2
run default_again
} > gpurun_out/ro_slots.log 2>&1
cat gpurun_out/ro_slots.log
