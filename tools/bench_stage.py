"""Device-resident stage timings (LDE + Merkle, FRI) through the C ABI's bench entry points.

    python tools/bench_stage.py [log_n ...]        # default 16 18 20
"""
import sys
from pathlib import Path

sys.path.insert(0, str(Path(__file__).resolve().parent.parent))
import encrypt_zkvm_b200 as ezk

logs = [int(x) for x in sys.argv[1:]] or [16, 18, 20]
with ezk.ExecutionProver(ezk.ProofOptions(), [0, 0], [0] * 16, ezk.ServerKey()) as p:
    for log_n in logs:
        n = 1 << log_n
        for width in (28, 7):
            lde, mk = p.bench_lde_merkle(width, n, 3)
            algo = 9 * n * width * 16
            print(f"log_n={log_n} width={width} lde_ms={lde:.3f} ({algo / lde / 1e6:.0f} GB/s algorithmic) merkle_ms={mk:.3f}")
        print(f"log_n={log_n} fri_ms={p.bench_fri(n, 3):.3f}")
