"""B200-native STARK proving backend for Encrypt-zkVM (drop-in for `ExecutionProver::prove`).

The compute path is the CUDA library `libezkvm.so` (sm_100a); Python here is only the binding used by the
tests and the benchmark.  There is no CPU fallback: the first use of any name below loads the library and
fails loudly if it is missing (`python -m encrypt_zkvm_b200.build` builds it).
"""
import importlib

_EXPORTS = {
    "ExecutionProver": "prover", "LweParameters": "prover", "Proof": "prover", "ProofOptions": "prover",
    "ProverError": "prover", "VerifierError": "prover", "PublicInputs": "prover", "ServerKey": "prover", "device_count": "prover", "verify": "prover", "wire_compat": "prover", "LocalGroup": "prover", "prove_sharded_in_process": "prover",
    "kernel_launch_count": "prover", "profile_enable": "prover", "profile_reset": "prover", "profile_read": "prover",
    "Execution": "vm", "Program": "vm", "ProgramError": "vm", "ProcessorError": "vm", "ProgramInputs": "vm",
    "execute": "vm", "prove": "vm", "synthetic_case": "vm",
}
__all__ = sorted(_EXPORTS)


def __getattr__(name):  # PEP 562: keeps `python -m encrypt_zkvm_b200.build` importable before the .so exists
    if name in _EXPORTS:
        return getattr(importlib.import_module(f".{_EXPORTS[name]}", __name__), name)
    raise AttributeError(f"module {__name__!r} has no attribute {name!r}")
