"""B200-native STARK proving backend for Encrypt-zkVM (drop-in for `ExecutionProver::prove`).

The compute path is the CUDA library `libezkvm.so` (sm_100a); Python here is only the binding used by the
tests and the benchmark.  There is no CPU fallback.
"""
from .prover import (ExecutionProver, LweParameters, Proof, ProofOptions, ProverError, PublicInputs, ServerKey,
                     device_count, kernel_launch_count)
from .vm import Execution, Program, ProgramError, ProcessorError, ProgramInputs, execute, prove, synthetic_case

__all__ = ["ExecutionProver", "LweParameters", "Proof", "ProofOptions", "ProverError", "PublicInputs", "ServerKey",
           "device_count", "kernel_launch_count", "Execution", "Program", "ProgramError", "ProcessorError",
           "ProgramInputs", "execute", "prove", "synthetic_case"]
