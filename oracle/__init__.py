"""CPU oracle of the neural_raytracing hot path -- TEST INFRASTRUCTURE ONLY.

Only tests/, __graft_entry__.smoke() and bench.py's cpu_baseline / `--impl reference` legs may
import this package.  The product package (neural_raytracing_b200) never does.

  oracle.c_oracle  ctypes front-end of oracle/c/nrt_oracle.c (fixed-order fp32 restatement)
  oracle.port      torch-CPU restatement of the reference's eager op sequence (shading glue,
                   losses) -- also what the CPU baseline times
  oracle.ref_shim  imports the unmodified reference (build container only)

Parity status: PINNED.  The reference ships no tests or golden vectors, so
the oracle is pinned against outputs of the unmodified reference run in the build container:
tests/golden/*.npz, produced by tests/golden/make_golden.py, checked by
tests/test_oracle_vs_golden.py.
"""
