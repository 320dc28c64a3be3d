"""CPU oracle for the InteractVLM inference hot path -- TEST INFRASTRUCTURE, NOT PRODUCT CODE.

Plain numpy restatements of the reference's algorithms, each function citing the reference file:line it
follows.  Only tests/, __graft_entry__.smoke() and bench.py's cpu_baseline / --impl reference leg may
import this package, and only as the checker or the CPU baseline; interactvlm_b200/ never imports it.

Pinning: the reference (saidwivedi/InteractVLM) ships no tests or golden vectors, so the
oracle is pinned against outputs of the UNMODIFIED reference run in the build container under the import
shims of oracle/ref_shim.py; the generating script is oracle/make_goldens.py and the vectors live in
tests/golden/.  tests/test_oracle_golden.py checks every oracle function against them on CPU.
"""
