"""CPU oracle for the MOCHA per-frame characterization hot path — TEST INFRASTRUCTURE ONLY.

A NumPy restatement of the reference's algorithm (DK-Jang/MOCHA_SIGASIA2023), each function citing
the reference file:line it follows. Only tests/, __graft_entry__.smoke() and bench.py's
cpu_baseline / --impl reference legs may import it; the product path (mocha_sigasia2023_b200)
never does and fails loudly without its CUDA library.

Parity pin: the reference ships no tests or golden vectors, so this oracle is pinned
against outputs of the live reference code run in the build container: oracle/gen_golden.py imports
/root/reference, feeds it seeded inputs and the deterministic weights of
mocha_sigasia2023_b200/weights.py, and commits the outputs under tests/golden/;
tests/test_oracle_golden.py checks every oracle function against them.
"""
from . import nets, rot, inertial, matching, driver  # noqa: F401
