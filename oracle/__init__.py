"""CPU oracle for the attention-forward hot path.  TEST INFRASTRUCTURE ONLY.

Only tests/, __graft_entry__.smoke() and bench.py's cpu_baseline / --impl reference legs may
import this package; the product (quantizedmha_b200) never does.
"""
from .oracle import Oracle, RefLib, build_oracle, load_oracle, load_ref  # noqa: F401
