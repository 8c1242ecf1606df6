"""CPU oracle for the batched Clifford-algebra path -- TEST INFRASTRUCTURE ONLY.

Only tests/, __graft_entry__.smoke() and bench.py's CPU-baseline legs may import this
package; the product (largecrimsoncanine_b200/, lcc/, the `largecrimsoncanine` extension)
never does.  See oracle/lcc_oracle.c for what is restated and how it is pinned.
"""
from .oracle import *  # noqa: F401,F403
