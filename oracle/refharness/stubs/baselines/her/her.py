"""baselines.her.her.make_sample_her_transitions [upstream, recalled] -- the numpy restatement of
oracle/callers_oracle.py (call site: config.py:9,121)."""
from oracle.callers_oracle import make_sample_her_transitions  # noqa: F401
