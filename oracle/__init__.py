"""CPU oracle for the B200 optical-flow hot path.  TEST INFRASTRUCTURE ONLY.

Importable from tests/, __graft_entry__.smoke() and bench.py's cpu_baseline /
--impl reference legs.  The product package never imports it.
"""
from .oracle import *  # noqa: F401,F403
