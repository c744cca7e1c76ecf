"""CPU checkers for the charge-flux Ewald path. TEST INFRASTRUCTURE, NOT PRODUCT.

Only tests/, __graft_entry__.smoke() and bench.py's cpu_baseline / --impl reference legs import this.
"""
from .binding import Oracle, ReferenceBuild, oracle_available, reference_available  # noqa: F401
