"""B200-native charge-flux Ewald electrostatics behind the openmm-chargeflux plugin boundary.

Host-side mirror of the plugin's API (``CoulForce``) plus the ``CalcCoulForce`` kernel that runs the
hand-written sm_100a CUDA path through the C ABI in include/cfx_b200.h.
"""
from .force import CoulForce  # noqa: F401
