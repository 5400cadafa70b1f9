"""CPU oracle for the SURF path -- TEST INFRASTRUCTURE ONLY.

Only tests/, __graft_entry__.smoke() and bench.py's cpu_baseline / --impl reference legs may import
this package.  The product (surffeatures_b200) never does.  Parity is UNPINNED by the reference
(stropitek/SurfFeatures holds no golden vectors); see oracle/surf_oracle.h.
"""
from .surf_oracle import *  # noqa: F401,F403
