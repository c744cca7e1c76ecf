import sys, os
sys.path.insert(0,'.')
import numpy as np
from openmm_chargeflux_b200 import synthetic
from oracle import ReferenceBuild
PLUGIN = os.path.join('openmm_chargeflux_b200', 'plugin', 'libOpenMMCoulB200.so')
pos, box, force = synthetic.config("c1")
stage = sys.argv[1]
if stage == 'ref':
    r = ReferenceBuild(force, box, platform="Reference")
elif stage == 'b200':
    r = ReferenceBuild(force, box, platform="B200", plugin=os.path.abspath(PLUGIN))
    print(r.execute(pos, box)[0])
elif stage == 'cfx':
    from openmm_chargeflux_b200 import runtime
    c = runtime.CoulContext(force, box); print(c.evaluate(pos)[0])
print("importing torch"); sys.stdout.flush()
import torch
print("ok", torch.__version__)
