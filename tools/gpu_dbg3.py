import os, sys, warnings
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT); sys.path.insert(0, os.path.join(ROOT, "oracle")); sys.path.insert(0, os.path.join(ROOT, "tests"))
warnings.simplefilter("ignore")
import numpy as np
import lwfbrook90.jl_b200 as lw
from lwfbrook90.jl_b200 import capi, layout as lo
from conftest import inputs
s = lw.setup(lw.loadSPAC(inputs("DAV2020-full"), "DAV2020-full", simulate_isotopes=True))
h = capi.Handle(s.N, True, 0, 2); h.force_legacy(True); h.set_sites(s.forcing); h.set_members([0], s.scal, s.lay, s.u0); h.run()
print(h.get_daily()[0][:, :6])
