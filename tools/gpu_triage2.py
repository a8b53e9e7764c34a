import os, sys, warnings
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
for p in (ROOT, os.path.join(ROOT, "oracle"), os.path.join(ROOT, "tests"), os.path.join(ROOT, "tools")):
    sys.path.insert(0, p)
warnings.simplefilter("ignore")
import numpy as np
import lwfbrook90.jl_b200 as lw
from lwfbrook90.jl_b200 import ensemble as en
from conftest import inputs
from gpu_triage import compare
s = lw.setup(lw.loadSPAC(inputs("DAV2020-full"), "DAV2020-full", simulate_isotopes=True))
ens = en.make_ensemble(s, 64)
compare("dav " + os.environ.get("LWFB_LIB", "default"), ens, int(sys.argv[1]) if len(sys.argv) > 1 else 3, True, 12)
