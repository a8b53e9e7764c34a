import os, sys, warnings
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
for p in (ROOT, os.path.join(ROOT, "oracle"), os.path.join(ROOT, "tests"), os.path.join(ROOT, "tools")):
    sys.path.insert(0, p)
warnings.simplefilter("ignore")
import numpy as np
import lwfbrook90.jl_b200 as lw
from lwfbrook90.jl_b200 import capi, ensemble as en, layout as lo
import oracle as orc
from conftest import inputs
s = lw.setup(lw.loadSPAC(inputs("DAV2020-full"), "DAV2020-full", simulate_isotopes=True))
ens = en.make_ensemble(s, 48)
nd = int(sys.argv[1]) if len(sys.argv) > 1 else 40
orc.use("elem")
ref, st_o, cnt_o = orc.run(ens.N, ens.iso, 0, nd, ens.forcing, ens.site_of_member, ens.scal, ens.lay, ens.u0, n_threads=8)
h = capi.Handle(ens.N, True, 0, nd); h.set_sites(ens.forcing); h.set_members(ens.site_of_member, ens.scal, ens.lay, ens.u0); h.run()
g = h.get_daily(); st, cnt = h.get_status()
same = (g == ref) | (np.isnan(g) & np.isnan(ref))
bad = np.nonzero(~same.all(axis=(1, 2)))[0]
print("lib", os.environ.get("LWFB_LIB", "default"), "bit-exact members", ens.M - len(bad), "/", ens.M)
ix = lo.StateIndex(ens.N, True)
names = {}
for k, v in vars(ix).items():
    if isinstance(v, int): names[v] = k
    elif isinstance(v, slice): 
        for j in range(v.start, v.stop): names[j] = f"{k}[{j - v.start}]"
for i in bad[:4]:
    d0 = np.nonzero((~same[i]).any(axis=1))[0][0]; cols = np.nonzero(~same[i, d0])[0]
    print(f" member {i} first mismatch at record {d0}: ", [(names.get(c, c), g[i, d0, c], ref[i, d0, c]) for c in cols[:6]])
