import sys; sys.path.insert(0, ".")
import numpy as np, torch
from giwaxsim_b200 import engine
from oracle import giwaxs_oracle as ox
g = np.load("tests/golden/graphite262.npz")
codes, uniq = engine.encode_values(g["f_values"]); b = g["bounds"]
eng = engine.SliceEngine(g["coords"], float(g["r"]), g["q_axis"], int(g["grid_size"]), complex(g["avg_voxel_f"]), b[0], b[1], True, 5, species=codes, table=uniq)
probe = [int(i) for i in g["probe"]]
t = eng.prepare(g["phis"][probe])
N = eng.N; q_num = eng.q_num
cols = t["col"].cpu().numpy().reshape(-1, N)
for k, i in enumerate(probe):
    cm = g["colmask_%d" % i]
    got = cols[k] >= 0
    bad = np.where(got != cm)[0]
    print("slice", i, "phi", g["phis"][i], "mask mismatches", len(bad), bad[:10], "kept ref", cm.sum(), "kept gpu", got.sum())
    hx, hy, vz = ox.slice_q_axes(g["phis"][i], N, float(g["r"]))
    sn, cs, xl, xr, yl, yr = eng._phi_scalars(g["phis"][[i]])
    print("  host scalars xl,xr,yl,yr", xl, xr, yl, yr, " oracle ends", hx[0], hx[-1], hy[0], hy[-1])
    print("  qmin,qmax,dq", eng.qmin, eng.qmax, eng.dq)
    both = got & cm
    ref_packed = np.full(N, -1); ref_packed[cm] = g["iy_%d" % i] * q_num + g["ix_%d" % i]
    print("  index mismatches among kept", np.count_nonzero(cols[k][both] != ref_packed[both]))
    if len(bad):
        j = bad[0]; print("  first bad col", j, "hx", hx[j], "hy", hy[j], "gpu", cols[k][j])
