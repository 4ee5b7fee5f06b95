"""Diagnostic (GPU box): per-material error quantiles of the shade stage against the oracle on the small C2 scene."""
import json, sys, tempfile
import numpy as np
sys.path.insert(0, ".")
from btrc_b200 import api, scenes
from oracle import front
from tests.test_gpu_materials_media import _bounce_states, W2, H2

d = tempfile.mkdtemp()
path = scenes.make_material_test(d, width=W2, height=H2, spp=8, rings=12, segments=16, tex_res=64, state_count=60000)
root = json.load(open(path))
mats = [e["material"]["type"] + ("/tex" if isinstance(e["material"].get("color"), dict) else "") for e in root["scene"]["entities"]]
s = api.Scene(path); o = front.OracleScene(path)
for depth, (st, flag, detail, _, _, _, sh) in enumerate(_bounce_states(o, W2 * H2 * 4, 3)):
    nxt, shd = s.shade(st, flag, detail)
    keep = (sh["status"] & 1) != 0
    ref = {k: v[keep] for k, v in sh["next"].items()}
    inst = (flag & 0x3fffffff)[keep]
    ka = {int(k): i for i, k in enumerate(nxt["rng"][:, 0])}
    kb = {int(k): i for i, k in enumerate(ref["rng"][:, 0])}
    common = [k for k in ka if k in kb]
    ia = np.array([ka[k] for k in common]); ib = np.array([kb[k] for k in common])
    print(f"depth {depth}: gpu {len(ka)} oracle {len(kb)} common {len(common)}")
    a, b = nxt["beta_depth"][ia, :3], ref["beta_depth"][ib, :3]
    rel = (np.abs(a - b) / (np.abs(b) + 1e-5)).max(axis=1)
    dd = np.abs(nxt["d_t1"][ia, :3] - ref["d_t1"][ib, :3]).max(axis=1)
    ii = inst[ib]
    for m in sorted(set(ii)):
        sel = ii == m
        q = np.quantile(rel[sel], [0.5, 0.9, 0.99, 1.0])
        qd = np.quantile(dd[sel], [0.5, 0.99, 1.0])
        print(f"  inst {m:2d} {mats[m]:12s} n={sel.sum():5d} beta rel q50/90/99/max = {q[0]:.2e} {q[1]:.2e} {q[2]:.2e} {q[3]:.2e}   dir abs q50/99/max = {qd[0]:.1e} {qd[1]:.1e} {qd[2]:.1e}")
    worst = np.argsort(-rel)[:5]
    for w in worst:
        print("   worst", mats[ii[w]], a[w], b[w], nxt["d_t1"][ia[w], :3], ref["d_t1"][ib[w], :3], "pdf", nxt["ble"][ia[w], 3], ref["ble"][ib[w], 3])
