"""Diagnostic (GPU box): shadow rays of the Cornell shade stage that exist on one side only."""
import sys, tempfile
import numpy as np
sys.path.insert(0, ".")
from btrc_b200 import api, scenes
from oracle import front
from tests.test_gpu_parity import _oracle_bounce_states

d = tempfile.mkdtemp()
path = scenes.make_cornell(d, width=96, height=96, spp=8, state_count=20000)
s = api.Scene(path); o = front.OracleScene(path)
st, flag, detail, sh = _oracle_bounce_states(o, 96 * 96 * 2, 1)[0]
nxt, shd = s.shade(st, flag, detail)
has_sh = (sh["status"] & 2) != 0
print("gpu", shd["pixel"].shape[0], "oracle", has_sh.sum())
# hit point of every input state
t = detail[:, 0].view(np.float32)
hitp = st["o_med"][:, :3] + t[:, None] * st["d_t1"][:, :3]
# match each GPU shadow ray to an input state: same pixel, nearest hit point
from collections import defaultdict
bypix = defaultdict(list)
for i, (px, py) in enumerate(st["pixel"]):
    bypix[(int(px), int(py))].append(i)
owner = np.zeros(shd["pixel"].shape[0], np.int64)
for j, (px, py) in enumerate(shd["pixel"]):
    c = bypix[(int(px), int(py))]
    dist = [np.abs(hitp[i] - shd["o_med"][j, :3]).max() for i in c]
    owner[j] = c[int(np.argmin(dist))]
cnt = np.bincount(owner, minlength=st["pixel"].shape[0])
print("states with >1 gpu shadow ray:", (cnt > 1).sum(), " gpu ray but no oracle ray:", ((cnt > 0) & ~has_sh).sum(), " oracle ray but no gpu ray:", ((cnt == 0) & has_sh).sum())
bad = np.where((cnt > 0) & ~has_sh)[0][:8]
for i in bad:
    js = np.where(owner == i)[0]
    print("state", i, "pixel", st["pixel"][i], "flag", hex(flag[i]), "prim", detail[i, 1], "t", t[i], "hitp", hitp[i], "oracle status", sh["status"][i], "splat", sh["splat"][i])
    print("    ray d", st["d_t1"][i, :3])
    for j in js:
        print("    gpu shadow o", shd["o_med"][j, :3], "d", shd["d_t1"][j], "beta", shd["beta_li"][j, :3])
dup = np.where(cnt > 1)[0][:5]
for i in dup:
    js = np.where(owner == i)[0]
    print("dup state", i, "pixel", st["pixel"][i], [shd["o_med"][j, :3] for j in js], "same-pixel states", bypix[tuple(int(x) for x in st["pixel"][i])])
