import sys, ctypes, numpy as np, tempfile
sys.path.insert(0, '/root/repo')
from oracle import front
import bench
L = ctypes.CDLL('/root/repo/experiments/simt_sim/_simt.so')
L.sim_build.restype = ctypes.c_void_p; L.sim_build.argtypes = [ctypes.c_int, ctypes.c_void_p]
L.sim_run.argtypes = [ctypes.c_void_p, ctypes.c_int, ctypes.c_void_p, ctypes.c_void_p, ctypes.c_int, ctypes.c_int, ctypes.c_void_p]
wl = sys.argv[1] if len(sys.argv) > 1 else "c2_disney_1080p_256spp"
d = tempfile.mkdtemp()
path = bench.make_scene(wl, d, spp=1)
o = front.OracleScene(path)
verts, ids = o.triangles(); verts = np.ascontiguousarray(verts, np.float32)
h = L.sim_build(verts.shape[0], verts.ctypes.data)
C_NODE, C_TRI = 45.0, 60.0

def morton_key(om, dt, bits=5):
    p = om[:, :3]; lo = p.min(0); ext = np.maximum(p.max(0) - lo, 1e-6)
    q = np.clip(((p - lo) / ext * (2 ** bits - 1)).astype(np.uint32), 0, 2 ** bits - 1)
    key = np.zeros(len(p), np.uint32)
    for b in range(bits):
        for a in range(3):
            key |= ((q[:, a] >> b) & 1) << (3 * b + a)
    octant = ((dt[:, 0] < 0).astype(np.uint32) | ((dt[:, 1] < 0).astype(np.uint32) << 1) | ((dt[:, 2] < 0).astype(np.uint32) << 2))
    return (octant << (3 * bits)) | key

def report(tag, om, dt, anyhit):
    om = np.ascontiguousarray(om, np.float32); dt = np.ascontiguousarray(dt, np.float32)
    n = om.shape[0]
    row = []
    for name, order in (("asis", None),):
        a, b = (om, dt) if order is None else (np.ascontiguousarray(om[order]), np.ascontiguousarray(dt[order]))
        for mode, mname in ((0, "ww"), (3, "ww+refill<32"), (4, "ww+refill<24"), (5, "ww+refill<16")):
            out = np.zeros(4, np.float64)
            L.sim_run(h, n, a.ctypes.data, b.ctypes.data, 1 if anyhit else 0, mode, out.ctypes.data)
            cost = (C_NODE * out[0] + C_TRI * out[1]) / n * 32 / 32
            eff = (C_NODE * out[2] + C_TRI * out[3]) / (32 * (C_NODE * out[0] + C_TRI * out[1]))
            row.append("%s/%s cost %.0f eff %.2f (lane nodes %.1f tris %.1f)" % (name, mname, cost * 32, eff, out[2] / n, out[3] / n))
    print(tag, n, "rays:\n   " + "\n   ".join(row))

npix = o.width * o.height
st = o.generate(npix // 2 - 32768, 65536)
for bnc in range(4):
    n = st["o_med"].shape[0]
    if n < 1000: break
    report("closest bounce %d" % bnc, st["o_med"], st["d_t1"], False)
    flag, det = o.closest_hit(st["o_med"], st["d_t1"])
    sh = o.shade(st, flag, det)
    m = (sh["status"] & 2) != 0
    if m.sum() > 1000:
        report("shadow  bounce %d" % bnc, sh["sh_o_med"][m], sh["sh_d_t1"][m], True)
    k = (sh["status"] & 1) != 0
    st = {kk: np.ascontiguousarray(v[k]) for kk, v in sh["next"].items()}

# ---- BVH quality: lane node steps / triangle tests of the product's builder vs lib/bvh's sweep-SAH and binned-SAH on the same rays
from oracle.front import _ptr
R = front.ref_lib()
st = o.generate(npix // 2 - 32768, 65536)
flag, det = o.closest_hit(st["o_med"], st["d_t1"])
sh = o.shade(st, flag, det)
k = (sh["status"] & 1) != 0
b1 = {kk: np.ascontiguousarray(v[k]) for kk, v in sh["next"].items()}
for builder in (0, 1):
    hb = R.ref_bvh_build(verts.shape[0], _ptr(verts, ctypes.c_float), builder)
    for tag, rays in (("primary", st), ("bounce1", b1)):
        n = rays["o_med"].shape[0]
        prim = np.zeros(n, np.int32); tuv = np.zeros((n, 3), np.float32); stats = np.zeros(2, np.uint64)
        R.ref_bvh_closest(hb, n, _ptr(rays["o_med"], ctypes.c_float), _ptr(rays["d_t1"], ctypes.c_float), _ptr(prim, ctypes.c_int32), _ptr(tuv, ctypes.c_float), _ptr(stats, ctypes.c_uint64))
        print("lib/bvh builder %d %s: steps/ray %.1f tests/ray %.1f" % (builder, tag, stats[0] / n, stats[1] / n))
    R.ref_bvh_free(hb)
