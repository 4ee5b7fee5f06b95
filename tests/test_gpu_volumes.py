"""GPU parity for volume primitives (the scene file's "volumes" array; core/volume.cpp, core/volume/*.cpp): the medium stage,
shadow-ray transmittance and whole renders against the oracle on identical batches.  Tolerances as in
test_gpu_materials_media.py: integers exact; tracking decisions compare a uniform draw with a filtered density, so a few
walks per thousand may diverge and are counted."""
import numpy as np
import pytest

from tests import util
from tests.test_gpu_materials_media import _bounce_states, _compare_next, _film_want, _frac_close
from tests.util import f32

pytestmark = pytest.mark.gpu


@pytest.fixture(scope="module")
def pair_volumes(tmp_path_factory, built):
    from btrc_b200 import api, scenes
    from oracle import front

    d = tmp_path_factory.mktemp("vol_small")
    path = scenes.make_volume_scene(str(d), width=64, height=64, spp=8, grid_res=16, state_count=50000, n_extra=6)
    s = api.Scene(path)
    o = front.OracleScene(path)
    yield s, o, path
    s.close()


def test_volume_medium_stage_parity(pair_volumes):
    s, o, _ = pair_volumes
    assert s.info.has_medium == 1 and o.has_medium()
    n_scatter = 0
    for depth, (st, flag, detail, mo, st2, flag2, sh) in enumerate(_bounce_states(o, 64 * 64 * 4, 5, with_medium=True)):
        s.film_clear()
        g_st, g_flag, nxt, shd = s.medium(st, flag, detail)
        assert np.array_equal(g_st["o_med"][:, 3].view(np.uint32), st2["o_med"][:, 3].view(np.uint32))
        same = (g_flag == flag2)
        assert same.mean() > 0.995, (depth, same.mean())
        unsc = same & ((flag2 >> 31) == 0)
        agree = (g_st["rng"][unsc] == st2["rng"][unsc]).all(axis=1)
        assert agree.mean() > 0.995
        assert _frac_close(g_st["ble"][unsc], st2["ble"][unsc], 2e-3, 1e-6) > 0.99
        keep = (mo["status"] & 1) != 0
        ref = {k: v[keep] for k, v in mo["next"].items()}
        _compare_next(nxt, ref, rtol=5e-3, max_flip=0.02, max_outlier=0.02)
        has_sh = (mo["status"] & 2) != 0
        util.compare_shadow_rays(shd, st, has_sh, mo["sh_o_med"], mo["sh_d_t1"], mo["sh_beta_li"], rtol=5e-3, atol_dir=2e-3,
                                 max_flip=0.02, max_outlier=0.02)
        n_scatter += int(keep.sum())
    assert n_scatter > 1000


def test_volume_shadow_transmittance_parity(pair_volumes):
    s, o, _ = pair_volumes
    rows = _bounce_states(o, 64 * 64 * 4, 3, with_medium=True)
    got_any = False
    for st, flag, detail, mo, st2, flag2, sh in rows:
        for out, src in ((mo, st), (sh, st2)):  # shadow rays of the medium stage and of the surface shade
            has_sh = (out["status"] & 2) != 0
            if has_sh.sum() < 100:
                continue
            shd = dict(pixel=np.ascontiguousarray(src["pixel"][has_sh]), beta_li=np.ascontiguousarray(out["sh_beta_li"][has_sh]),
                       o_med=np.ascontiguousarray(out["sh_o_med"][has_sh]), d_t1=np.ascontiguousarray(out["sh_d_t1"][has_sh]))
            s.film_clear()
            s.shadow(shd)
            rad = np.zeros((64, 64, 4), f32)
            wgt = np.zeros((64, 64), f32)
            s.film_export(rad, wgt)
            occ = o.any_hit(shd["o_med"], shd["d_t1"]).astype(bool)
            tr = o.shadow_tr(shd["o_med"], shd["d_t1"])
            assert (tr[~occ] < 0.999).mean() > 0.05  # the volumes do attenuate a share of the rays
            want = _film_want((64, 64), shd["pixel"][~occ], shd["beta_li"][~occ, :3] * tr[~occ])
            ok = np.isclose(rad[..., :3], want, rtol=2e-2, atol=1e-4).all(axis=-1)
            assert ok.mean() > 0.98, ok.mean()
            got_any = True
    assert got_any


def test_render_matches_oracle_volumes(tmp_path, built):
    from btrc_b200 import api, scenes
    from oracle import front

    path = scenes.make_volume_scene(str(tmp_path), width=64, height=64, spp=16, grid_res=16, state_count=40000, n_extra=2)
    s = api.Scene(path)
    o = front.OracleScene(path)
    color, _, _, stats = s.render()
    ref = o.render()
    assert stats["n_gen"] == ref["stats"]["n_gen"]
    for k in ("n_closest", "n_hit", "n_next", "n_shadow", "n_shadow_miss"):
        assert abs(stats[k] - ref["stats"][k]) <= 1e-2 * ref["stats"][k], (k, stats[k], ref["stats"][k])
    a, b = color[..., :3], ref["color"][..., :3]
    assert np.isfinite(a).all()
    assert abs(a.mean() - b.mean()) / b.mean() < 2e-2
    assert util.rel_mse(a, b) < 0.15
    # the volumes are visible: the same scene without them is brighter behind the cloud
    s.close()
