"""Camera ray generation: oracle vs the reference's own ``Cameras.generate_rays`` output, the
kernel body (host emulation) vs the same golden vectors, and - on a GPU - the kernel and ``snrf_render_camera``."""
import ctypes as C
import os

import numpy as np
import pytest
import torch

from oracle import raygen_oracle as RO

GOLDEN = os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden", "raygen.npz")
CASES = ["perspective", "perspective_distorted", "fisheye_distorted", "equirectangular"]
# stated tolerance: unit directions agree to 2e-6 absolute (fp32 divide / sqrt / sin / cos rounding and, on the device,
# fma contraction inside the Newton iteration); pixel areas to 1e-4 relative
DIR_ATOL, AREA_RTOL = 2e-6, 1e-4


def _case(name):
    z = np.load(GOLDEN)
    fx, fy, cx, cy, w, h, typ = z[f"{name}.camera"].tolist()
    dist = z[f"{name}.dist"].tolist()
    return dict(fx=fx, fy=fy, cx=cx, cy=cy, w=int(w), h=int(h), type=int(typ), dist=dist or None,
                c2w=torch.from_numpy(z[f"{name}.c2w"]), origins=z[f"{name}.origins"], directions=z[f"{name}.directions"],
                pixel_area=z[f"{name}.pixel_area"], sub_directions=z[f"{name}.sub_directions"])


@pytest.mark.parametrize("name", CASES)
def test_oracle_matches_reference(name):
    c = _case(name)
    ys, xs = RO.full_image_pixels(c["h"], c["w"])
    o, d, a = RO.generate_rays(c["fx"], c["fy"], c["cx"], c["cy"], c["c2w"], ys, xs, c["type"], c["dist"])
    assert np.array_equal(o.numpy(), c["origins"])
    np.testing.assert_allclose(d.numpy(), c["directions"], rtol=0, atol=1e-7)
    np.testing.assert_allclose(a.numpy(), c["pixel_area"], rtol=1e-5, atol=0)


def _emu_rays(c, rows, cols, patch, want_area=True, aabb=None):
    from emu.build_emu import load

    lib = load()
    n_rows = len(rows) if rows is not None else c["h"]
    n_cols = len(cols) if cols is not None else c["w"]
    n = n_rows * n_cols
    o, d, a = np.empty((n, 3), np.float32), np.empty((n, 3), np.float32), np.empty((n,), np.float32)
    intr = np.array([c["fx"], c["fy"], c["cx"], c["cy"]], np.float32)
    dist = np.array(c["dist"] or [0] * 6, np.float32)
    c2w = np.ascontiguousarray(c["c2w"].numpy(), np.float32)
    r = None if rows is None else np.ascontiguousarray(rows, np.int32)
    cc = None if cols is None else np.ascontiguousarray(cols, np.int32)
    p = lambda x: None if x is None else x.ctypes.data_as(C.c_void_p)
    box = None if aabb is None else np.ascontiguousarray(aabb, np.float32)
    nr, fr = np.empty((n,), np.float32), np.empty((n,), np.float32)
    lib.emu_generate_rays(p(intr), c["type"], int(c["dist"] is not None), p(dist), p(c2w), p(r), n_rows, p(cc), n_cols, patch,
                          p(o), p(d), p(a) if want_area else None, p(box), p(nr), p(fr))
    return (o, d, a) if aabb is None else (o, d, a, nr, fr)


@pytest.mark.parametrize("name", CASES)
def test_kernel_body_matches_reference(name):
    """The __host__ __device__ body of raygen_kernel, run on the CPU, against the reference's output."""
    c = _case(name)
    o, d, a = _emu_rays(c, None, None, 1)
    assert np.array_equal(o.reshape(c["h"], c["w"], 3), c["origins"])
    np.testing.assert_allclose(d.reshape(c["h"], c["w"], 3), c["directions"], rtol=0, atol=DIR_ATOL)
    np.testing.assert_allclose(a.reshape(c["h"], c["w"], 1), c["pixel_area"], rtol=AREA_RTOL, atol=0)
    # explicit row / column lists (the LOOP B sub-grid)
    ys = torch.linspace(0, c["h"] - 1, 4, dtype=torch.long).numpy()
    xs = torch.linspace(0, c["w"] - 1, 8, dtype=torch.long).numpy()
    _, d2, _ = _emu_rays(c, ys, xs, 1, want_area=False)
    np.testing.assert_allclose(d2.reshape(4, 8, 3), c["sub_directions"], rtol=0, atol=DIR_ATOL)


@pytest.mark.parametrize("name", CASES)
def test_crop_box_near_far(name):
    """Viewer crop box (generate_rays(aabb_box=...), cameras.py:463-482): oracle and kernel body against the
    reference's nears / fars; misses report 1e10 twice."""
    z = np.load(GOLDEN)
    c = _case(name)
    box = z["aabb"]
    ys, xs = RO.full_image_pixels(c["h"], c["w"])
    o, d, _ = RO.generate_rays(c["fx"], c["fy"], c["cx"], c["cy"], c["c2w"], ys, xs, c["type"], c["dist"])
    tn, tf = RO.intersect_aabb(o.reshape(-1, 3), d.reshape(-1, 3), torch.from_numpy(box))
    want_n, want_f = z[f"{name}.aabb_nears"].reshape(-1), z[f"{name}.aabb_fars"].reshape(-1)
    np.testing.assert_allclose(tn.numpy(), want_n, rtol=1e-5, atol=1e-6)
    np.testing.assert_allclose(tf.numpy(), want_f, rtol=1e-5, atol=1e-6)
    _, _, _, nr, fr = _emu_rays(c, None, None, 1, want_area=False, aabb=box)
    hit = want_n < 1e9
    assert hit.any() and (~hit).any()
    assert np.array_equal(nr < 1e9, hit)                      # same hit / miss decision for every ray
    np.testing.assert_allclose(nr[hit], want_n[hit], rtol=2e-5, atol=2e-6)
    np.testing.assert_allclose(fr[hit], want_f[hit], rtol=2e-5, atol=2e-6)
    assert (nr[~hit] == 1e10).all() and (fr[~hit] == 1e10).all()


def test_kernel_body_patch_major_order():
    """Ray order of the feature grid: (fh, p, fw, p) -> (fh, fw, p, p) flattened (sam_model.py:376-379)."""
    c = _case("perspective")
    fh, fw, p = 2, 3, 2
    hind, wind = RO.feature_grid_pixels(c["h"], c["w"], fh, fw, p)
    _, want, _ = RO.generate_rays(c["fx"], c["fy"], c["cx"], c["cy"], c["c2w"], hind, wind, c["type"], c["dist"])
    rows = torch.linspace(0, c["h"] - 1, fh * p, dtype=torch.long).numpy()
    cols = torch.linspace(0, c["w"] - 1, fw * p, dtype=torch.long).numpy()
    _, got, _ = _emu_rays(c, rows, cols, p, want_area=False)
    np.testing.assert_allclose(got, want.numpy(), rtol=0, atol=DIR_ATOL)


def test_oracle_reproduces_the_synthetic_orbit_camera():
    """bench.py's 800x800 frame comes from synthetic.orbit_rays; the camera path must give the same rays."""
    from samnerf_b200.synthetic import look_at, orbit_rays

    o_ref, d_ref = orbit_rays(40, 40, 40.0)
    c2w = look_at((1.2, 0.0, 0.4))
    ys, xs = RO.full_image_pixels(40, 40)
    o, d, _ = RO.generate_rays(40.0, 40.0, 20.0, 20.0, c2w[:3, :4], ys, xs)
    assert torch.equal(o.reshape(-1, 3), o_ref.reshape(-1, 3))
    np.testing.assert_allclose(d.numpy().reshape(-1, 3), d_ref.numpy().reshape(-1, 3), rtol=0, atol=1e-6)


# ---------------------------------------------------------------------------------------------------------
@pytest.mark.gpu
@pytest.mark.parametrize("name", CASES)
def test_gpu_kernel_matches_oracle(name):
    from samnerf_b200 import SAMNeRFConfig
    from samnerf_b200.renderer import Camera, Renderer

    c = _case(name)
    r = Renderer(SAMNeRFConfig.tiny(), device=0)
    cam = Camera(c["fx"], c["fy"], c["cx"], c["cy"], c["w"], c["h"], c["c2w"], c["type"], c["dist"])
    o, d, a = r.generate_rays(cam, pixel_area=True)
    torch.cuda.synchronize()
    assert np.array_equal(o.cpu().numpy().reshape(c["h"], c["w"], 3), c["origins"])
    np.testing.assert_allclose(d.cpu().numpy().reshape(c["h"], c["w"], 3), c["directions"], rtol=0, atol=DIR_ATOL)
    np.testing.assert_allclose(a.cpu().numpy().reshape(c["h"], c["w"], 1), c["pixel_area"], rtol=AREA_RTOL, atol=0)
    # a big image against the oracle, and the patch-major sub-grid
    big = Camera(800.0, 790.0, 400.0, 401.5, 800, 800, c["c2w"], c["type"], c["dist"])
    o, d, a = r.generate_rays(big, pixel_area=True)
    ys, xs = RO.full_image_pixels(800, 800)
    _, dw, aw = RO.generate_rays(800.0, 790.0, 400.0, 401.5, c["c2w"], ys, xs, c["type"], c["dist"])
    np.testing.assert_allclose(d.cpu().numpy(), dw.numpy().reshape(-1, 3), rtol=0, atol=DIR_ATOL)
    hind, wind = RO.feature_grid_pixels(800, 800, 64, 64, 4)
    rows = torch.linspace(0, 799, 256, dtype=torch.long)
    _, d2, _ = r.generate_rays(big, rows=rows, cols=rows, patch=4)
    _, dw2, _ = RO.generate_rays(800.0, 790.0, 400.0, 401.5, c["c2w"], hind, wind, c["type"], c["dist"])
    np.testing.assert_allclose(d2.cpu().numpy(), dw2.numpy(), rtol=0, atol=DIR_ATOL)


@pytest.mark.gpu
def test_gpu_render_camera_equals_render_frame():
    from helpers import make_renderer, model_pair
    from samnerf_b200.renderer import Camera
    from samnerf_b200.synthetic import look_at

    cfg, params, _ = model_pair("tiny", "scene", 5, False, 4)
    r = make_renderer(cfg, params)
    cam = Camera(60.0, 60.0, 32.0, 24.0, 64, 48, look_at((1.1, 0.6, 0.45))[:3, :4])
    o, d, _ = r.generate_rays(cam)
    ref = r.render_frame(o, d, get_feature=("sam",), chunk=1024)
    out = r.render_camera(cam, get_feature=("sam",), chunk=1024)
    torch.cuda.synchronize()
    for k in ("rgb", "depth", "accumulation", "prop_depth_0", "sam"):
        assert torch.equal(out[k], ref[k]), k
    # patch-aggregated feature map on the strided sub-grid == the per-chunk shim path
    rows = torch.linspace(0, 47, 12 * 4, dtype=torch.long)
    cols = torch.linspace(0, 63, 16 * 4, dtype=torch.long)
    o2, d2, _ = r.generate_rays(cam, rows=rows, cols=cols, patch=4)
    want = r.render(o2, d2, get_feature=("sam",), patch=True)["sam"]
    got = r.render_camera(cam, rows=rows, cols=cols, get_feature=("sam",), patch=True, chunk=1024)["sam"]
    torch.cuda.synchronize()
    assert torch.equal(got, want)
    with pytest.raises(RuntimeError, match="outside the image"):
        r.generate_rays(cam, rows=torch.tensor([0, 48]), cols=cols)


@pytest.mark.gpu
def test_gpu_model_from_camera_equals_model_from_ray_bundle():
    """SAMModel.get_outputs_for_camera == get_outputs_for_camera_ray_bundle on the rays of the same camera."""
    from helpers import model_pair
    from samnerf_b200.nerfstudio_api import RayBundle, SAMModel
    from samnerf_b200.renderer import Camera
    from samnerf_b200.synthetic import look_at

    cfg, params, _ = model_pair("tiny", "scene", 6, True, 4)
    m = SAMModel(cfg)
    m.load_state_dict(params)
    cam = Camera(32.0, 32.0, 16.0, 12.0, 32, 24, look_at((1.1, 0.6, 0.45))[:3, :4])
    o, d, _ = m.renderer.generate_rays(cam)
    bundle = RayBundle(origins=o.view(24, 32, 3), directions=d.view(24, 32, 3), pixel_area=torch.ones(24, 32, 1, device=o.device),
                       camera_indices=torch.zeros(24, 32, 1, dtype=torch.long, device=o.device))
    want = m.get_outputs_for_camera_ray_bundle(bundle)
    got = m.get_outputs_for_camera(cam)
    torch.cuda.synchronize()
    assert set(got) == set(want)
    for k in want:
        assert torch.equal(got[k], want[k]), k
    # clicks: both entry points lift / remember / re-project the same prompts, and the camera path neither drops what the
    # ray-bundle path remembered nor needs the caller to hand it intrinsics (they are the camera's)
    intrin = torch.tensor([[32.0, 0.0, 16.0], [0.0, 32.0, 12.0], [0.0, 0.0, 1.0]])
    clicks = [[10, 8], [20, 15]]
    a = m.get_outputs_for_camera_ray_bundle(bundle, points=clicks, intrin=intrin, c2w=cam.camera_to_world)
    remembered = m.prompts.clone()
    b = m.get_outputs_for_camera(cam)              # interleaved frame without new clicks: prompts stay
    assert torch.equal(m.prompts, remembered) and torch.equal(b["prompt_points"], a["prompt_points"])
    m.prompts = None
    c = m.get_outputs_for_camera(cam, points=clicks)
    assert torch.allclose(m.prompts, remembered) and torch.equal(c["prompt_points"], a["prompt_points"])
    m.get_outputs_for_camera(cam, points=[])       # an empty click list clears them, as in the reference
    assert m.prompts is None


@pytest.mark.gpu
def test_gpu_crop_box_render():
    """snrf_render_camera with a crop box == snrf_render with the nears / fars snrf_generate_rays reports."""
    from helpers import make_renderer, model_pair
    from samnerf_b200.renderer import Camera
    from samnerf_b200.synthetic import look_at

    z = np.load(GOLDEN)
    cfg, params, _ = model_pair("tiny", "scene", 5, False, 1)
    r = make_renderer(cfg, params)
    c = _case("perspective")
    cam = Camera(c["fx"], c["fy"], c["cx"], c["cy"], c["w"], c["h"], c["c2w"], aabb=z["aabb"])
    o, d, _, nr, fr = r.generate_rays(cam)
    np.testing.assert_allclose(nr.cpu().numpy().reshape(-1), z["perspective.aabb_nears"].reshape(-1), rtol=2e-5, atol=2e-6)
    np.testing.assert_allclose(fr.cpu().numpy().reshape(-1), z["perspective.aabb_fars"].reshape(-1), rtol=2e-5, atol=2e-6)
    want = r.render(o, d, nears=nr, fars=fr, get_feature=("sam",))
    got = r.render_camera(cam, get_feature=("sam",))
    torch.cuda.synchronize()
    for k in ("rgb", "depth", "accumulation", "sam"):
        assert torch.equal(torch.nan_to_num(got[k]), torch.nan_to_num(want[k])), k
