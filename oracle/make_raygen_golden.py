"""Generate ``tests/golden/raygen.npz`` by running the REFERENCE'S OWN ``Cameras.generate_rays``
(nerfstudio/cameras/cameras.py:312-482,490-726, camera_utils.py:298-401) on a few small cameras.

TEST INFRASTRUCTURE ONLY.  Runs only in the build container (needs /root/reference):

    python -m oracle.make_raygen_golden
"""
from __future__ import annotations

import os

import numpy as np
import torch

from oracle.make_golden import GOLDEN, _install_stubs


AABB = torch.tensor([[-0.4, -0.5, -0.3], [0.5, 0.35, 0.45]])  # crop box used for every camera


def cases():
    """name -> camera description (shared with tests/test_raygen.py through the npz itself)."""
    from samnerf_b200.synthetic import look_at

    c2w_a = look_at((1.1, 0.6, 0.45))[:3, :4]
    c2w_b = look_at((-0.7, 1.3, 0.2))[:3, :4]
    return {
        "perspective": dict(h=12, w=16, fx=14.0, fy=13.5, cx=8.2, cy=5.9, type=1, dist=None, c2w=c2w_a),
        "perspective_distorted": dict(h=10, w=14, fx=11.0, fy=11.5, cx=7.0, cy=5.0, type=1,
                                      dist=[0.08, -0.02, 0.004, -0.001, 0.003, -0.002], c2w=c2w_b),
        "fisheye_distorted": dict(h=9, w=11, fx=4.5, fy=4.5, cx=5.3, cy=4.6, type=2,
                                  dist=[0.03, 0.005, 0.0, 0.0, 0.0, 0.0], c2w=c2w_a),
        "equirectangular": dict(h=8, w=16, fx=8.0, fy=8.0, cx=8.0, cy=4.0, type=3, dist=None, c2w=c2w_b),
    }


def main():
    _install_stubs()
    from nerfstudio.cameras.cameras import Cameras, CameraType

    out = {}
    for name, c in cases().items():
        cam = Cameras(
            camera_to_worlds=c["c2w"][None].clone(), fx=c["fx"], fy=c["fy"], cx=c["cx"], cy=c["cy"], width=c["w"],
            height=c["h"], distortion_params=None if c["dist"] is None else torch.tensor([c["dist"]]),
            camera_type=CameraType(c["type"]),
        )
        rb = cam.generate_rays(camera_indices=0, keep_shape=True)
        assert tuple(rb.origins.shape) == (c["h"], c["w"], 3)
        out[f"{name}.origins"] = rb.origins.numpy()
        out[f"{name}.directions"] = rb.directions.numpy()
        out[f"{name}.pixel_area"] = rb.pixel_area.numpy()
        out[f"{name}.camera"] = np.array([c["fx"], c["fy"], c["cx"], c["cy"], c["w"], c["h"], c["type"]], dtype=np.float64)
        out[f"{name}.dist"] = np.array(c["dist"] if c["dist"] is not None else [], dtype=np.float32)
        out[f"{name}.c2w"] = c["c2w"].numpy()
        # an explicit coordinate list too (the LOOP B sub-grid path, sam_model.py:368-379)
        ys = torch.linspace(0, c["h"] - 1, 4, dtype=torch.long)
        xs = torch.linspace(0, c["w"] - 1, 8, dtype=torch.long)
        sub = rb[ys[:, None], xs[None, :]]
        out[f"{name}.sub_directions"] = sub.directions.numpy()
        # viewer crop box: nears / fars from the ray-box intersection (cameras.py:463-482, utils/math.py:201-238)
        from nerfstudio.data.scene_box import SceneBox

        rbx = cam.generate_rays(camera_indices=0, keep_shape=True, aabb_box=SceneBox(aabb=AABB.clone()))
        out[f"{name}.aabb_nears"] = rbx.nears.numpy()
        out[f"{name}.aabb_fars"] = rbx.fars.numpy()
        print(name, rb.directions.shape, "nan:", int(torch.isnan(rb.directions).sum()),
              "box hits:", int((rbx.nears < 1e9).sum()), "of", rbx.nears.numel())
    out["aabb"] = AABB.flatten().numpy()
    path = os.path.join(GOLDEN, "raygen.npz")
    np.savez_compressed(path, **out)
    print("wrote", path, os.path.getsize(path), "bytes")


if __name__ == "__main__":
    main()
