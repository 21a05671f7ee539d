"""Generate tests/golden/cameras.npz from the UNMODIFIED reference camera classes (unidepth/utils/camera.py,
imported from /root/reference through oracle/ref_shims): for every camera model, the sequence `infer` applies to a
camera argument (unidepthv2.py:267-303,361-362: BatchCamera.from_camera -> crop(-pads) -> resize(factor) ->
get_rays) plus unproject / project / reconstruct on seeded inputs.

Run here (CPU container, has /root/reference):   python oracle/make_golden_cameras.py
TEST INFRASTRUCTURE ONLY.
"""
import os
import sys
import warnings

import numpy as np
import torch

HERE = os.path.dirname(os.path.abspath(__file__))
REF = "/root/reference"
sys.path[:0] = [REF, os.path.join(HERE, "ref_shims"), HERE]

H, W = 30, 44                      # ray-map size used for every case
PADS = (3, 3, 5, 5)                # left, right, top, bottom as infer passes them (negated)
FACTOR = 0.73

# name -> (class name, params).  One camera per object: the reference's own unproject reshapes to the batch of the pixel
# grid (1), so a Pinhole holding several K's cannot produce rays (camera.py:255-267); batches are BatchCameras built with
# torch.cat (the "mixed" case below).
#  Parameter layouts follow the reference classes (camera.py:229,276,331,412,697,977).
CASES = {
    "pinhole_params": ("Pinhole", [[40.0, 38.0, 21.5, 15.25]]),
    "eucm": ("EUCM", [[30.0, 31.0, 22.0, 15.0, 0.6, 1.1]]),
    "spherical": ("Spherical", [[1.0, 1.0, 22.0, 15.0, 44.0, 30.0, 1.2, 0.7]]),
    "opencv_radial": ("OPENCV", [[40.0, 38.0, 21.5, 15.25, -0.12, 0.03, -0.004, 0, 0, 0, 0, 0, 0, 0, 0, 0]]),
    "opencv_full": ("OPENCV", [[40.0, 38.0, 21.5, 15.25, -0.10, 0.02, 0.0, 0, 0, 0, 0.004, -0.003, 0.002, 0.001, -0.002, 0.001]]),
    "fisheye624": ("Fisheye624", [[24.0, 24.5, 22.0, 15.0, 0.05, -0.01, 0.004, -0.001, 0.0, 0.0, 0.002, -0.001, 0.001, 0.0005, -0.001, 0.0005]]),
    "fisheye624_radial": ("Fisheye624", [[24.0, 24.5, 22.0, 15.0, 0.05, -0.01, 0.004, -0.001, 0.0, 0.0, 0, 0, 0, 0, 0, 0]]),
    "mei": ("MEI", [[45.0, 44.0, 22.0, 15.0, -0.08, 0.01, 0.002, -0.001, 0.9]]),
    "mei_plain": ("MEI", [[45.0, 44.0, 22.0, 15.0, 0.0, 0.0, 0.0, 0.0, 0.9]]),
}


def main():
    warnings.simplefilter("ignore")
    import unidepth.utils.camera as C
    from unidepth.utils.coordinate import coords_grid
    out = {}
    g = torch.Generator().manual_seed(7)
    pts = torch.randn(2, 3, 6, 8, generator=g)
    pts[:, 2] = pts[:, 2].abs() + 1.0                     # in front of the camera
    depth = torch.rand(2, 1, H, W, generator=g) * 5 + 0.5
    out["points"] = pts.numpy()
    out["depth"] = depth.numpy()
    for name, (cls, params) in CASES.items():
        p = torch.tensor(params, dtype=torch.float32)
        b = p.shape[0]
        make = lambda: getattr(C, cls)(params=p.clone())
        cam = make()
        out[f"{name}/params"] = p.numpy()
        out[f"{name}/K"] = cam.K.numpy()
        uv = coords_grid(b, H, W)
        out[f"{name}/unproject"] = cam.unproject(uv).numpy()
        out[f"{name}/get_rays"] = make().get_rays((b, H, W)).numpy()
        try:
            cam2 = make()
            out[f"{name}/project"] = cam2.project(pts[:b].clone()).numpy()
            if cam2.projection_mask is not None:
                out[f"{name}/projection_mask"] = cam2.projection_mask.numpy()
        except Exception as e:                             # MEI.project / Spherical need nothing special; record failures
            print(name, "project failed:", type(e).__name__, e)
        out[f"{name}/reconstruct"] = make().reconstruct(depth[:b].clone()).numpy()
        # what infer does with a camera object
        bc = C.BatchCamera.from_camera(make())
        bc = bc.crop(left=-PADS[0], top=-PADS[2], right=-PADS[1], bottom=-PADS[3])
        bc = bc.resize(FACTOR)
        out[f"{name}/infer_params"] = bc.params.numpy()
        out[f"{name}/infer_K"] = bc.K.numpy()
        out[f"{name}/infer_rays"] = bc.get_rays(shapes=(b, H, W)).numpy()
        print(name, "ok", {k.split('/')[1]: v.shape for k, v in out.items() if k.startswith(name + "/")})
    # Pinhole built from K (the `camera=K` tensor branch wraps K this way, unidepthv2.py:273-278)
    K = torch.tensor([[[40.0, 0, 21.5], [0, 38.0, 15.25], [0, 0, 1]]])
    cam = C.Pinhole(K=K.clone())
    out["pinhole_K/K_in"] = K.numpy()
    out["pinhole_K/params"] = cam.params.numpy()
    out["pinhole_K/get_rays"] = cam.get_rays((1, H, W)).numpy()
    out["pinhole_K/pinhole_rays"] = cam.get_pinhole_rays((1, H, W)).numpy()
    out["pinhole_K/hfov"] = cam.hfov.numpy()
    out["pinhole_K/vfov"] = cam.vfov.numpy()
    fl = C.Pinhole(K=K.clone()).flip(H, W, "horizontal")
    out["pinhole_K/flip_params"] = fl.params.numpy()
    # a batch of two different camera models (torch.cat over Camera objects, camera.py:182-207)
    mixed = torch.cat([C.BatchCamera.from_camera(C.Pinhole(params=torch.tensor(CASES["pinhole_params"][1]))),
                       C.BatchCamera.from_camera(C.EUCM(params=torch.tensor(CASES["eucm"][1])))])
    out["mixed/params"] = mixed.params.numpy()
    out["mixed/K"] = mixed.K.numpy()
    out["mixed/get_rays"] = mixed.get_rays((2, H, W)).numpy()
    out["mixed/classes"] = np.array(mixed.original_class)
    path = os.path.join(HERE, "..", "tests", "golden", "cameras.npz")
    np.savez_compressed(path, **out)
    print("wrote", path, os.path.getsize(path), "bytes")


if __name__ == "__main__":
    main()
