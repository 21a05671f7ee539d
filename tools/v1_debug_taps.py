"""Stage-by-stage comparison of the V1 engine with the oracle: runs udb_infer_v1 eagerly with UDB_V1_DUMP set (the engine
then writes named intermediates, csrc/engine_v1.cu `tap`) and prints each tap's error against the oracle's taps.
    python tools/v1_debug_taps.py            (GPU box; debugging aid, not a test)"""
import json
import os
import sys
import tempfile

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path[:0] = [ROOT, os.path.join(ROOT, "oracle"), os.path.join(ROOT, "tests")]
dump = tempfile.mkdtemp()
os.environ["UDB_V1_DUMP"] = dump

import numpy as np  # noqa: E402
import torch  # noqa: E402

import unidepth_v1_oracle as O1  # noqa: E402
from test_oracle_golden import v1_case_inputs  # noqa: E402
from unidepth_b200 import UniDepthV1  # noqa: E402

cfg, sd, rgb, K, meta, z = v1_case_inputs(os.path.join(ROOT, "tests", "golden"), "v1_cnvnxtl_480x640")
taps = {}
torch.set_num_threads(32)
ref = O1.infer_v1(sd, cfg, rgb, None, taps=taps)
m = UniDepthV1(cfg)
m.load_state_dict(sd, strict=True)
m = m.to("cuda:0").eval()
m.use_cuda_graph = False
out = m.infer(rgb)
torch.cuda.synchronize()


def load(name, dtype, shape):
    a = np.fromfile(os.path.join(dump, name + ".bin"), dtype=dtype)
    return torch.from_numpy(a.astype(np.float32)).reshape(shape)


def report(name, got, want):
    got, want = got.float(), want.float()
    floor = 0.1 * want.abs().mean()
    e = (got - want).abs() / want.abs().clamp(min=floor)
    print(f"{name:14s} max {e.max().item():.3e} mean {e.mean().item():.3e}   (|ref| mean {want.abs().mean().item():.3e})")


B = 1
report("enc_last", load("enc_last", np.float32, (B, 14, 19, 1536)), taps["enc_last"])
feat = taps["features"]                                     # [B, nq, hid, 4]
report("tokens", load("tokens", np.float16, (B, 4, 1064, 512)), feat.permute(0, 3, 1, 2))
kn = taps["K_net"]
report("intr4", load("intr4", np.float32, (B, 4)), torch.stack([kn[:, 0, 0], kn[:, 1, 1], kn[:, 0, 2], kn[:, 1, 2]], 1))
report("lat16", load("lat16", np.float32, (B, 1064, 512)), taps["latents_16"])
for name, s in (("out8", 2), ("out4", 4), ("out2", 8)):
    report(name, load(name, np.float32, (B, 1, 28 * s, 38 * s)), taps[name])
for k in ("intrinsics", "depth", "points"):
    report("final " + k, out[k].cpu(), ref[k])
