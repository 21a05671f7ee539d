"""One eager `infer` step between cudaProfilerStart/Stop, for ncu:
  ncu --profile-from-start off --metrics gpu__time_duration.sum --clock-control none --csv \
      --log-file gpurun_out/launches.csv python tools/profile_step.py
  ncu --profile-from-start off --set full --clock-control none --import-source on -k regex:gemm_f16 -s 30 -c 3 \
      -o gpurun_out/gemm python tools/profile_step.py
Numbers printed under a profiler are never bench values."""
import copy
import json
import os
import sys
import warnings

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import torch  # noqa: E402

from unidepth_b200 import UniDepthV1, UniDepthV2  # noqa: E402
from unidepth_b200.synthetic import synthetic_state_dict, synthetic_state_dict_v1  # noqa: E402

warnings.simplefilter("ignore")
# usage: profile_step.py [batch] [default|hires|v1]
WL = sys.argv[2] if len(sys.argv) > 2 else "default"
B = int(sys.argv[1]) if len(sys.argv) > 1 else {"default": 8, "hires": 4, "v1": 16}[WL]
H, W = (1024, 1536) if WL == "hires" else (480, 640)
if WL == "v1":
    cfg = json.load(open(os.path.join(ROOT, "tests", "golden", "config_v1_cnvnxtl.json")))
    model = UniDepthV1(copy.deepcopy(cfg))
    model.load_state_dict(synthetic_state_dict_v1(cfg, 0, device="cuda"), strict=True)
else:
    cfg = json.load(open(os.path.join(ROOT, "tests", "golden", "config_v2_vitl14.json")))
    model = UniDepthV2(copy.deepcopy(cfg))
    model.load_state_dict(synthetic_state_dict(cfg, 0, device="cuda"), strict=True)
model = model.to("cuda").eval()
model.use_cuda_graph = False
rgb = torch.randint(0, 256, (B, 3, H, W), dtype=torch.uint8, device="cuda")
model.infer(rgb)
torch.cuda.synchronize()
torch.cuda.cudart().cudaProfilerStart()
model.infer(rgb)
torch.cuda.synchronize()
torch.cuda.cudart().cudaProfilerStop()
print("done")
