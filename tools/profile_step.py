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

from unidepth_b200 import UniDepthV2  # noqa: E402
from unidepth_b200.synthetic import synthetic_state_dict  # noqa: E402

warnings.simplefilter("ignore")
B = int(sys.argv[1]) if len(sys.argv) > 1 else 8
cfg = json.load(open(os.path.join(ROOT, "tests", "golden", "config_v2_vitl14.json")))
model = UniDepthV2(copy.deepcopy(cfg))
model.load_state_dict(synthetic_state_dict(cfg, 0, device="cuda"), strict=True)
model = model.to("cuda").eval()
model.use_cuda_graph = False
rgb = torch.randint(0, 256, (B, 3, 480, 640), dtype=torch.uint8, device="cuda")
model.infer(rgb)
torch.cuda.synchronize()
torch.cuda.cudart().cudaProfilerStart()
model.infer(rgb)
torch.cuda.synchronize()
torch.cuda.cudart().cudaProfilerStop()
print("done")
