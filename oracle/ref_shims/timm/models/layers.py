from torch.nn.init import trunc_normal_  # noqa: F401
from timm.layers import DropPath  # noqa: F401
