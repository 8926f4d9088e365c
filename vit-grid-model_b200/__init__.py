"""vit_grid_model_b200 -- B200 (sm_100a) implementation of the MaxViT / MetNet3 grid-forecast hot path of
jhsk777/VIT-Grid-Model behind the reference's own nn.Module API.

    from vit_grid_model_b200 import MetNet3, MaxViT, FocalRLoss

All compute runs in libvitgrid.so (hand-written CUDA, C ABI in include/vitgrid.h); there is no CPU fallback.
"""
from .maxvit import MaxViT, Attention, MBConv          # noqa: F401
from .metnet3 import MetNet3, MetNet3_with_stn_imgs     # noqa: F401
from .focal_r import FocalRLoss, focal_r_loss           # noqa: F401
from ._lib import VitGridError                          # noqa: F401
from .parallel import DataParallel                      # noqa: F401
from .optim import FlatAdamW                            # noqa: F401
from .pipeline import HostPipeline, pack_host           # noqa: F401
from .eval_metrics import EvalMetrics                   # noqa: F401

__version__ = "0.1.0"
