"""B200 side of the reference's self-supervised path (src/self_supervised/dinov2/): the losses of SURVEY.md 8f row f2 on
the row kernels of csrc/ssl.cu.  The backbone of that path is `apla_b200.apla` (packed multi-crop FusedAplaBlock)."""
from .dino_head import DINOHead
from .loss import DINOLoss, KoLeoLoss, iBOTPatchLoss, update_teacher

__all__ = ["DINOHead", "DINOLoss", "iBOTPatchLoss", "KoLeoLoss", "update_teacher"]
