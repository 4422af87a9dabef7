"""b200mm — B200-native (sm_100a) implementation of AntMMF's ViT+BERT contrastive hot path.

Python host (registry-compatible modules, autograd glue) over hand-written CUDA behind a C-ABI (include/b200mm.h).
Import as `import b200mm` (see b200mm.py at the repository root).
"""
from . import _lib, ops  # noqa: F401
from ._lib import B200mmError, LIB_PATH  # noqa: F401

__version__ = "0.1.0"

# importing the package registers the encoders and losses with the (AntMMF or local) registries
from . import contrastive, convert as _convert_mod, cross, distributed, encoders, functional, gradcache, losses, moco, modules, registry, retrieval, video, vtp  # noqa: E402,F401
from .contrastive import clip_contrastive_loss, mil_nce_loss  # noqa: E402,F401
from .distributed import gather_tensor  # noqa: E402,F401
from .modules import CNCLIP, CONFIGS, BertModel, VisionTransformer  # noqa: E402,F401
from .convert import convert  # noqa: E402,F401
