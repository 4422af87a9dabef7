from .bert import BertConfig, BertEmbeddings, BertEncoder, BertLayer, BertModel  # noqa: F401
from .cnclip import CONFIGS, CNCLIP, CNCLIPImageEncoder, CNCLIPLanguageEncoder, available_models, build_model, load  # noqa: F401
from .vit import ResidualAttentionBlock, Transformer, VisionTransformer  # noqa: F401
from .beit3 import BEiT3, M2_CONFIGS, M2Encoder  # noqa: F401
from .beit3 import Encoder as M2TransformerEncoder  # noqa: F401
