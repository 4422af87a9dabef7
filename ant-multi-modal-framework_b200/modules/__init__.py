from .bert import BertConfig, BertEmbeddings, BertEncoder, BertLayer, BertModel  # noqa: F401
from .cnclip import CONFIGS, CNCLIP  # noqa: F401
from .vit import ResidualAttentionBlock, Transformer, VisionTransformer  # noqa: F401
