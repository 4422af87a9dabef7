"""Imports the UNMODIFIED reference CNCLIP from baseline/_ref/ (see install_ref.py) for the two baseline legs of bench.py:
`--impl reference` / `cpu_baseline` (host cores, fp32) and `gpu_eager_baseline` (the same modules after .cuda().bfloat16()).

`import antmmf` as a package is impossible in this image (omegaconf, jsonlines, torchtext ... missing), so bare package
objects are pre-seeded in sys.modules and only the plain-PyTorch files of the path are imported (cn_model.py, model.py,
modeling_bert.py, configuration_bert.py, cn_tokenizer.py). Nothing of the reference is edited or re-implemented here."""
import contextlib
import importlib
import os
import sys
import types

REF_DIR = os.path.join(os.path.dirname(os.path.abspath(__file__)), "_ref")
_CN = "antmmf/modules/vision/backbone/clip/cn_model.py"


def available():
    return os.path.isfile(os.path.join(REF_DIR, _CN))


def _pkg(name, path):
    m = types.ModuleType(name)
    m.__path__ = [path]
    sys.modules[name] = m
    return m


def load_cn_model():
    if not available():
        raise RuntimeError(f"reference copies not found under {REF_DIR} (python baseline/install_ref.py)")
    mod = sys.modules.get("antmmf.modules.vision.backbone.clip.cn_model")
    if mod is not None:
        return mod
    if "antmmf" in sys.modules and not getattr(sys.modules["antmmf"], "_b200mm_stub", False):
        raise RuntimeError("a real 'antmmf' package is already imported")
    a = _pkg("antmmf", REF_DIR + "/antmmf")
    a._b200mm_stub = True
    c = _pkg("antmmf.common", REF_DIR + "/antmmf/common")
    c.configurable = lambda f=None, **kw: f if f is not None else (lambda g: g)
    c.Configuration = type("Configuration", (dict,), {})
    for p in ["modules", "modules/vision", "modules/vision/backbone", "modules/vision/backbone/clip", "utils"]:
        _pkg("antmmf." + p.replace("/", "."), REF_DIR + "/antmmf/" + p)
    g = types.ModuleType("antmmf.utils.general")
    g.nullcontext = contextlib.nullcontext
    sys.modules["antmmf.utils.general"] = g
    fa = types.ModuleType("flash_attn.flash_attention")  # modeling_bert.py:23-24 probes the flash-attn 1.x symbol
    fa.FlashMHA = None
    sys.modules["flash_attn.flash_attention"] = fa
    return importlib.import_module("antmmf.modules.vision.backbone.clip.cn_model")


def build_cnclip(name, seed=0, dropout=0.0):
    """Reference CNCLIP(**CONFIGS[name]) with the reference initialisation; `text_projection` (torch.empty, cn_model.py:190-192) gets
    N(0, hidden^-0.5); dropout probabilities as given (0 for the timing legs, like the B200 arm)."""
    import torch

    cn = load_cn_model()
    cfg = dict(cn.CONFIGS[name])
    cfg["text_attention_probs_dropout_prob"] = dropout
    cfg["text_hidden_dropout_prob"] = dropout
    torch.manual_seed(seed)
    m = cn.CNCLIP(**cfg)
    with torch.no_grad():
        m.text_projection.normal_(0.0, cfg["text_hidden_size"] ** -0.5)
    return m
