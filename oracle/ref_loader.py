"""TEST INFRASTRUCTURE — not product code.

Loads the UNMODIFIED AntMMF reference modules of the hot path straight from the read-only reference tree
(``/root/reference`` in the build container). ``import antmmf`` itself cannot work here (omegaconf, jsonlines,
torchtext ... are not installed), so bare package objects are pre-seeded in ``sys.modules`` and only
the plain-PyTorch files of the path are imported:

  antmmf/modules/vision/backbone/clip/model.py          (VisionTransformer, ResidualAttentionBlock, LayerNorm, QuickGELU)
  antmmf/modules/vision/backbone/clip/modeling_bert.py  (BertModel and its layers)
  antmmf/modules/vision/backbone/clip/cn_model.py       (CNCLIP, CONFIGS)
  antmmf/utils/distributed_utils.py                     (gather_tensor / GradientAllGather)

The pure loss functions of ``prj/base_vtp/roi_univl/univl/model/univl_video_ret.py`` (module not importable) are
extracted by AST and exec'd without modification (``get_mil_nce_loss`` :146-197, ``get_l1_simi_matrix`` :199-226,
``reduce_clips`` :345-355) and likewise ``MocoUtils.moco_loss`` (``moco_utils.py:71-81``).

Only ``tests/``, ``oracle/make_golden.py`` and ``bench.py --impl reference`` may import this module. The reference
tree does not exist on the GPU box: callers must check :func:`available` first.
"""
import ast
import contextlib
import importlib
import os
import sys
import textwrap
import types

REF_ROOT = os.environ.get("B200MM_REFERENCE_ROOT", "/root/reference")


def available() -> bool:
    return os.path.isfile(os.path.join(REF_ROOT, "antmmf/modules/vision/backbone/clip/cn_model.py"))


_loaded = {}


def _pkg(name, path):
    m = types.ModuleType(name)
    m.__path__ = [path]
    sys.modules[name] = m
    return m


def load():
    """Returns a namespace with the reference modules: .model, .modeling_bert, .cn_model, .distributed_utils."""
    if "ns" in _loaded:
        return _loaded["ns"]
    if not available():
        raise RuntimeError(f"reference tree not found under {REF_ROOT}")
    R = REF_ROOT
    if "antmmf" in sys.modules and not getattr(sys.modules["antmmf"], "_b200mm_stub", False):
        raise RuntimeError("a real 'antmmf' package is already imported; the stub loader would shadow it")
    a = _pkg("antmmf", R + "/antmmf")
    a._b200mm_stub = True
    c = _pkg("antmmf.common", R + "/antmmf/common")
    c.configurable = lambda f=None, **kw: f if f is not None else (lambda g: g)
    c.Configuration = type("Configuration", (dict,), {})
    for p in ["modules", "modules/vision", "modules/vision/backbone", "modules/vision/backbone/clip", "utils"]:
        _pkg("antmmf." + p.replace("/", "."), R + "/antmmf/" + p)
    g = types.ModuleType("antmmf.utils.general")
    g.nullcontext = contextlib.nullcontext
    sys.modules["antmmf.utils.general"] = g
    # modeling_bert.py:23-24 imports the flash-attn 1.x symbol whenever flash_attn is importable
    fa = types.ModuleType("flash_attn.flash_attention")
    fa.FlashMHA = None
    sys.modules["flash_attn.flash_attention"] = fa
    ns = types.SimpleNamespace()
    ns.cn_model = importlib.import_module("antmmf.modules.vision.backbone.clip.cn_model")
    ns.model = importlib.import_module("antmmf.modules.vision.backbone.clip.model")
    ns.modeling_bert = importlib.import_module("antmmf.modules.vision.backbone.clip.modeling_bert")
    ns.configuration_bert = importlib.import_module("antmmf.modules.vision.backbone.clip.configuration_bert")
    ns.distributed_utils = importlib.import_module("antmmf.utils.distributed_utils")
    _loaded["ns"] = ns
    return ns


def _extract_functions(relpath, names, extra_ns=None):
    import numpy as np
    import torch
    import torch.nn.functional as F

    src = open(os.path.join(REF_ROOT, relpath)).read()
    tree = ast.parse(src)
    ns = {"torch": torch, "np": np, "F": F, "get_package_version": lambda _: torch.__version__}
    if extra_ns:
        ns.update(extra_ns)
    found = {}
    for node in ast.walk(tree):
        if isinstance(node, ast.FunctionDef) and node.name in names and node.name not in found:
            code = textwrap.dedent(ast.get_source_segment(src, node))
            exec(compile(code, relpath, "exec"), ns)
            found[node.name] = ns[node.name]
    missing = set(names) - set(found)
    if missing:
        raise RuntimeError(f"functions {missing} not found in {relpath}")
    return found


def load_loss_functions():
    """Reference loss arithmetic, unmodified, callable with ``self=None`` (or a namespace with ``.T`` for moco)."""
    if "loss" in _loaded:
        return _loaded["loss"]
    ret = _extract_functions(
        "prj/base_vtp/roi_univl/univl/model/univl_video_ret.py",
        ["get_mil_nce_loss", "get_l1_simi_matrix", "reduce_clips"],
    )
    ret.update(_extract_functions("prj/base_vtp/roi_univl/univl/model/moco_utils.py", ["moco_loss"]))
    ns = types.SimpleNamespace(**ret)
    _loaded["loss"] = ns
    return ns


def load_retrieval_metric_functions():
    """Reference retrieval-metric arithmetic, unmodified (antmmf/modules/metrics/global_retrieval_recall.py:12-102)."""
    if "ret_metric" in _loaded:
        return _loaded["ret_metric"]
    fns = _extract_functions("antmmf/modules/metrics/global_retrieval_recall.py", ["_compute_retrieval_metrics", "_cal_sym_recall", "_cal_recall"])
    # the three functions call each other through module globals: they share the exec namespace, so this already resolves
    ns = types.SimpleNamespace(**fns)
    _loaded["ret_metric"] = ns
    return ns


def build_cnclip(name_or_cfg, seed=0, dropout=0.0):
    """Reference CNCLIP with the reference initialisation; ``text_projection`` (torch.empty in cn_model.py:190-192)
    gets N(0, hidden^-0.5). Dropout probabilities are overridden (parity runs use 0)."""
    import torch

    ns = load()
    cfg = dict(ns.cn_model.CONFIGS[name_or_cfg]) if isinstance(name_or_cfg, str) else dict(name_or_cfg)
    cfg["text_attention_probs_dropout_prob"] = dropout
    cfg["text_hidden_dropout_prob"] = dropout
    torch.manual_seed(seed)
    m = ns.cn_model.CNCLIP(**cfg)
    with torch.no_grad():
        m.text_projection.normal_(0.0, cfg["text_hidden_size"] ** -0.5)
    return m


def load_m2():
    """Unmodified M²-Encoder (BEiT-3 multiway) reference modules from ``prj/M2_Encoder`` (SURVEY.md §8c recipe):
    .BEiT3 (vlmo/torchscale/model/BEiT3.py), .Encoder / .EncoderLayer (architecture/encoder.py), .EncoderConfig
    (architecture/config.py), .heads (vlmo/modules/heads.py). ``fairscale``, ``pytorch_lightning`` and ``timm`` are
    not installed: they are replaced by identity stubs (checkpoint wrappers, rank_zero_info, drop_path at rate 0 —
    none of them does arithmetic on the path). ``vlmo`` / ``vlmo.modules`` are pre-seeded as bare packages because
    ``vlmo/modules/__init__.py`` would import the Lightning module (tokenizers, timm model zoo)."""
    if "m2" in _loaded:
        return _loaded["m2"]
    if not available():
        raise RuntimeError(f"reference tree not found under {REF_ROOT}")
    import torch

    root = os.path.join(REF_ROOT, "prj", "M2_Encoder")

    def stub(name, **attrs):
        m = sys.modules.get(name)
        if m is None:
            m = types.ModuleType(name)
            m.__path__ = []
            sys.modules[name] = m
        for k, v in attrs.items():
            setattr(m, k, v)
        return m

    ident = lambda module, *a, **k: module  # noqa: E731
    if "fairscale" not in sys.modules:
        stub("fairscale")
        stub("fairscale.nn", checkpoint_wrapper=ident, wrap=ident)
    if "pytorch_lightning" not in sys.modules:
        stub("pytorch_lightning")
        stub("pytorch_lightning.utilities")
        stub("pytorch_lightning.utilities.distributed", rank_zero_info=lambda *a, **k: None)
    if "timm" not in sys.modules:

        def drop_path(x, drop_prob=0.0, training=False):
            if drop_prob != 0.0 and training:
                raise NotImplementedError("stub: drop_path only at rate 0")
            return x

        stub("timm")
        stub("timm.models")
        stub("timm.models.layers", drop_path=drop_path, trunc_normal_=torch.nn.init.trunc_normal_)
    _pkg("vlmo", root + "/vlmo")
    _pkg("vlmo.modules", root + "/vlmo/modules")
    ns = types.SimpleNamespace()
    ns.BEiT3 = importlib.import_module("vlmo.torchscale.model.BEiT3").BEiT3
    enc = importlib.import_module("vlmo.torchscale.architecture.encoder")
    ns.Encoder, ns.EncoderLayer = enc.Encoder, enc.EncoderLayer
    ns.EncoderConfig = importlib.import_module("vlmo.torchscale.architecture.config").EncoderConfig
    ns.heads = importlib.import_module("vlmo.modules.heads")
    ns.multiway = importlib.import_module("vlmo.torchscale.component.multiway_network")
    _loaded["m2"] = ns
    return ns


def load_stage2():
    """Unmodified stage-2 (cross-modal) retrieval arithmetic of prj/base_vtp, AST-extracted because the modules import the whole
    antmmf package: `_cross_similarity` (univl_video_ret.py:33-89), `_cross_similarity_hard_mining` (:91-144), `get_mil_nce_loss`
    (:146-197), `get_cross_output` / `_align_text_to_video_clips` (univl_video_base.py:229-307), `split_encoder_output`
    (univl_base.py:17-36). `gather_tensor` / `all_gather` / `get_rank` come from the real antmmf/utils/distributed_utils.py
    (single process: identity / [x] / 0). Returns a namespace of plain functions taking `self` first."""
    if "stage2" in _loaded:
        return _loaded["stage2"]
    ns0 = load()
    du = ns0.distributed_utils
    split = _extract_functions("prj/base_vtp/roi_univl/univl/model/univl_base.py", ["split_encoder_output"])
    base = _extract_functions("prj/base_vtp/roi_univl/univl/model/univl_video_base.py", ["get_cross_output", "_align_text_to_video_clips"],
                              extra_ns=split)
    ret = _extract_functions("prj/base_vtp/roi_univl/univl/model/univl_video_ret.py",
                             ["_cross_similarity", "_cross_similarity_hard_mining", "get_mil_nce_loss"],
                             extra_ns=dict(gather_tensor=du.gather_tensor, all_gather=du.all_gather, get_rank=du.get_rank))
    ns = types.SimpleNamespace(**split, **base, **ret)
    _loaded["stage2"] = ns
    return ns


def build_stage2_self(hidden=64, heads=2, layers=2, inter=128, out_dim=32, seed=0, re_sample_method="top_k"):
    """A stand-in for `UnivlForVideoTextRetrieval` holding exactly the members the extracted stage-2 functions touch:
    .module.{cross_encoder = reference BertEncoder, arch_type 'clip', text_encoder.text_projection, get_cross_output, …},
    .similarity_dense (univl_video_ret.py:24-28), .dropout (p set to 0 for parity), .config.re_sample_method."""
    import torch
    from torch import nn

    ns = load()
    s2 = load_stage2()
    torch.manual_seed(seed)
    cfg = ns.configuration_bert.BertConfig(vocab_size_or_config_json_file=64, hidden_size=hidden, num_hidden_layers=layers,
                                           num_attention_heads=heads, intermediate_size=inter, hidden_act="gelu", hidden_dropout_prob=0.0,
                                           attention_probs_dropout_prob=0.0, max_position_embeddings=64)
    module = nn.Module()
    module.cross_encoder = ns.modeling_bert.BertEncoder(cfg)
    for m in module.cross_encoder.modules():
        if isinstance(m, nn.Linear):
            m.weight.data.normal_(0.0, 0.05)
            m.bias.data.normal_(0.0, 0.05)
    module.arch_type = "clip"
    module.text_encoder = nn.Module()
    module.text_encoder.text_projection = nn.Parameter(torch.randn(hidden, out_dim) * hidden ** -0.5)
    module.get_cross_output = types.MethodType(s2.get_cross_output, module)
    module._align_text_to_video_clips = types.MethodType(s2._align_text_to_video_clips, module)
    me = nn.Module()
    me.module = module
    me.dropout = nn.Dropout(0.0)
    me.similarity_dense = nn.Sequential(nn.Linear(out_dim, out_dim * 2), nn.ReLU(True), nn.Linear(out_dim * 2, 1))
    me.config = types.SimpleNamespace(re_sample_method=re_sample_method, hidden_size=out_dim)
    for name in ["_cross_similarity", "_cross_similarity_hard_mining", "get_mil_nce_loss"]:
        setattr(me, name, types.MethodType(getattr(s2, name), me))
    return me
