"""base_vtp video-text retrieval model on the b200mm path — the caller of rows a9–a16 and f3, arch_type 'clip'.

Mirrors, with the same attribute names / state-dict prefixes / method signatures / output dictionaries:
  UnivlVideoBase              prj/base_vtp/roi_univl/univl/model/univl_video_base.py:14-316
      .text_encoder, .img_encoder (built through the TextEncoder / VisualEncoder registries from config.text_encoder /
      config.image_encoder, :24-29), .img_proj, .cross_embeddings / .cross_encoder (= the text encoder's, :47-48),
      forward_img_encoder, forward_text_encoder, prepare_cross_text, prepare_cross_visual, build_transformer_input,
      get_cross_output, get_l2_input
  UnivlForVideoTextRetrieval  prj/base_vtp/roi_univl/univl/model/univl_video_ret.py:16-470
      .module, .similarity_dense, .moco_utils, forward(img_input, caption_input) -> {"losses": {"level1_similarity_loss",
      "level2_similarity_loss"}, "l1_simi", "l2_simi"}, forward_stage1 / forward_stage2 / get_simi_logits

Everything numeric runs in the b200mm kernels through the modules this file composes (video.py, contrastive.py, moco.py, cross.py);
the only tensor work done here is index / mask plumbing. Differences from the reference, all raising or documented:
arch_type 'univl' (needs attention probabilities and the HuggingFace BertModel) is not built; dropout is p = 0; with MoCo only
n_clips == 1 (the fused queue loss takes one positive per row); the level-1 matrix handed to hard mining is computed once by a plain
GEMM under no_grad (the reference detaches its clone too, :395-397).
"""
import torch
from torch import nn

from . import functional as Fn
from . import ops
from .contrastive import mil_nce_loss
from .cross import PairScorer
from .distributed import gather_tensor, gathered_sizes, get_rank, get_world_size
from .moco import B200MocoUtils
from .registry import TextEncoder, VisualEncoder
from .video import forward_img_encoder as _forward_img_encoder
from .video import forward_text_encoder as _forward_text_encoder

BF16 = torch.bfloat16


def _get(config, key, default=None):
    if hasattr(config, "get"):
        return config.get(key, default)
    return getattr(config, key, default)


def _bf16(t):
    return t if t.dtype == BF16 else t.to(BF16)


class B200UnivlVideoBase(nn.Module):
    def __init__(self, config, **kwargs):
        super().__init__()
        self.config = config
        self.arch_type = _get(config, "arch_type", "clip")
        if self.arch_type != "clip":
            raise NotImplementedError("b200mm UnivlVideoBase: only arch_type='clip' (ViT + BERT, univl_video_base.py:144-149,252-258)")
        self.with_cross_encoder = kwargs.get("with_cross_encoder", None)
        if self.with_cross_encoder is None:
            self.with_cross_encoder = _get(config, "with_cross_encoder", False)
        self.text_encoder = TextEncoder(_get(config, "text_encoder")).module
        self.img_encoder = VisualEncoder(_get(config, "image_encoder")).module
        hidden = _get(config, "hidden_size")
        self.img_proj = None
        if self.img_encoder.out_dim != hidden:
            self.img_proj = nn.Parameter(torch.randn(self.img_encoder.out_dim, hidden) * self.img_encoder.out_dim ** -0.5)
        # shared with the text tower, exactly as the reference aliases them (:47-48); plain attributes, not re-registered
        object.__setattr__(self, "cross_embeddings", self.text_encoder.embeddings)
        object.__setattr__(self, "cross_encoder", self.text_encoder.encoder)

    # ---- :56-166
    def forward_img_encoder(self, image_data, image_pad_mask, image_n_clips, image_num_frames, img_encoder=None, **kwargs):
        return _forward_img_encoder(img_encoder or self.img_encoder, image_data, image_pad_mask, image_n_clips, image_num_frames, self.img_proj)

    def forward_text_encoder(self, input_ids, input_mask, txt_encoder=None):
        return _forward_text_encoder(txt_encoder or self.text_encoder, input_ids, input_mask, self.arch_type)

    # ---- :168-227
    def prepare_cross_text(self, input_ids, input_mask):
        cap_embed = self.cross_embeddings(input_ids=input_ids, token_type_ids=torch.zeros_like(input_ids))
        return cap_embed, input_mask, cap_embed.shape[0]

    def prepare_cross_visual(self, visual_embed, visual_mask=None):
        bsz, num_clip = visual_embed.shape[0], visual_embed.shape[1]
        if visual_mask is None:
            visual_mask = torch.zeros((bsz, num_clip), device=visual_embed.device).bool()
        sep = torch.full((bsz,), 102, dtype=torch.long, device=visual_embed.device)
        sep_token_embeds = self.cross_embeddings.word_embeddings(sep).unsqueeze(1)
        visual_mask = visual_mask.logical_not().long()
        visual_inputs_embeds = torch.cat([_bf16(visual_embed), _bf16(sep_token_embeds)], 1)
        token_type_ids = torch.ones(visual_inputs_embeds.shape[:2], dtype=torch.long, device=visual_embed.device)
        new_visual_embed = self.cross_embeddings(inputs_embeds=visual_inputs_embeds, token_type_ids=token_type_ids)
        new_visual_mask = torch.cat([visual_mask, visual_mask.new_ones((bsz, 1))], 1)
        return new_visual_embed, new_visual_mask, num_clip

    def build_transformer_input(self, visual_embed_dict, text_embed_dict, caption_input):
        cap_embed, cap_mask, batch_size = self.prepare_cross_text(caption_input["caption_input_ids"], caption_input["caption_input_mask"])
        visual_embed, visual_mask, num_clip = self.prepare_cross_visual(visual_embed_dict["visual_embed"], visual_embed_dict["visual_mask"])
        return cap_embed, visual_embed, cap_mask, visual_mask, num_clip, batch_size

    # ---- :229-271 (aligned pairs, n_clips = 1 after fusion)
    def get_cross_output(self, cap_embed, visual_embed, cap_mask, visual_mask, n_clips):
        if n_clips > 1:
            cap_embed = cap_embed.repeat_interleave(n_clips, dim=0)
            cap_mask = cap_mask.repeat_interleave(n_clips, dim=0)
        embed = torch.cat([_bf16(cap_embed), _bf16(visual_embed)], 1)
        mask = torch.cat([cap_mask, visual_mask], 1)
        ext = (1.0 - mask.unsqueeze(1).unsqueeze(2).float()) * -10000.0
        seq = self.cross_encoder(embed, attention_mask=ext, head_mask=[None] * len(self.cross_encoder.layer))[0]
        B, S, H = seq.shape
        proj = self.text_encoder.text_projection
        if proj is not None:
            pooled = Fn.ClsHeadFn.apply(seq.reshape(B * S, H), None, None, _bf16(proj), B, S, 0.0)
        else:
            pooled = seq[:, 0, :]
        St = cap_embed.shape[1]
        return seq[:, :St], seq[:, St:-1], pooled

    # ---- :273-300
    def get_l2_input(self, img_input, caption_input):
        visual_embed_dict = self.forward_img_encoder(**img_input)
        text_embed_dict = self.forward_text_encoder(caption_input["caption_raw_input_ids"], caption_input["caption_input_mask"])
        if self.with_cross_encoder:
            cap_embed, visual_embed, cap_mask, visual_mask, num_clips, batch_size = self.build_transformer_input(
                visual_embed_dict, text_embed_dict, caption_input)
        else:  # stage 1 only: the cross-encoder inputs are never consumed
            cap_embed = visual_embed = None
            cap_mask, visual_mask = caption_input["caption_input_mask"], None
            num_clips, batch_size = visual_embed_dict["visual_embed"].shape[1], caption_input["caption_input_mask"].shape[0]
        cap_input = (cap_embed, cap_mask, text_embed_dict["pooled_output"], batch_size)
        vis_input = (visual_embed, visual_mask, visual_embed_dict["clip_feature"], num_clips)
        return cap_input, vis_input, text_embed_dict, visual_embed_dict


class B200VideoTextRetrieval(nn.Module):
    """UnivlForVideoTextRetrieval (univl_video_ret.py:16-470). config keys as in prj/base_vtp/configs/univl/video/*: training_stage,
    hidden_size, text_encoder, image_encoder, with_moco, K / M / T, hard_example_mining, re_sample_method, re_weight_method."""

    def __init__(self, config, max_pairs=8192):
        super().__init__()
        self.config = config
        stage = _get(config, "training_stage", "stage1")
        self.training_stage = stage
        with_cross = "stage2" in stage
        self.module = B200UnivlVideoBase(config, with_cross_encoder=with_cross)
        if with_cross:
            h = _get(config, "hidden_size")
            self.similarity_dense = nn.Sequential(nn.Linear(h, h * 2), nn.ReLU(True), nn.Linear(h * 2, 1))
        self.with_moco = _get(config, "with_moco", True)
        self.moco_utils = None
        self.max_pairs = max_pairs

    # ---- level 1 -----------------------------------------------------------------------------------------------------------
    @torch.no_grad()
    def _l1_matrix(self, text_embed_l1, video_embed_l1, num_clips, cal_cross=True):
        """get_simi_logits level 'l1' + get_l1_simi_matrix + reduce_clips (:199-226, :313-327, :345-355), no gradient (only the detached
        copy is consumed downstream: hard mining / metrics).
          training, cal_cross : embeddings of all ranks are gathered first (:313-325) -> [Bg_text, Bg_video]
          eval, cal_cross     : the local matrix [bsz_text, bsz_video] (no collective: ranks may evaluate different shards)
          cal_cross == False  : inputs are aligned pairs -> the paired clip scores reduced over clips, [bsz_text]"""
        t, v = text_embed_l1, video_embed_l1
        if not cal_cross:
            own = (t.float()[:, None, :] * v.view(t.shape[0], num_clips, -1).float()).sum(-1)  # [bsz_text, n_clips]
            return own.logsumexp(-1)
        if self.training and get_world_size() > 1:
            t = gather_tensor(t, method="cat", back_gradient=False, pad_tensors=True)
            v = gather_tensor(v, method="cat", back_gradient=False, pad_tensors=True)
        nv = v.shape[0]
        pad = (-nv) % 8
        vp = torch.cat([v, v.new_zeros(pad, v.shape[1])]) if pad else v
        sim = ops.gemm(_bf16(t).contiguous(), _bf16(vp).contiguous(), out_f32=True)[:, :nv]
        if num_clips > 1:  # logsumexp over the clips of a video (small [Bg, Bg, n] tensor)
            sim = sim.reshape(t.shape[0], nv // num_clips, num_clips).logsumexp(-1)
        return sim

    def _moco_l1_loss(self, vis_input, cap_input):
        """get_simi_logits level 'l1' with MoCo (:263-312)."""
        (_, _, text_embed_l1, _, caption_input) = cap_input
        (_, _, video_embed_l1, num_clips, img_input) = vis_input
        if num_clips != 1:
            raise NotImplementedError("b200mm: the fused MoCo queue loss takes one positive per row (n_clips == 1)")
        if self.moco_utils is None:
            cfg = dict(hidden_size=_get(self.config, "hidden_size"), K=_get(self.config, "K", 16384), M=_get(self.config, "M", 0.9999),
                       T=_get(self.config, "T", 0.05))
            self.moco_utils = B200MocoUtils(cfg, img_encoder=self.module.img_encoder, txt_encoder=self.module.text_encoder).to(text_embed_l1.device)
        mu = self.moco_utils
        with torch.no_grad():
            mu.momentum_update_key_encoder()
            key_v = self.module.forward_img_encoder(**img_input, img_encoder=mu.img_encoder_k)["clip_feature"]
            key_t = self.module.forward_text_encoder(caption_input["caption_raw_input_ids"], caption_input["caption_input_mask"],
                                                     txt_encoder=mu.txt_encoder_k)["pooled_output"]
        loss_v = mu.moco_loss_fused(video_embed_l1, key_t, "txt")
        loss_t = mu.moco_loss_fused(text_embed_l1, key_v, "img")
        mu.dequeue_and_enqueue(key_v, key_t)
        return (loss_t + loss_v) / 2.0

    def forward_stage1(self, vis_input, cap_input, output_dict=None, cal_cross=True):
        output_dict = dict(losses={}) if output_dict is None else output_dict
        (_, _, text_embed_l1, _, _) = cap_input
        (_, _, video_embed_l1, num_clips, _) = vis_input
        if self.training and self.with_moco:
            loss = self._moco_l1_loss(vis_input, cap_input)
        elif cal_cross and text_embed_l1.shape[0] * num_clips == video_embed_l1.shape[0]:
            # forward_stage1 :369-381: MIL-NCE over the gathered batch; the [B·n, B·n] repeat is never built
            # (under data parallelism contrastive.py returns W x the rank's share: the mean over ranks is the reference's global loss and
            # DDP's gradient averaging yields its exact gradient)
            # in eval the reference does not gather (:313): the loss is that of the local matrix
            loss = mil_nce_loss(video_embed_l1, text_embed_l1, None if self.training else "local", n_clips=num_clips)
        else:
            loss = text_embed_l1.new_zeros((), dtype=torch.float32)
        output_dict["losses"]["level1_similarity_loss"] = loss
        output_dict["l1_simi"] = self._l1_matrix(text_embed_l1, video_embed_l1, num_clips, cal_cross)
        return output_dict

    # ---- level 2 -----------------------------------------------------------------------------------------------------------
    def _scorer(self):
        return PairScorer(self.module.text_encoder, self.similarity_dense, self.max_pairs)

    def _cross_similarity(self, sequence_output, visual_output, attention_mask, video_mask, num_clips):
        return self._scorer().cross_similarity(sequence_output, visual_output, attention_mask, video_mask, num_clips)

    def _cross_similarity_hard_mining(self, vis_input, cap_input, l1_simi_matrix):
        return self._scorer().cross_similarity_hard_mining(vis_input, cap_input, l1_simi_matrix, _get(self.config, "re_sample_method", "top_k"))

    def forward_stage2(self, vis_input, cap_input, output_dict=None, cal_cross=True):
        output_dict = dict(losses={}) if output_dict is None else output_dict
        (cap_embed, cap_mask, _, batch_size, _) = cap_input
        (visual_embed, visual_mask, _, num_clips, _) = vis_input
        hard = self.training and _get(self.config, "hard_example_mining", False)
        scorer = self._scorer()
        if hard:
            l1 = output_dict["l1_simi"].detach()
            l2_simi = self._cross_similarity_hard_mining(vis_input, cap_input, l1)
        elif cal_cross:
            l2_simi = self._cross_similarity(cap_embed, visual_embed, cap_mask, visual_mask, num_clips)
        else:  # aligned pairs only (inference, :243-248)
            n = cap_embed.shape[0]
            idx = torch.arange(n, device=cap_embed.device)
            l2_simi = scorer.score_pair_list(cap_embed, cap_mask, visual_embed, visual_mask, idx, idx).view(-1, 1)
        if cal_cross and l2_simi.shape[0] == l2_simi.shape[1]:
            weighted = hard and _get(self.config, "re_weight_method", None) == "median"
            # row offset of this rank inside the gathered level-1 matrix = sum of the batch sizes of the ranks before it (:436-438)
            beg_idx = sum(gathered_sizes(cap_embed)[: get_rank()]) if weighted and get_world_size() > 1 else 0
            loss = scorer.level2_loss(l2_simi, output_dict["l1_simi"] if weighted else None, beg_idx,
                                      "median" if weighted else None, _get(self.config, "re_sample_method", "top_k"))
        else:
            loss = l2_simi.new_zeros(())
        output_dict["losses"]["level2_similarity_loss"] = loss
        output_dict["l2_simi"] = l2_simi
        return output_dict

    def forward_stage(self, cap_input, vis_input, cal_cross=True):
        output_dict = None
        if "stage1" in self.training_stage:
            output_dict = self.forward_stage1(vis_input, cap_input, output_dict, cal_cross=cal_cross)
        if "stage2" in self.training_stage:
            output_dict = self.forward_stage2(vis_input, cap_input, output_dict, cal_cross=cal_cross)
        return output_dict

    def forward(self, img_input, caption_input, ocr_input=None, region_input=None, caption_output=None, sample_list=None):
        cap_input, vis_input, _, _ = self.module.get_l2_input(img_input, caption_input)
        return self.forward_stage(cap_input + (caption_input,), vis_input + (img_input,), True)


# ======================================================================================================================
# Model plug-in: antmmf `registry.register_model` API (antmmf/common/registry.py:415-440, models/build.py:9-25)
# ======================================================================================================================
try:  # pragma: no cover - only with a full AntMMF installation
    from antmmf.models.base_model import BaseModel as _BaseModel  # type: ignore
except Exception:  # noqa: BLE001
    _BaseModel = None


class B200Univl(_BaseModel if _BaseModel is not None else nn.Module):
    """`Univl` (prj/base_vtp/roi_univl/univl/model/univl_model.py:16-125) for training_head_type 'video_text_retrieval': same
    build() / group_inputs() / forward(sample_list) -> {"losses": {...}, "l1_simi", "l2_simi"} / get_optimizer_parameters(config).
    Registered as "b200_univl"; `install_as_univl()` also claims the key "univl", which the unmodified trainer requires
    (base_trainer.py:552-556 reads config.model_attributes.univl on every iteration)."""

    def __init__(self, config):
        if _BaseModel is not None:
            super().__init__(config)
        else:
            super().__init__()
            self.config = config

    def build(self):
        head = _get(self.config, "training_head_type", "video_text_retrieval")
        if head != "video_text_retrieval":
            raise NotImplementedError(f"b200mm Univl: training_head_type {head!r}; only 'video_text_retrieval' is on the hot path")
        self.model = B200VideoTextRetrieval(self.config)
        self.get_l2_input = self.model.module.get_l2_input  # modality features for retrieval evaluation (univl_model.py:30-31)

    @staticmethod
    def group_inputs(sample_list):
        groups = {"ocr": None, "caption": None, "region": None, "image": None, "generation": None}
        for sample_key in sample_list.keys():
            for input_key in groups:
                if sample_key.startswith(input_key):
                    if groups[input_key] is None:
                        groups[input_key] = {}
                    groups[input_key][sample_key] = sample_list[sample_key]
        return groups

    def forward(self, sample_list, *args, **kwargs):
        g = self.group_inputs(sample_list)
        out = self.model(g["image"], g["caption"], g["ocr"], g["region"], caption_output=g["generation"], sample_list=sample_list)
        return {"logits": out} if isinstance(out, torch.Tensor) else out

    def get_optimizer_parameters(self, config):
        """UnivlForVideoTextRetrieval.get_optimizer_parameters (univl_video_ret.py:478-537): pretrained tower parameters at
        lr * encoder_lr_decay, new modules at lr; biases / LayerNorm affine without weight decay."""
        lr = config.optimizer_attributes.params.lr
        weight_decay = config.optimizer_attributes.params.weight_decay
        decay = _get(self.config, "encoder_lr_decay", 0.01)
        no_decay = ["bias", "LayerNorm.bias", "LayerNorm.weight"]
        pretrain_prefix = ["text_encoder.embeddings.", "text_encoder.encoder.", "text_encoder.pooler.", "img_embeddings.", "img_encoder."]
        groups = {(d, c): [] for d in (True, False) for c in (True, False)}
        for n, p in self.model.named_parameters():
            groups[(not any(nd in n for nd in no_decay), any(pre in n for pre in pretrain_prefix))].append(p)
        return [
            {"params": groups[(True, True)], "weight_decay": weight_decay, "lr": lr * decay},
            {"params": groups[(True, False)], "weight_decay": weight_decay},
            {"params": groups[(False, True)], "weight_decay": 0.0, "lr": lr * decay},
            {"params": groups[(False, False)], "weight_decay": 0.0},
        ]


from .registry import registry as _registry  # noqa: E402

_registry.register_model("b200_univl")(B200Univl)


def install_as_univl():
    """Claim the model key "univl" (the only key the unmodified trainer can run, see the class docstring)."""
    _registry.register_model("univl")(B200Univl)
