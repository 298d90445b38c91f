"""UniBind — top-level multimodal module, mirror of lhrs/models/UniBind.py:24-242.

Same constructor, attribute names (``rgb``, ``rgb_pooler``, ``text``), ``forward`` / ``generate`` / ``encode_image`` /
``prepare_for_training`` / ``custom_load_state_dict`` / ``custom_save_checkpoint`` signatures and state-dict keys as the
reference, so ``main_pretrain_stage{1,2,3}.py`` and ``cli_qa.py`` call it unchanged.  One deliberate deviation, stated in
SURVEY.md §0 fact 11 / BASELINE.json: ``freeze_text=True`` really freezes the LLaMA body (the reference leaves it trainable).
"""
from __future__ import annotations

import logging
import os
import pathlib
from typing import Dict, Tuple

import torch
import torch.nn as nn

from .common_arch import AttnPooler, LayerNorm, LayerNormFp32
from .rgb_vision_modal import VisionModal
from .text_modal import TextModal

logger = logging.getLogger("train")

MODAL_MAPPING = {"rgb": VisionModal, "text": TextModal}


class UniBind(nn.Module):
    def __init__(self, activate_modal: Tuple[str, str], config):
        assert len(activate_modal) > 0, "activate_modal should not be empty"
        for name in activate_modal:
            assert name in MODAL_MAPPING.keys(), f"Modal {name} is not supported"
        super().__init__()
        self.modal = activate_modal
        self.stage = config.stage
        if config.adjust_norm:
            norm_layer = LayerNormFp32 if config.dtype in ("float16", "bfloat16") else LayerNorm
        else:
            norm_layer = LayerNorm
        for modal in activate_modal:
            self.add_module(modal, MODAL_MAPPING[modal](config))
            if modal == "rgb":
                self.rgb_pooler = AttnPooler(
                    num_query=config.rgb_vision.attn_pooler.num_query,
                    num_layers=config.rgb_vision.attn_pooler.num_layers,
                    num_attention_heads=config.rgb_vision.attn_pooler.num_attn_heads,
                    encoder_hidden_size=VisionModal.EMBEDDING_DIM[config.rgb_vision.arch],
                    hidden_size=VisionModal.EMBEDDING_DIM[config.rgb_vision.arch],
                    output_size=config.text.hidden_size,
                    norm_layer=norm_layer,
                    checkpoint=getattr(config, "use_checkpoint", False),
                )

    # ------------------------------------------------------------------ dtype surface
    def to(self, *args, **kwargs):
        """``model.to(type_dict[config.dtype])`` is how cli_qa.py:89-90 and the eval drivers cast the model, and every shipped
        yaml says ``dtype: float16``: a float16 request is served in bfloat16 (``runtime.resolve_compute_dtype``; logged once)."""
        from . import runtime
        device, dtype, non_blocking, fmt = torch._C._nn._parse_to(*args, **kwargs)
        if dtype == torch.float16:
            dtype = runtime.resolve_compute_dtype(dtype)
            if device is not None:
                super().to(device=device, non_blocking=non_blocking)
            return super().to(dtype=dtype)
        return super().to(*args, **kwargs)

    def half(self):
        return self.to(torch.float16)

    # ------------------------------------------------------------------ checkpoints (UniBind.py:59-117)
    def load_rgb_encoder(self, path: str):
        assert hasattr(self, "rgb"), "rgb modal is not activated"
        ckpt = torch.load(path, map_location="cpu")
        if "model" in ckpt:
            ckpt = ckpt["model"]
        msg = self.rgb.encoder.load_state_dict(ckpt, strict=False)
        logger.info(f"Loading RGB Model: {msg}")

    def custom_save_checkpoint(self, file_name: str):
        """Returns ``{rgb_ckpt, other_ckpt}`` in the reference's FINAL.pt layout (UniBind.py:68-81).  Given a DeepSpeed ZeRO
        checkpoint directory it folds that to fp32 like the reference; given a file name it reads the live parameters (this
        package's trainer keeps no ZeRO shards on disk)."""
        if os.path.isdir(str(file_name)) and os.path.isfile(os.path.join(str(file_name), "latest")):
            # the reference's own call: a DeepSpeed ZeRO checkpoint directory is folded to fp32 (UniBind.py:68-70, 275-302)
            from .zero_checkpoint import get_fp32_state_dict_from_zero_checkpoint
            fp32 = get_fp32_state_dict_from_zero_checkpoint(str(file_name))
            rgb_ckpt = {k[len("rgb."):]: v.cpu() for k, v in fp32.items() if k.startswith("rgb.")}
            other_ckpt = dict(rgb_pooler={}, text_proj={}, embed_tokens={}, lm_head={})
            for k, v in fp32.items():
                for name in ("rgb_pooler", "embed_tokens"):
                    if name in k:
                        other_ckpt[name][k.split(name + ".")[-1]] = v
        else:
            rgb_ckpt = {k: v.detach().float().cpu() for k, v in self.rgb.state_dict().items()}
            other_ckpt = {"rgb_pooler": {k: v.detach().float().cpu() for k, v in self.rgb_pooler.state_dict().items()}}
        if self.stage >= 2 and self.text.text_encoder.has_lora():
            file_name = pathlib.Path(file_name)
            lora_dir = (file_name.parent if file_name.suffix else file_name) / "TextLoRA"
            self.text.text_encoder.save_pretrained(str(lora_dir))
        return dict(rgb_ckpt=rgb_ckpt, other_ckpt=other_ckpt)

    def custom_load_state_dict(self, state_dict_path, strict=False):
        if os.path.isdir(state_dict_path):
            # a DeepSpeed ZeRO checkpoint directory (UniBind.py:84-88): merged without DeepSpeed (zero_checkpoint.py), adapters
            # folded into the base weights as the reference does for a PeftModel
            from .zero_checkpoint import load_state_dict_from_zero_checkpoint
            load_state_dict_from_zero_checkpoint(self, state_dict_path)
            if self.text.text_encoder.has_lora():
                self.text.text_encoder = self.text.text_encoder.merge_and_unload()
            return None
        ckpt = torch.load(state_dict_path, map_location="cpu")
        if "model" in ckpt.keys():
            ckpt = ckpt["model"]
        text_path = pathlib.Path(state_dict_path).parent / "TextLoRA"
        msg = self.rgb.load_state_dict(ckpt["rgb_ckpt"], strict=strict)
        logger.info(f"After loading RGB encoder: Missing: {msg.missing_keys}. Unexpected: {msg.unexpected_keys}")
        self.rgb_pooler.load_state_dict(ckpt["other_ckpt"]["rgb_pooler"])
        del ckpt
        if text_path.exists():
            self.text.text_encoder.load_adapter(str(text_path), is_trainable=self.stage > 2)
            if self.stage == 0:  # Eval
                self.text.text_encoder = self.text.text_encoder.merge_and_unload()
        return None

    # ------------------------------------------------------------------ trainable set (UniBind.py:119-176)
    def prepare_for_training(self, freeze_vision: bool = False, freeze_text: bool = False, tune_rgb_pooler: bool = False,
                             model_path: str = False, tune_im_start: bool = False,
                             compute_dtype: torch.dtype = torch.float32):
        from . import runtime
        compute_dtype = runtime.resolve_compute_dtype(compute_dtype)   # float16 (every shipped yaml) -> bfloat16, logged once
        self.train()
        for param in self.rgb.parameters():
            param.requires_grad = not freeze_vision
            param.data = param.data.to(dtype=compute_dtype)
        for name, buffer in self.rgb.named_buffers():
            if "index" not in name and "id" not in name:
                buffer.data = buffer.data.to(dtype=compute_dtype)

        te = self.text.get_text_encoder()
        if freeze_text:
            self.text.eval()
            # Deviation (documented): the reference only freezes the embeddings here (UniBind.py:140-146) and leaves the
            # 6.48 B body parameters trainable; BASELINE.json defines stage 1 as "LLaMA-7B frozen, projector-only grads".
            for n, p in te.named_parameters():
                if "lora_" not in n:
                    p.requires_grad = False
        for p in te.get_input_embeddings().parameters():
            p.requires_grad = False
        for p in te.get_output_embeddings().parameters():
            p.requires_grad = False

        if hasattr(self, "rgb_pooler"):
            for param in self.rgb_pooler.parameters():
                param.requires_grad = bool(tune_rgb_pooler)
                param.data = param.data.to(dtype=compute_dtype)

        if tune_im_start:
            raise NotImplementedError("tune_im_start=True is not used by any shipped yaml")
        if model_path is not None and model_path is not False:
            msg = self.custom_load_state_dict(model_path)
            logger.info(f"After loading ckpt {model_path}: {msg}")

    # ------------------------------------------------------------------ forward (UniBind.py:178-199)
    def forward(self, data: Dict):
        out = dict()
        total_loss = 0.0
        image_embedding = self.rgb(data)
        for modal in self.modal:
            if modal == "rgb":
                image_embedding = self.rgb_pooler(image_embedding)
                continue
            output = getattr(self, modal)(data, image_embedding=image_embedding)
            if modal == "text":
                text_loss = output
                total_loss += text_loss
                out.update({"text_loss": text_loss})
        out.update({"total_loss": total_loss})
        return out

    def encode_image(self, image, pool):
        assert hasattr(self, "rgb"), "rgb modal is not activated"
        image_embedding = self.rgb.encode(image)
        if hasattr(self, "rgb_pooler"):
            image_embedding = self.rgb_pooler(image_embedding)
        if pool:
            return image_embedding.mean(dim=1)
        return image_embedding

    @torch.no_grad()   # HF generate never tracks gradients; callers need not wrap (cli_qa.py:176 uses inference_mode anyway)
    def generate(self, input_ids: torch.Tensor, images: torch.Tensor = None, do_sample: bool = True,
                 temperature: float = 0.2, max_new_tokens: int = 1024, streamer=None, use_cache: bool = True,
                 stopping_criteria=None, **kwargs):
        assert hasattr(self, "text"), "text modal is not activate"
        if (images is not None and hasattr(self, "rgb_pooler") and input_ids is not None and input_ids.shape[0] == 1 and
                input_ids.shape[1] > 1 and os.environ.get("LHRS_SCATTER_SPLICE", "1") != "0"):
            # generation fast path: the splice plan is made first (text rows + row map), and the AttnPooler's out_proj epilogue
            # writes its 144 rows per image straight into inputs_embeds (scatter epilogue of lhrs_gemm_bf16; common_arch.py:167-173
            # followed by text_modal.py:340-407 without the dense intermediate)
            feats = self.rgb.encode(images)
            embeds, row_map = self.text.splice_for_scatter(input_ids, feats.shape[0], self.rgb_pooler.num_query)
            self.rgb_pooler(feats, scatter_into=embeds.view(-1, embeds.shape[-1]), row_map=row_map)
            return self.text.generate(input_ids=input_ids, image_embedding=None, inputs_embeds=embeds, do_sample=do_sample,
                                      temperature=temperature, max_new_tokens=max_new_tokens, streamer=streamer,
                                      use_cache=use_cache, stopping_criteria=stopping_criteria, **kwargs)
        image_embedding = self.encode_image(images, pool=False) if images is not None else None
        return self.text.generate(input_ids=input_ids, image_embedding=image_embedding, do_sample=do_sample,
                                  temperature=temperature, max_new_tokens=max_new_tokens, streamer=streamer,
                                  use_cache=use_cache, stopping_criteria=stopping_criteria, **kwargs)
