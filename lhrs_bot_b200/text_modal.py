"""TextModal — LLaMA-2 decoder behind the reference's interface, mirror of lhrs/models/text_modal.py:63-667.

``self.text_encoder`` carries the HF ``LlamaForCausalLM`` parameter tree (``model.embed_tokens``, ``model.layers.N.*``,
``model.norm``, ``lm_head``; with LoRA the peft names ``<proj>.base_layer`` / ``<proj>.lora_A.default`` /
``<proj>.lora_B.default``), but no HF / peft code runs on the hot path:

* ``prepare_inputs_for_multimodal`` (:296-526) -> ``lhrs_splice_scan`` + ``lhrs_splice_fill`` (one tiny D2H of the lengths)
* ``decode`` (:258-294)                        -> ``lhrs_llama_fwd`` + ``lhrs_lm_head`` + ``lhrs_ce_fwd``
* ``generate`` (:528-627)                      -> prefill into a paged KV cache + ``lhrs_llama_decode_step`` per token
"""
from __future__ import annotations

import ctypes as C
import math
import os
from types import SimpleNamespace
from typing import Dict, List, Optional, Tuple, Union

import torch
import torch.nn as nn

from . import _lib, ops, runtime
from ._lib import LhrsKvCache, LhrsLlamaWeights, check
from .base_modal import BaseModal
from .constants import (DEFAULT_IM_END_TOKEN, DEFAULT_IM_START_TOKEN, DEFAULT_IMAGE_PATCH_TOKEN, DEFAULT_IMAGE_TOKEN,
                        IGNORE_INDEX, IMAGE_TOKEN_INDEX)

type_dict = {"float32": torch.float32, "float16": torch.float16, "bfloat16": torch.bfloat16}

PROJ_NAMES = ("q_proj", "k_proj", "v_proj", "o_proj", "gate_proj", "up_proj", "down_proj")


# ---------------------------------------------------------------------------------------------- parameter containers
class LoraLinearParams(nn.Module):
    """peft ``lora.Linear`` parameter layout: base_layer + lora_A/lora_B['default'] (peft==0.7.1 naming)."""

    def __init__(self, base: nn.Linear, r: int, alpha: float, dropout: float):
        super().__init__()
        self.base_layer = base
        self.r, self.lora_alpha, self.scaling = r, alpha, alpha / r
        self.lora_dropout_p = dropout
        dev, dt = base.weight.device, base.weight.dtype
        self.lora_A = nn.ModuleDict({"default": nn.Linear(base.in_features, r, bias=False, device=dev, dtype=dt)})
        self.lora_B = nn.ModuleDict({"default": nn.Linear(r, base.out_features, bias=False, device=dev, dtype=dt)})
        nn.init.kaiming_uniform_(self.lora_A["default"].weight, a=math.sqrt(5))   # peft reset_lora_parameters
        nn.init.zeros_(self.lora_B["default"].weight)
        base.weight.requires_grad_(False)

    @property
    def weight(self):
        return self.base_layer.weight


class _LlamaAttentionParams(nn.Module):
    def __init__(self, d):
        super().__init__()
        self.q_proj = nn.Linear(d, d, bias=False)
        self.k_proj = nn.Linear(d, d, bias=False)
        self.v_proj = nn.Linear(d, d, bias=False)
        self.o_proj = nn.Linear(d, d, bias=False)


class _LlamaMLPParams(nn.Module):
    def __init__(self, d, f):
        super().__init__()
        self.gate_proj = nn.Linear(d, f, bias=False)
        self.up_proj = nn.Linear(d, f, bias=False)
        self.down_proj = nn.Linear(f, d, bias=False)


class _RMSNormParams(nn.Module):
    def __init__(self, d, eps):
        super().__init__()
        self.weight = nn.Parameter(torch.ones(d))
        self.variance_epsilon = eps


class _LlamaLayerParams(nn.Module):
    def __init__(self, d, f, eps):
        super().__init__()
        self.self_attn = _LlamaAttentionParams(d)
        self.mlp = _LlamaMLPParams(d, f)
        self.input_layernorm = _RMSNormParams(d, eps)
        self.post_attention_layernorm = _RMSNormParams(d, eps)


class _LlamaModelParams(nn.Module):
    def __init__(self, cfg):
        super().__init__()
        self.embed_tokens = nn.Embedding(cfg.vocab_size, cfg.hidden_size)
        self.layers = nn.ModuleList([_LlamaLayerParams(cfg.hidden_size, cfg.intermediate_size, cfg.rms_norm_eps)
                                     for _ in range(cfg.num_hidden_layers)])
        self.norm = _RMSNormParams(cfg.hidden_size, cfg.rms_norm_eps)


class CustomLlamaForCausalLM(nn.Module):
    """Parameter tree of the reference's ``CustomLlamaForCausalLM`` (text_modal.py:30-60)."""

    def __init__(self, cfg):
        super().__init__()
        self.config = cfg
        self.model = _LlamaModelParams(cfg)
        self.lm_head = nn.Linear(cfg.hidden_size, cfg.vocab_size, bias=False)
        std = getattr(cfg, "initializer_range", 0.02)
        for m in self.modules():
            if isinstance(m, (nn.Linear, nn.Embedding)):
                m.weight.data.normal_(mean=0.0, std=std)
        self.peft_config = None

    def get_input_embeddings(self):
        return self.model.embed_tokens

    def get_output_embeddings(self):
        return self.lm_head

    def gradient_checkpointing_enable(self):   # activations are stashed, not recomputed; kept for API parity
        pass

    def enable_input_require_grads(self):
        pass

    # ---- LoRA (peft get_peft_model on every nn.Linear except lm_head, text_modal.py:133-151, :658-667)
    def add_lora(self, r: int, alpha: float, dropout: float) -> None:
        for p in self.parameters():
            p.requires_grad_(False)
        for layer in self.model.layers:
            for holder, names in ((layer.self_attn, PROJ_NAMES[:4]), (layer.mlp, PROJ_NAMES[4:])):
                for n in names:
                    base = getattr(holder, n)
                    if not isinstance(base, LoraLinearParams):
                        setattr(holder, n, LoraLinearParams(base, r, alpha, dropout))
        self.peft_config = SimpleNamespace(r=r, lora_alpha=alpha, lora_dropout=dropout, target_modules=list(PROJ_NAMES))

    def has_lora(self) -> bool:
        return self.peft_config is not None

    def merge_and_unload(self):
        """W += (alpha/r) * B·A and drop the adapters (peft semantics; eval path UniBind.py:114-115)."""
        for layer in self.model.layers:
            for holder, names in ((layer.self_attn, PROJ_NAMES[:4]), (layer.mlp, PROJ_NAMES[4:])):
                for n in names:
                    m = getattr(holder, n)
                    if isinstance(m, LoraLinearParams):
                        a, b = m.lora_A["default"].weight, m.lora_B["default"].weight
                        delta = ops.gemm(b.contiguous(), a.t().contiguous(), alpha=m.scaling, out_f32=True)  # [out,in]
                        m.base_layer.weight.data = (m.base_layer.weight.data.float() + delta).to(m.base_layer.weight.dtype)
                        setattr(holder, n, m.base_layer)
        self.peft_config = None
        return self

    def save_pretrained(self, path: str, safe_serialization: bool = True) -> None:
        """Adapter-only save in peft==0.7.1's on-disk layout (trainer.py:296-300 and UniBind.py:79 call this on text_encoder):
        ``adapter_model.safetensors`` (peft's default; ``adapter_model.bin`` with ``safe_serialization=False``) holding
        ``base_model.model.<module path>.lora_{A,B}.weight`` and an ``adapter_config.json`` a stock peft can read back."""
        os.makedirs(path, exist_ok=True)
        sd = {("base_model.model." + k).replace(".default", ""): v.detach().cpu().contiguous()
              for k, v in self.state_dict().items() if "lora_" in k}
        if safe_serialization:
            from safetensors.torch import save_file
            save_file(sd, os.path.join(path, "adapter_model.safetensors"), metadata={"format": "pt"})
        else:
            torch.save(sd, os.path.join(path, "adapter_model.bin"))
        if self.peft_config is not None:
            import json
            with open(os.path.join(path, "adapter_config.json"), "w") as f:
                json.dump(dict(peft_type="LORA", task_type="CAUSAL_LM", base_model_name_or_path=None, inference_mode=True,
                               r=self.peft_config.r, lora_alpha=self.peft_config.lora_alpha,
                               lora_dropout=self.peft_config.lora_dropout, target_modules=list(PROJ_NAMES), bias="none",
                               fan_in_fan_out=False, init_lora_weights=True, modules_to_save=None, layers_to_transform=None,
                               layers_pattern=None, rank_pattern={}, alpha_pattern={}, revision=None), f, indent=2, sort_keys=True)

    def load_adapter(self, path: str, is_trainable: bool) -> None:
        """``PeftModel.from_pretrained(model, path, is_trainable=...)`` for a LoRA adapter (UniBind.py:107-112): reads
        ``adapter_config.json`` and ``adapter_model.safetensors`` (or the older ``adapter_model.bin``)."""
        import json
        with open(os.path.join(path, "adapter_config.json")) as f:
            ac = json.load(f)
        if ac.get("peft_type", "LORA") != "LORA":
            raise NotImplementedError(f"adapter type {ac.get('peft_type')!r}: the reference trains LoRA only (text_modal.py:133-151)")
        targets = ac.get("target_modules") or list(PROJ_NAMES)
        if set(targets) != set(PROJ_NAMES):
            raise NotImplementedError(f"adapter targets {sorted(targets)}: the reference wraps all seven projections (text_modal.py:658-667)")
        self.add_lora(ac["r"], ac["lora_alpha"], ac.get("lora_dropout", 0.0))
        st_path = os.path.join(path, "adapter_model.safetensors")
        if os.path.exists(st_path):
            from safetensors.torch import load_file
            sd = load_file(st_path)
        else:
            sd = torch.load(os.path.join(path, "adapter_model.bin"), map_location="cpu")
        own = self.state_dict()
        loaded = 0
        for k, v in sd.items():
            kk = k.replace("base_model.model.", "", 1).replace("lora_A.weight", "lora_A.default.weight").replace(
                "lora_B.weight", "lora_B.default.weight")
            if kk not in own:
                raise KeyError(f"adapter tensor {k!r} has no counterpart in the model ({kk!r})")
            own[kk].copy_(v)
            loaded += 1
        expect = sum(1 for k in own if "lora_" in k)
        if loaded != expect:
            raise KeyError(f"adapter holds {loaded} tensors, the model expects {expect}")
        for n, p in self.named_parameters():
            if "lora_" in n:
                p.requires_grad_(is_trainable)


def dropout_call_seed(base: int, call: int) -> int:
    """Seed of the `call`-th training forward (1-based) under base seed `base` — one mask set per forward, reused by its backward."""
    return (int(base) + 0x9E3779B97F4A7C15 * int(call)) & 0xFFFFFFFFFFFFFFFF


def find_all_linear_names(model) -> List[str]:
    """text_modal.py:658-667"""
    names = set()
    for name, module in model.named_modules():
        if isinstance(module, nn.Linear):
            names.add(name)
    names.discard("lm_head")
    return list(names)


def _llama_config(config) -> SimpleNamespace:
    t = config.text
    return SimpleNamespace(
        vocab_size=t.vocab_size, hidden_size=t.hidden_size, intermediate_size=t.intermediate_size,
        num_hidden_layers=t.num_hidden_layers, num_attention_heads=t.num_attention_heads,
        max_position_embeddings=t.max_position_embeddings, rms_norm_eps=float(t.rms_norm_eps),
        initializer_range=getattr(t, "initializer_range", 0.02), rope_theta=float(getattr(t, "rope_theta", 10000.0)),
        pad_token_id=getattr(t, "pad_token_id", 0), bos_token_id=getattr(t, "bos_token_id", 1),
        eos_token_id=getattr(t, "eos_token_id", 2))


def _load_llama(config, device, dtype) -> CustomLlamaForCausalLM:
    path = config.text.path
    if os.path.isdir(str(path)):
        from transformers import AutoConfig, AutoModelForCausalLM   # only to read a local HF checkpoint's tensors
        hc = AutoConfig.from_pretrained(path)
        cfg = SimpleNamespace(vocab_size=hc.vocab_size, hidden_size=hc.hidden_size, intermediate_size=hc.intermediate_size,
                              num_hidden_layers=hc.num_hidden_layers, num_attention_heads=hc.num_attention_heads,
                              max_position_embeddings=hc.max_position_embeddings, rms_norm_eps=hc.rms_norm_eps,
                              initializer_range=hc.initializer_range, rope_theta=getattr(hc, "rope_theta", 10000.0),
                              pad_token_id=hc.pad_token_id, bos_token_id=hc.bos_token_id, eos_token_id=hc.eos_token_id)
        with torch.device("meta"):
            lm = CustomLlamaForCausalLM.__new__(CustomLlamaForCausalLM)
            nn.Module.__init__(lm)
            lm.config, lm.peft_config = cfg, None
            lm.model = _LlamaModelParams(cfg)
            lm.lm_head = nn.Linear(cfg.hidden_size, cfg.vocab_size, bias=False)
        lm = lm.to_empty(device=device).to(dtype)
        hf = AutoModelForCausalLM.from_pretrained(path, torch_dtype=dtype)
        lm.load_state_dict(hf.state_dict(), strict=False)
        return lm
    if not getattr(config, "random_init", False):
        raise FileNotFoundError(
            f"text.path={path!r} is not a local checkpoint directory and there is no network; "
            f"set random_init=True to build the architecture with seeded random weights")
    with torch.device(device):
        lm = CustomLlamaForCausalLM(_llama_config(config))
    return lm.to(dtype)


# ---------------------------------------------------------------------------------------------- the modal
class TextModal(BaseModal):
    def __init__(self, config):
        super().__init__(config)
        self.embedding_dim = config.text.hidden_size
        self.tune_pooler = config.tune_rgb_pooler
        self.tune_im_start = config.tune_im_start
        self.tune_im_patch = config.tune_im_patch
        self.num_query = config.rgb_vision.attn_pooler.num_query
        if self.tune_im_start or self.tune_im_patch:
            raise NotImplementedError("tune_im_start / tune_im_patch are False in every shipped yaml (text_modal.py:353-387 is dead code there)")
        if config.bits in [4, 8]:
            # Config/multi_modal_stage{2,3}.yaml:76 ship ``bits: 8``: the reference then loads the frozen LLaMA through
            # bitsandbytes LLM.int8() (text_modal.py:91-110) to fit 7B next to activations on a 24-80 GB GPU.  A B200 holds the
            # bf16 weights (13.5 GB of 180 GB) and runs them on the tensor cores directly: the quantised storage is not
            # reproduced — the frozen weights stay bf16 (strictly closer to the fp16 checkpoint than their int8 quantisation).
            runtime.warn_once("bits", f"lhrs_bot_b200: bits={config.bits} (bitsandbytes quantised storage of the frozen LLaMA) is "
                                      "not reproduced: the frozen weights are kept in bfloat16 in HBM - documented deviation, "
                                      "see DESIGN.md section 4")
        compute_dtype = runtime.resolve_compute_dtype(config.dtype)

        if getattr(config, "is_distribute", False):
            device = torch.device("cuda", getattr(config, "local_rank", 0))
        elif torch.cuda.is_available():
            device = torch.device("cuda", torch.cuda.current_device())
        else:
            device = torch.device("cpu")   # construction only (CPU tests of the host logic); forward needs CUDA
        self.text_encoder = _load_llama(config, device, compute_dtype)
        self.tokenizer = self.init_tokenizer(config.text.path)

        if config.lora.enable:
            self.text_encoder.add_lora(config.lora.lora_r, config.lora.lora_alpha, config.lora.lora_dropout)
        if getattr(config, "use_checkpoint", False):
            self.text_encoder.gradient_checkpointing_enable()
        self._table = None
        self._table_sig = None
        self._rope = None

    def get_text_encoder(self):
        return self.text_encoder

    def init_tokenizer(self, tokenizer_name: str):
        if not os.path.isdir(str(tokenizer_name)):
            return None   # no network: callers pass token ids (synthetic workloads) — text_modal.py:191-240 needs the files
        from transformers import AutoTokenizer
        tokenizer = AutoTokenizer.from_pretrained(tokenizer_name)
        tokenizer.pad_token_id = tokenizer.unk_token_id
        return tokenizer

    def get_modal_input(self, x: Dict[str, Union[str, torch.Tensor]]):
        return dict(input_ids=x["input_ids"], labels=x["labels"], attention_mask=x["attention_mask"])

    def encode(self, x: Dict) -> Dict:
        return x

    # ------------------------------------------------------------------ weight table
    def _rope_tables(self, device):
        if self._rope is None or self._rope[0].device != device:
            cfg = self.text_encoder.config
            hd = cfg.hidden_size // cfg.num_attention_heads
            inv = 1.0 / (cfg.rope_theta ** (torch.arange(0, hd, 2, dtype=torch.int64).float() / hd))
            freqs = torch.outer(torch.arange(cfg.max_position_embeddings, dtype=torch.float32), inv)
            # HF casts cos/sin to the activation dtype before use; keep the rounded values in fp32 tables
            cos = freqs.cos().to(torch.bfloat16).float().to(device).contiguous()
            sin = freqs.sin().to(torch.bfloat16).float().to(device).contiguous()
            self._rope = (cos, sin)
        return self._rope

    def weights(self) -> LhrsLlamaWeights:
        te = self.text_encoder
        params = list(te.parameters())
        sig = runtime.signature(params)
        if getattr(self, "_bwd_copies", False):     # in-place writes to a frozen weight (load_state_dict) must refresh its K-major copy
            sig = sig + tuple(p._version for n, p in te.named_parameters() if "lora_" not in n and p.dim() == 2)
        if self._table is not None and sig == self._table_sig:
            return self._table[0]
        for p in params:
            runtime.require_bf16_cuda(p, "TextModal parameter")
        runtime.contiguous_params(te)
        cfg = te.config
        layers = list(te.model.layers)
        dev = params[0].device
        cos, sin = self._rope_tables(dev)
        w = LhrsLlamaWeights()
        w.num_layers, w.dim, w.ffn, w.heads = len(layers), cfg.hidden_size, cfg.intermediate_size, cfg.num_attention_heads
        w.vocab, w.max_pos, w.eps = te.lm_head.out_features, cfg.max_position_embeddings, cfg.rms_norm_eps
        keep = [cos, sin]

        def base_w(m):
            return m.base_layer.weight if isinstance(m, LoraLinearParams) else m.weight

        def arr(fn):
            a = runtime.PtrArray([fn(l) for l in layers])
            keep.append(a)
            return a.ptr()

        w.ln1_w = arr(lambda l: l.input_layernorm.weight)
        w.q_w, w.k_w = arr(lambda l: base_w(l.self_attn.q_proj)), arr(lambda l: base_w(l.self_attn.k_proj))
        w.v_w, w.o_w = arr(lambda l: base_w(l.self_attn.v_proj)), arr(lambda l: base_w(l.self_attn.o_proj))
        w.ln2_w = arr(lambda l: l.post_attention_layernorm.weight)
        w.gate_w, w.up_w = arr(lambda l: base_w(l.mlp.gate_proj)), arr(lambda l: base_w(l.mlp.up_proj))
        w.down_w = arr(lambda l: base_w(l.mlp.down_proj))
        w.norm_w, w.lm_head = te.model.norm.weight.data_ptr(), te.lm_head.weight.data_ptr()
        w.embed = te.model.embed_tokens.weight.data_ptr()
        w.rope_cos, w.rope_sin = cos.data_ptr(), sin.data_ptr()
        if te.has_lora():
            a_list, b_list = [], []
            for l in layers:
                for holder, names in ((l.self_attn, PROJ_NAMES[:4]), (l.mlp, PROJ_NAMES[4:])):
                    for n in names:
                        m = getattr(holder, n)
                        a_list.append(m.lora_A["default"].weight)
                        b_list.append(m.lora_B["default"].weight)
            pa, pb = runtime.PtrArray(a_list), runtime.PtrArray(b_list)
            keep += [pa, pb]
            w.lora_r, w.lora_scale = te.peft_config.r, te.peft_config.lora_alpha / te.peft_config.r
            w.lora_a, w.lora_b = pa.ptr(), pb.ptr()
        else:
            w.lora_r, w.lora_scale = 0, 0.0
        w.lora_dropout, w.lora_seed = 0.0, 0          # set per training call by autograd.LlamaLossFunction
        if getattr(self, "_bwd_copies", False):
            # K-major (transposed) copies of the frozen projection weights for the dX GEMMs of the backward pass: +12.95 GB for
            # LLaMA-2-7B on a 180 GB part buys the faster form of the tcgen05 GEMM (LhrsLlamaWeights::*_wt, csrc/models_bwd.cu).
            # Rebuilt with the table whenever a base weight is re-pointed (checkpoint load, cast).
            def packed_t(l, names, holder):
                ws = [base_w(getattr(getattr(l, holder), n)).detach() for n in names]
                return torch.cat(ws, 0).t().contiguous() if len(ws) > 1 else ws[0].t().contiguous()
            qkv_t = [packed_t(l, ("q_proj", "k_proj", "v_proj"), "self_attn") for l in layers]
            o_t = [packed_t(l, ("o_proj",), "self_attn") for l in layers]
            gu_t = [packed_t(l, ("gate_proj", "up_proj"), "mlp") for l in layers]
            down_t = [packed_t(l, ("down_proj",), "mlp") for l in layers]
            lm_t = te.lm_head.weight.detach().t().contiguous()
            arrs = [runtime.PtrArray(x) for x in (qkv_t, o_t, gu_t, down_t)]
            keep += arrs + [lm_t]
            w.qkv_wt, w.o_wt, w.gu_wt, w.down_wt = (a.ptr() for a in arrs)
            w.lm_head_wt = lm_t.data_ptr()
        self._table, self._table_sig = (w, keep), sig
        return w

    def enable_backward_copies(self, on: bool = True) -> None:
        """Keep K-major copies of the frozen LLaMA projections for the backward dX GEMMs (training with a frozen body only:
        the copies would go stale if the base weights were updated).  Called by training.SftStepper."""
        if bool(on) != bool(getattr(self, "_bwd_copies", False)):
            self._bwd_copies = bool(on)
            self._table = self._table_sig = None

    # ------------------------------------------------------------------ LoRA dropout (peft lora.Linear input dropout)
    def lora_dropout_p(self) -> float:
        """Dropout probability of the LoRA branch for the NEXT training forward: peft applies nn.Dropout(lora_dropout) to the
        branch input while the module is in train mode (the trainer calls model.train(), IterBasedTrainer.py:87)."""
        pc = self.text_encoder.peft_config
        if pc is None or not self.training:
            return 0.0
        return float(getattr(pc, "lora_dropout", 0.0) or 0.0)

    def set_lora_dropout_seed(self, seed: int) -> None:
        """Base seed of the dropout masks (default: drawn from torch's generator at first use, so it follows
        torch.manual_seed and differs per rank under the reference's seed + rank rule)."""
        self._drop_base, self._drop_calls = int(seed) & ((1 << 62) - 1), 0

    def next_lora_dropout_seed(self) -> int:
        if getattr(self, "_drop_base", None) is None:
            self.set_lora_dropout_seed(int(torch.randint(0, 2 ** 62, (1,)).item()))
        self._drop_calls += 1
        return dropout_call_seed(self._drop_base, self._drop_calls)

    def lora_pairs(self):
        """[(lora_A.weight, lora_B.weight)] in weight-table order [layer*7 + proj]; empty without LoRA."""
        te = self.text_encoder
        if not te.has_lora():
            return []
        pairs = []
        for l in te.model.layers:
            for holder, names in ((l.self_attn, PROJ_NAMES[:4]), (l.mlp, PROJ_NAMES[4:])):
                for n in names:
                    m = getattr(holder, n)
                    pairs.append((m.lora_A["default"].weight, m.lora_B["default"].weight))
        return pairs

    # ------------------------------------------------------------------ splice
    def prepare_inputs_for_multimodal(self, input_ids: torch.Tensor, attention_mask: Optional[torch.Tensor],
                                      labels: Optional[torch.Tensor], past_key_values=None,
                                      image_embedding: Optional[torch.Tensor] = None):
        """text_modal.py:296-526.  Returns (input_ids|None, attention_mask, past_key_values, inputs_embeds, labels)."""
        if image_embedding is None or input_ids.shape[1] == 1:
            if past_key_values is not None and image_embedding is not None and input_ids.shape[1] == 1:
                attention_mask = torch.ones((attention_mask.shape[0], past_key_values.seq_len + 1),
                                            dtype=attention_mask.dtype, device=attention_mask.device)
            return input_ids, attention_mask, past_key_values, None, labels
        embeds, new_labels, new_mask, _, _ = self._splice(input_ids, attention_mask, labels, image_embedding)
        return None, new_mask, past_key_values, embeds, new_labels

    def splice_for_scatter(self, input_ids, n_images: int, num_query: int):
        """Text half of the splice for the generation path: inputs_embeds with the text rows filled and a row map telling where
        row (slot, i) of the pooler output belongs — ``AttnPooler.forward(scatter_into=embeds, row_map=...)`` then writes the image
        rows straight from its out_proj epilogue (no dense (B, 144, 4096) tensor, no splice copy of it)."""
        embeds, _, _, row_map, _ = self._splice(input_ids, None, None, None, fill_image=False, nq=num_query, n_slots=n_images)
        return embeds, row_map

    def _splice(self, input_ids, attention_mask, labels, image_embedding, fill_image: bool = True, nq: int = 0, n_slots: int = 0):
        table = self.text_encoder.model.embed_tokens.weight
        runtime.require_bf16_cuda(table, "embed_tokens.weight")
        B, T = input_ids.shape
        if image_embedding is not None:
            runtime.require_bf16_cuda(image_embedding, "image_embedding")
            nq = image_embedding.shape[1]
            n_slots = image_embedding.shape[0]
        info = ops.splice_scan(input_ids, nq)
        head = info[: (B + 1) * 4].cpu().view(B + 1, 4)      # the one D2H sync of the splice (B+1 small rows)
        total_slots, s_out = int(head[B, 0]), int(head[B, 1])
        if total_slots > n_slots:
            raise IndexError(f"prepare_inputs_for_multimodal: the batch consumes {total_slots} image slots but "
                             f"image_embedding holds {n_slots} (text_modal.py:343 would raise IndexError)")
        if attention_mask is not None and labels is None and int(head[B, 2]) != s_out:
            raise NameError("ragged multimodal batch with attention_mask but no labels: the reference reads `_new_labels` "
                            "before assignment here (text_modal.py:474)")
        want_bool = attention_mask is not None and attention_mask.dtype == torch.bool
        embeds, new_labels, new_mask, row_map = ops.splice_fill(
            input_ids, labels, attention_mask, info, table, image_embedding if fill_image else None, s_out, nq, n_slots)
        if new_mask is not None:
            new_mask = new_mask.bool() if want_bool else new_mask.to(attention_mask.dtype)
        return embeds, new_labels, new_mask, row_map, head

    # ------------------------------------------------------------------ training / scoring forward
    def decode(self, input_ids: torch.Tensor, image_embedding: torch.Tensor = None,
               attention_mask: Optional[torch.Tensor] = None, labels: torch.Tensor = None):
        """text_modal.py:258-294 -> scalar loss (fp32)."""
        if image_embedding is not None and image_embedding.requires_grad and torch.is_grad_enabled() or \
                (torch.is_grad_enabled() and any(p.requires_grad for p in self.text_encoder.parameters())):
            from .autograd import text_loss_with_grad
            return text_loss_with_grad(self, input_ids, image_embedding, attention_mask, labels)
        _, mask, _, embeds, new_labels = self.prepare_inputs_for_multimodal(input_ids, attention_mask, labels, None, image_embedding)
        if embeds is None:   # text-only call without image features
            embeds = self.embed(input_ids)
            new_labels = labels
        from .autograd import ragged_plan, supervised_rows
        sel = supervised_rows(new_labels)
        # scoring a right-padded batch: the decoder stack runs on the real rows only (same loss; autograd.ragged_plan)
        plan = ragged_plan(mask) if sel is not None else None
        if plan is not None:
            from . import autograd as _ag
            _ag.RAGGED_CALLS += 1
            B, S, D = embeds.shape
            hidden = self.llama_forward_ragged(embeds.view(B * S, D).index_select(0, plan.real_rows), B, plan.s_max, plan.seq_off,
                                               plan.positions)
        else:
            hidden = self.llama_forward(embeds, mask)
        if sel is not None:      # lm_head + CE over the supervised rows only (identical loss)
            rows, ce_labels = sel
            if plan is not None:
                rows = plan.packed_of.index_select(0, rows)
            hc = torch.zeros((rows.numel() + 1, hidden.shape[-1]), device=hidden.device, dtype=hidden.dtype)
            hc[:-1] = hidden.reshape(-1, hidden.shape[-1]).index_select(0, rows)
            logits = self.lm_head(hc).unsqueeze(0)
        else:
            logits, ce_labels = self.lm_head(hidden), new_labels
        loss_sum, count, _ = ops.ce_fwd(logits, ce_labels)
        return loss_sum[0] / count[0].to(torch.float32)

    def embed(self, input_ids: torch.Tensor) -> torch.Tensor:
        table = self.text_encoder.model.embed_tokens.weight
        B, T = input_ids.shape
        info = ops.splice_scan(input_ids, 1)
        embeds, _, _, _ = ops.splice_fill(input_ids, None, None, info, table, None, T, 1, 0, want_row_map=False)
        return embeds

    def llama_forward(self, inputs_embeds: torch.Tensor, attention_mask: Optional[torch.Tensor], stash=None, kv=None):
        lib = _lib.load()
        w = self.weights()
        B, S, D = inputs_embeds.shape
        runtime.require_bf16_cuda(inputs_embeds, "inputs_embeds")
        hidden = torch.empty_like(inputs_embeds)
        km = None
        if attention_mask is not None:
            km = attention_mask.to(torch.uint8).contiguous()
        ws_bytes = lib.lhrs_llama_workspace_bytes(C.byref(w), B, S)
        ws = runtime.workspace(ws_bytes, inputs_embeds.device)
        check(lib.lhrs_llama_fwd(C.byref(w), inputs_embeds.contiguous().data_ptr(), B, S, None if km is None else km.data_ptr(),
                                 hidden.data_ptr(), None if stash is None else stash.data_ptr(),
                                 None if kv is None else C.byref(kv), ws.data_ptr(), ws.numel(), runtime.stream()),
              "lhrs_llama_fwd")
        return hidden

    def llama_forward_ragged(self, embeds_rows: torch.Tensor, B: int, s_max: int, seq_off: torch.Tensor, positions: torch.Tensor,
                             stash=None) -> torch.Tensor:
        """The decoder stack over a padding-free batch: ``embeds_rows`` [rows, dim] holds the B sequences back to back
        (lhrs_llama_fwd_ragged).  Same results on every real position as ``llama_forward`` on the right-padded batch."""
        lib = _lib.load()
        w = self.weights()
        runtime.require_bf16_cuda(embeds_rows, "inputs_embeds")
        rows = embeds_rows.shape[0]
        hidden = torch.empty_like(embeds_rows)
        ws_bytes = lib.lhrs_llama_workspace_bytes(C.byref(w), B, s_max)
        ws = runtime.workspace(ws_bytes, embeds_rows.device)
        check(lib.lhrs_llama_fwd_ragged(C.byref(w), embeds_rows.contiguous().data_ptr(), B, s_max, rows, seq_off.data_ptr(),
                                        positions.data_ptr(), hidden.data_ptr(), None if stash is None else stash.data_ptr(),
                                        ws.data_ptr(), ws.numel(), runtime.stream()), "lhrs_llama_fwd_ragged")
        return hidden

    def lm_head(self, hidden: torch.Tensor) -> torch.Tensor:
        lib = _lib.load()
        w = self.weights()
        rows = hidden.numel() // hidden.shape[-1]
        logits = torch.empty(tuple(hidden.shape[:-1]) + (w.vocab,), device=hidden.device, dtype=torch.bfloat16)
        check(lib.lhrs_lm_head(C.byref(w), hidden.contiguous().data_ptr(), rows, logits.data_ptr(), runtime.stream()), "lhrs_lm_head")
        return logits

    def logits(self, input_ids, image_embedding, attention_mask=None):
        """Full-sequence logits (B, S, V) bf16 — used by the parity tests."""
        _, mask, _, embeds, _ = self.prepare_inputs_for_multimodal(input_ids, attention_mask, None, None, image_embedding)
        if embeds is None:
            embeds = self.embed(input_ids)
        return self.lm_head(self.llama_forward(embeds, mask))

    # ------------------------------------------------------------------ generation
    def generate(self, image_embedding: torch.Tensor = None, prompt=None, input_ids: Optional[torch.LongTensor] = None,
                 do_sample: bool = True, temperature: float = 0.2, max_new_tokens: int = 1024, streamer=None,
                 use_cache: bool = True, stopping_criteria=None, attention_mask=None, **kwargs):
        """``inputs_embeds=`` (extension, used by UniBind.generate): the already spliced prompt (see ``splice_for_scatter``)."""
        from .generation import generate as _generate
        if input_ids is None:
            raise NotImplementedError("caption-prompt generation (text_modal.py:543-579) needs the tokenizer files; pass input_ids")
        return _generate(self, input_ids, image_embedding, do_sample=do_sample, temperature=temperature,
                         max_new_tokens=max_new_tokens, streamer=streamer, stopping_criteria=stopping_criteria,
                         attention_mask=attention_mask, **kwargs)


def tokenizer_image_token(prompt, tokenizer, image_token_index=IMAGE_TOKEN_INDEX, return_tensors=None):
    """text_modal.py:630-655 — host-side string/token utility, restated with the same behaviour."""
    chunks = [tokenizer(chunk).input_ids for chunk in prompt.split("<image>")]
    input_ids: List[int] = []
    offset = 0
    if len(chunks) > 0 and len(chunks[0]) > 0 and chunks[0][0] == tokenizer.bos_token_id:
        offset = 1
        input_ids.append(chunks[0][0])
    sep = [image_token_index] * (offset + 1)
    for i, chunk in enumerate(chunks):
        if i > 0:
            input_ids.extend(sep[offset:])
        input_ids.extend(chunk[offset:])
    if return_tensors is not None:
        if return_tensors == "pt":
            return torch.tensor(input_ids, dtype=torch.long)
        raise ValueError(f"Unsupported tensor type: {return_tensors}")
    return input_ids
