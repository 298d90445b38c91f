"""Generate the golden fixtures that pin ``oracle/`` against the REFERENCE ITSELF (run in the build container only).

The reference ships no tests or golden vectors.  This script imports the reference's own code from
/root/reference and records its outputs on seeded tiny inputs:

  pooler_ref.pt   lhrs/models/common_arch.py AttnPooler (imported by file path; it needs torch only)
  splice_ref.pt   lhrs/models/text_modal.py TextModal.prepare_inputs_for_multimodal, run under an import-stub finder
                  (the 15 missing third-party packages are faked; only `peft` among them carries hot-path arithmetic
                  and it is not touched by this function)
  unibind_ref.pt  lhrs/models/UniBind.py UniBind.forward (loss) assembled from tiny HF CLIP / LLaMA modules + the
                  reference's VisionModal / AttnPooler / TextModal classes, under the same stubs
  hf_ref.pt       installed transformers (5.5.0) CLIPVisionModel hidden-state taps, LlamaForCausalLM logits / loss and
                  stock greedy ``generate`` tokens on tiny configs — the third-party arithmetic the reference delegates to
                  (pinned version 4.36.1 is absent; same published formulae)

/root/reference does not exist on the GPU box, so nothing at test time imports it: tests read these .pt files.
Usage:  python tests/golden/make_golden.py
"""
import importlib.abc
import importlib.machinery
import importlib.util
import os
import sys
import types
from unittest import mock

import torch

REF = "/root/reference"
OUT = os.path.dirname(os.path.abspath(__file__))

STUBS = {"albumentations", "braceexpand", "deepspeed", "geopandas", "ml_collections", "mmcv", "mmengine", "peft",
         "pycocoevalcap", "pycocotools", "termcolor", "thop", "timm", "torchmetrics", "webdataset"}


class _StubModule(types.ModuleType):
    def __getattr__(self, name):
        if name.startswith("__"):
            raise AttributeError(name)
        if name[:1].isupper():
            cls = type(name, (), {})
            setattr(self, name, cls)
            return cls
        m = mock.MagicMock(name=f"{self.__name__}.{name}")
        setattr(self, name, m)
        return m


class _StubFinder(importlib.abc.MetaPathFinder, importlib.abc.Loader):
    def find_spec(self, fullname, path, target=None):
        if fullname.split(".")[0] in STUBS:
            return importlib.machinery.ModuleSpec(fullname, self, is_package=True)
        return None

    def create_module(self, spec):
        m = _StubModule(spec.name)
        m.__path__ = []
        return m

    def exec_module(self, module):
        pass


def import_reference_models():
    sys.meta_path.insert(0, _StubFinder())
    import transformers
    shim = types.ModuleType("transformers.models.llama.tokenization_llama_fast")
    shim.LlamaTokenizerFast = getattr(transformers, "LlamaTokenizerFast", getattr(transformers, "LlamaTokenizer", object))
    sys.modules["transformers.models.llama.tokenization_llama_fast"] = shim
    sys.path.insert(0, REF)
    import lhrs.models  # noqa: F401
    return sys.modules["lhrs.models"]


def golden_pooler():
    spec = importlib.util.spec_from_file_location("ref_common_arch", os.path.join(REF, "lhrs/models/common_arch.py"))
    mod = importlib.util.module_from_spec(spec)
    spec.loader.exec_module(mod)
    cases = []
    for seed, (dim, heads, layers, out, B) in enumerate([(64, 4, 2, 96, 2), (32, 2, 3, 48, 3)]):
        torch.manual_seed(seed)
        pool = mod.AttnPooler(num_query=144, num_layers=layers, num_attention_heads=heads, encoder_hidden_size=dim,
                              hidden_size=dim, output_size=out, norm_layer=mod.LayerNorm).eval()
        with torch.no_grad():
            for n, p in pool.named_parameters():
                if "ln_" in n and n.endswith("weight"):
                    p.add_(0.1 * torch.randn_like(p))
                elif n.endswith("bias"):
                    p.add_(0.05 * torch.randn_like(p))
            x = torch.randn(B, 768, dim)
            y = pool(x)
        cases.append(dict(dim=dim, heads=heads, layers=layers, out=out, state_dict={k: v.clone() for k, v in pool.state_dict().items()},
                          x=x, y=y))
    torch.save(cases, os.path.join(OUT, "pooler_ref.pt"))
    print("pooler_ref.pt", [tuple(c["y"].shape) for c in cases])


def _tiny_llama(vocab=96, hidden=32, layers=2, heads=4, inter=64):
    from transformers import LlamaConfig
    return LlamaConfig(vocab_size=vocab, hidden_size=hidden, intermediate_size=inter, num_hidden_layers=layers,
                       num_attention_heads=heads, num_key_value_heads=heads, max_position_embeddings=2048,
                       rms_norm_eps=1e-5, rope_theta=10000.0, attention_bias=False, tie_word_embeddings=False,
                       pad_token_id=0, bos_token_id=1, eos_token_id=2)


def _ref_text_modal(models, lm, num_query=144):
    from lhrs.models.text_modal import TextModal
    tm = TextModal.__new__(TextModal)
    torch.nn.Module.__init__(tm)
    tm.text_encoder = lm
    tm.tune_pooler, tm.tune_im_start, tm.tune_im_patch = True, False, False
    tm.num_query = num_query
    return tm


def golden_splice(models):
    from lhrs.models.text_modal import CustomLlamaForCausalLM
    torch.manual_seed(0)
    lm = CustomLlamaForCausalLM(_tiny_llama()).eval()
    tm = _ref_text_modal(models, lm)
    table = lm.model.embed_tokens.weight.detach().clone()
    g = torch.Generator().manual_seed(1)
    cases = []

    def run(ids, mask, labels, n_img):
        img = torch.randn(n_img, 144, 32, generator=g)
        with torch.no_grad():
            _, m, _, e, l = tm.prepare_inputs_for_multimodal(ids, mask, labels, None, img)
        cases.append(dict(input_ids=ids, attention_mask=mask, labels=labels, image=img, out_mask=m, out_embeds=e, out_labels=l))

    # equal-length, one image each
    ids = torch.randint(3, 96, (3, 12), generator=g); ids[:, 3] = -200
    run(ids, torch.ones(3, 12, dtype=torch.bool), ids.clone(), 3)
    # ragged: different image positions is still equal length; add a text-only sample and right padding
    ids = torch.randint(3, 96, (4, 14), generator=g); ids[0, 1] = -200; ids[1, 5] = -200; ids[3, 2] = -200
    mask = torch.ones(4, 14, dtype=torch.bool); mask[1, 10:] = False; mask[2, 12:] = False
    labels = ids.clone(); labels[:, :2] = -100; labels[1, 10:] = -100
    run(ids, mask, labels, 4)
    # two images in one sample
    ids = torch.randint(3, 96, (2, 10), generator=g); ids[0, 2] = -200; ids[0, 7] = -200; ids[1, 4] = -200
    run(ids, torch.ones(2, 10, dtype=torch.bool), ids.clone(), 3)
    # no labels / no mask (generation path)
    ids = torch.randint(3, 96, (1, 9), generator=g); ids[0, 5] = -200
    run(ids, None, None, 1)
    # int64 mask dtype, all text-only
    ids = torch.randint(3, 96, (2, 6), generator=g)
    run(ids, torch.ones(2, 6, dtype=torch.long), ids.clone(), 2)
    torch.save(dict(table=table, cases=cases), os.path.join(OUT, "splice_ref.pt"))
    print("splice_ref.pt", [tuple(c["out_embeds"].shape) for c in cases])


def golden_hf_and_unibind(models):
    from transformers import CLIPVisionConfig, CLIPVisionModel, LlamaForCausalLM
    from lhrs.models.text_modal import CustomLlamaForCausalLM
    # ---- third-party arithmetic: HF CLIP taps / LLaMA logits, loss, greedy
    torch.manual_seed(0)
    vc = CLIPVisionConfig(hidden_size=32, intermediate_size=64, num_hidden_layers=6, num_attention_heads=2, image_size=56,
                          patch_size=14, hidden_act="quick_gelu", layer_norm_eps=1e-5)
    clip = CLIPVisionModel(vc).eval()
    px = torch.randn(2, 3, 56, 56)
    with torch.no_grad():
        hs = clip(px, output_hidden_states=True, return_dict=True).hidden_states
    taps = [6 // 3 - 1, 6 // 3 * 2 - 1, 6 - 2]
    feats = torch.cat([hs[s][:, 1:, :] for s in taps], dim=1)
    lm = LlamaForCausalLM(_tiny_llama()).eval()
    emb = torch.randn(2, 20, 32) * 0.5
    mask = torch.ones(2, 20, dtype=torch.long); mask[1, 15:] = 0
    labels = torch.randint(0, 96, (2, 20)); labels[:, :5] = -100; labels[1, 15:] = -100
    with torch.no_grad():
        out = lm(inputs_embeds=emb, attention_mask=mask, labels=labels)
        gen = lm.generate(inputs_embeds=emb[:1, :11], do_sample=False, max_new_tokens=12, use_cache=True, eos_token_id=None,
                          pad_token_id=0)
    torch.save(dict(clip_sd={k: v.clone() for k, v in clip.state_dict().items()}, px=px, taps=taps, feats=feats,
                    llama_sd={k: v.clone() for k, v in lm.state_dict().items()}, emb=emb, mask=mask, labels=labels,
                    logits=out.logits, loss=out.loss, greedy=gen[0]), os.path.join(OUT, "hf_ref.pt"))
    print("hf_ref.pt feats", tuple(feats.shape), "loss", float(out.loss), "greedy", gen[0].tolist())

    # ---- the reference's own UniBind.forward on tiny modules
    import lhrs.models.common_arch as ca
    from lhrs.models.rgb_vision_modal import VisionModal
    from lhrs.models.UniBind import UniBind
    torch.manual_seed(1)
    vc2 = CLIPVisionConfig(hidden_size=32, intermediate_size=64, num_hidden_layers=6, num_attention_heads=2, image_size=224,
                           patch_size=14, hidden_act="quick_gelu", layer_norm_eps=1e-5)
    vm = VisionModal.__new__(VisionModal)
    torch.nn.Module.__init__(vm)
    vm.arch = "vit_large"
    vm.encoder = CLIPVisionModel(vc2)
    vm.extract_stage = [6 // 3 - 1, 6 // 3 * 2 - 1, 6 - 2]
    lm2 = CustomLlamaForCausalLM(_tiny_llama())
    tm = _ref_text_modal(models, lm2)
    ub = UniBind.__new__(UniBind)
    torch.nn.Module.__init__(ub)
    ub.modal, ub.stage = ("rgb", "text"), 1
    ub.add_module("rgb", vm)
    ub.rgb_pooler = ca.AttnPooler(144, 2, 2, 32, 32, 32, norm_layer=ca.LayerNorm)
    ub.add_module("text", tm)
    ub.eval()
    g = torch.Generator().manual_seed(2)
    B, T = 3, 20
    ids = torch.randint(3, 96, (B, T), generator=g); ids[0, 3] = -200; ids[1, 1] = -200   # sample 2 is text-only
    labels = ids.clone(); labels[:, :4] = -100
    mask = torch.ones(B, T, dtype=torch.bool); mask[1, 16:] = False; labels[1, 16:] = -100
    data = dict(rgb=torch.randn(B, 3, 224, 224, generator=g), input_ids=ids, labels=labels, attention_mask=mask)
    with torch.no_grad():
        out = ub(data)
        img = ub.encode_image(data["rgb"], pool=False)
    torch.save(dict(vit_sd={k: v.clone() for k, v in vm.encoder.state_dict().items()},
                    pooler_sd={k: v.clone() for k, v in ub.rgb_pooler.state_dict().items()},
                    llama_sd={k: v.clone() for k, v in lm2.state_dict().items()}, data=data,
                    total_loss=out["total_loss"], image_embedding=img), os.path.join(OUT, "unibind_ref.pt"))
    print("unibind_ref.pt loss", float(out["total_loss"]), "img", tuple(img.shape))


if __name__ == "__main__":
    golden_pooler()
    models = import_reference_models()
    golden_splice(models)
    golden_hf_and_unibind(models)
