"""oracle/ — TEST INFRASTRUCTURE ONLY.  A CPU (plain PyTorch fp32) restatement of the LHRS-Bot hot path.

Nothing in the product package (``lhrs_bot_b200``) may import this package.  Only ``tests/``,
``__graft_entry__.smoke()`` and the ``cpu_baseline`` / ``--impl reference`` legs of ``bench.py`` use it,
and only as the checker or the timed CPU baseline — never as the thing shipped.

What it restates (paths relative to the reference checkout, commit 84d9fddf):
  * ``pooler.py``   lhrs/models/common_arch.py:134-173 (AttnPooler.forward), :315-333 (ResidualAttentionBlock)
  * ``splice.py``   lhrs/models/text_modal.py:296-526 (TextModal.prepare_inputs_for_multimodal)
  * ``vit.py``      lhrs/models/rgb_vision_modal.py:159-184 (taps 7/15/22, CLS dropped) over HF CLIPVisionModel math
  * ``llama.py``    lhrs/models/text_modal.py:258-294 (decode -> HF LlamaForCausalLM loss), :36-60 (generation rule),
                    :133-151 + :658-667 (LoRA targets)
  * ``unibind.py``  lhrs/models/UniBind.py:178-242 (forward / encode_image / generate orchestration)

Pinning status.  The reference ships no tests, fixtures or golden vectors, so nothing of the
reference's own pins this path.  The oracle is instead pinned against outputs of the reference code itself,
generated in the build container by ``tests/golden/make_golden.py``:
  * pooler:  the reference's own ``common_arch.py`` imported by file path (it depends on torch only);
  * splice:  the reference's own ``TextModal.prepare_inputs_for_multimodal`` run under an import-stub finder;
  * ViT / LLaMA / LoRA arithmetic lives in un-vendored third-party packages (transformers==4.36.1, peft==0.7.1,
    pyproject.toml:16,23).  The installed transformers 5.5.0 implements the same published formulae and is used
    as the cross-check (``tests/test_oracle.py`` compares against it live — it is present on the GPU box too);
    peft is absent everywhere, so LoRA follows its published definition h = Wx + (alpha/r)·B(A·x): "parity unpinned"
    for LoRA beyond that definition.
"""
