"""Oracle (test infrastructure): embed lookup + image splice.

Restates ``TextModal.prepare_inputs_for_multimodal`` (lhrs/models/text_modal.py:296-526) for the shipped
configuration ``tune_im_start: False`` (Config/multi_modal_stage{1,2,3}.yaml) — the ``tune_im_start`` variant
(:353-387) is dead under every shipped yaml.  Written as explicit per-sample loops over integer positions so the
label / mask / length outputs can be compared bit-for-bit with the CUDA kernel.
"""
from __future__ import annotations

from typing import List, Optional, Tuple

import torch

IGNORE_INDEX = -100        # lhrs/models/__init__.py:1
IMAGE_TOKEN_INDEX = -200   # lhrs/models/__init__.py:2


def splice_plan(input_ids: torch.Tensor, num_query: int) -> Tuple[List[int], List[int], List[int]]:
    """Per sample: (#image tokens, spliced length, first image slot).  text_modal.py:320-339 (text-only samples
    still consume an image slot), :340-407 (each -200 is replaced by num_query rows)."""
    n_img, new_len, slot_base = [], [], []
    slot = 0
    T = input_ids.shape[1]
    for row in input_ids.tolist():
        n = sum(1 for t in row if t == IMAGE_TOKEN_INDEX)
        n_img.append(n)
        new_len.append(T + n * (num_query - 1))
        slot_base.append(slot)
        slot += n if n > 0 else 1
    return n_img, new_len, slot_base


def prepare_inputs_for_multimodal(
    input_ids: torch.Tensor,
    attention_mask: Optional[torch.Tensor],
    labels: Optional[torch.Tensor],
    embed_table: torch.Tensor,
    image_embedding: Optional[torch.Tensor],
):
    """Returns (attention_mask, inputs_embeds, labels) exactly as text_modal.py:296-526 does for
    ``past_key_values=None`` and a non-None image_embedding with more than one input position."""
    assert image_embedding is not None and input_ids.shape[1] != 1  # :304 fast path is a pass-through
    B, T = input_ids.shape
    nq, dim = image_embedding.shape[1], image_embedding.shape[2]
    new_embeds: List[torch.Tensor] = []
    new_labels: Optional[List[torch.Tensor]] = [] if labels is not None else None
    cur_image_idx = 0
    for b in range(B):
        ids = input_ids[b]
        img_pos = (ids == IMAGE_TOKEN_INDEX).nonzero().flatten().tolist()
        if len(img_pos) == 0:                                     # :321-339
            new_embeds.append(embed_table[ids])
            if labels is not None:
                new_labels.append(labels[b])
            cur_image_idx += 1
            continue
        pieces, lab_pieces = [], []
        start = 0
        for p in img_pos:                                         # :346-407
            pieces.append(embed_table[ids[start:p]])
            pieces.append(image_embedding[cur_image_idx].to(embed_table.dtype))
            if labels is not None:
                lab_pieces.append(labels[b, start:p])
                lab_pieces.append(torch.full((nq,), IGNORE_INDEX, dtype=labels.dtype, device=labels.device))
            cur_image_idx += 1
            start = p + 1
        if start < T:                                             # :409-423
            pieces.append(embed_table[ids[start:]])
            if labels is not None:
                lab_pieces.append(labels[b, start:])
        new_embeds.append(torch.cat(pieces, 0))
        if labels is not None:
            new_labels.append(torch.cat(lab_pieces, 0))

    lens = [e.shape[0] for e in new_embeds]
    max_len = max(lens)
    if any(l != lens[0] for l in lens):                           # ragged branch :440-505
        embeds = torch.zeros(B, max_len, dim, dtype=embed_table.dtype, device=embed_table.device)
        for b, e in enumerate(new_embeds):
            embeds[b, : e.shape[0]] = e
        out_labels = None
        if labels is not None:
            out_labels = torch.full((B, max_len), IGNORE_INDEX, dtype=labels.dtype, device=labels.device)
            for b, l in enumerate(new_labels):
                out_labels[b, : l.shape[0]] = l
        out_mask = None
        if attention_mask is not None:
            # :474-497 — note: requires labels in the reference (it sizes the left pad from the new labels)
            out_mask = torch.zeros(B, max_len, dtype=attention_mask.dtype, device=attention_mask.device)
            for b in range(B):
                grown = lens[b] - T
                out_mask[b, :grown] = True
                out_mask[b, grown: lens[b]] = attention_mask[b]
    else:                                                         # equal-length branch :506-524
        embeds = torch.stack(new_embeds, 0)
        out_labels = torch.stack(new_labels, 0) if labels is not None else None
        out_mask = None
        if attention_mask is not None:
            left = torch.full((B, max_len - T), True, dtype=attention_mask.dtype, device=attention_mask.device)
            out_mask = torch.cat((left, attention_mask), dim=1)
    return out_mask, embeds, out_labels
