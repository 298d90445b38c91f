"""Training step of the hot path (stage-1/2/3): forward + backward over pooler / LoRA, flat-gradient allreduce, AdamW.

Filled in once the backward kernels are in place; bench.py reads ``AVAILABLE``.
"""
AVAILABLE = False


class SftStepper:
    def __init__(self, model, world_size=1):
        raise NotImplementedError("the backward path is not built yet")
