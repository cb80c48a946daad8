"""ORACLE shim (test infrastructure): stand-in for the un-installed `pytora` package that
/root/reference/models/local_eval_models.py imports at :5.  The oracle never asks for LoRA (lora_rank=None)."""


def apply_lora(model, lora_r=None, **kwargs):
    raise NotImplementedError("the oracle runs the reference's U-Net with lora_rank=None only")
