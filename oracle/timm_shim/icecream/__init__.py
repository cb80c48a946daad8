"""ORACLE shim: the reference does `from icecream import ic`
(/root/reference/models/global_eval_models.py:919); icecream is not installed."""


def ic(*args):
    return args[0] if len(args) == 1 else args
