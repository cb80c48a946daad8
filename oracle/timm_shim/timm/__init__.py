"""ORACLE shim (test infrastructure): a stand-in for the un-vendored, un-pinned, un-installed
`timm` package, exposing exactly the three calls the reference scorer makes
(/root/reference/models/global_eval_models.py:315, :333, :334).  `create_model` ignores
`pretrained` (no network) and returns the seeded pure-PyTorch trunk from oracle/trunks.py.
Put this directory on sys.path ONLY inside oracle/reference_loader.py."""
from . import data  # noqa: F401

SEED = 0
CALIBRATE_BN = True


def create_model(name, pretrained=False, **kwargs):
    from oracle.trunks import build_trunk

    return build_trunk(name, seed=SEED, calibrate_bn=CALIBRATE_BN)
