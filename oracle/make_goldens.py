"""ORACLE (test infrastructure): generate tests/golden/scorer_goldens.json by running the
UNMODIFIED reference file (/root/reference/models/global_eval_models.py, via oracle/reference_loader)
on seeded synthetic inputs.  Run in the build container (the only place /root/reference exists):

    python -m oracle.make_goldens

Inputs and weights are regenerated from the seeds stored in each record (oracle/synth.py), so the
file stays small.  CPU float arithmetic may differ in the last bits between hosts (MKL-DNN picks
kernels by ISA, and BN calibration is a forward pass), hence consumers compare with rtol 5e-5.
"""
from __future__ import annotations

import json
import os
import sys

import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))

from oracle import reference_loader as rl  # noqa: E402
from oracle.synth import make_pairs, set_head  # noqa: E402

OUT = os.path.join(os.path.dirname(os.path.dirname(os.path.abspath(__file__))), "tests", "golden", "scorer_goldens.json")

CASES = [
    # trunk, depth, n_pairs, input seed, head mode
    ("resnet50", 3, 8, 0, "abs"),          # BASELINE.json configs[0]
    ("resnet50", 3, 8, 0, "signed"),
    ("resnet50", 2, 5, 1, "abs"),          # reference training batch size 5 (CLIPLPIPS...py:169)
    ("resnet50", 1, 5, 2, "abs"),
    ("resnet50", 0, 1, 3, "abs"),
    ("resnet50_clip.openai", 3, 8, 0, "abs"),
    ("resnet50_clip.openai", 3, 8, 0, "signed"),
    ("resnet50_clip.openai", 1, 5, 2, "abs"),
    ("resnet50_clip.openai", 0, 1, 3, "abs"),
]


def main():
    torch.set_num_threads(os.cpu_count())
    records = []
    for trunk, depth, n, seed, head in CASES:
        model = rl.build_reference_scorer(trunk, depth, seed=0)
        set_head(model, head)
        gt, sr = make_pairs(n, seed=seed)
        with torch.no_grad():
            scores = model(gt, sr)
            relu = model.final_relu
            model.final_relu = torch.nn.Identity()   # attribute swap on the instance; file untouched
            pre = model(gt, sr)
            model.final_relu = relu
            taps = list(model.outputs.values())      # taps of the last trunk pass (= sr)
        records.append({
            "trunk": trunk, "depth": depth, "n_pairs": n, "input_seed": seed, "weight_seed": 0, "head": head,
            "scores": [float(x) for x in scores], "pre_relu": [float(x) for x in pre],
            "sr_tap_abs_mean": [float(t.abs().mean()) for t in taps],
            "sr_tap_shapes": [list(t.shape) for t in taps],
            "state_dict_keys": len(model.state_dict()),
        })
        print(trunk, depth, n, head, records[-1]["scores"][:4], flush=True)
    meta = {"generator": "oracle/make_goldens.py", "torch": torch.__version__,
            "reference_file": rl.REFERENCE_FILE, "note": "outputs of the unmodified reference file through oracle/timm_shim"}
    with open(OUT, "w") as f:
        json.dump({"meta": meta, "records": records}, f, indent=1)
    print("wrote", OUT)


UNET_OUT = os.path.join(os.path.dirname(OUT), "unet_goldens.json")


def main_unet():
    """tests/golden/unet_goldens.json: the reference's local-map U-Nets (local_eval_models.py:7-339, executed verbatim via
    reference_loader.build_reference_unet) on seeded pairs, decoder calibrated by restated.calibrate_unet_decoder.  The
    224x224 maps are stored on a 16-pixel grid (196 values per pair)."""
    from oracle.restated import calibrate_unet_decoder

    torch.set_num_threads(os.cpu_count())
    records = []
    for trunk in ("resnet50", "resnet50_clip.openai"):
        model = calibrate_unet_decoder(rl.build_reference_unet(trunk, seed=0))
        gt, sr = make_pairs(2, seed=5)
        with torch.no_grad():
            m = model(gt, sr)
        records.append({"trunk": trunk, "n_pairs": 2, "input_seed": 5, "weight_seed": 0, "shape": list(m.shape),
                        "grid": 16, "map": [[float(v) for v in row] for row in m[:, 0, 8::16, 8::16].reshape(2, -1)],
                        "mean": float(m.mean()), "state_dict_keys": len(model.state_dict())})
        print(trunk, records[-1]["mean"], flush=True)
    meta = {"generator": "oracle/make_goldens.py unet", "torch": torch.__version__, "reference_file": rl.REFERENCE_LOCAL_FILE,
            "note": "outputs of the reference file's first two classes (lines 1-339, executed verbatim) through oracle/timm_shim"}
    with open(UNET_OUT, "w") as f:
        json.dump({"meta": meta, "records": records}, f, indent=1)
    print("wrote", UNET_OUT)


if __name__ == "__main__":
    main_unet() if "unet" in sys.argv[1:] else main()
