for mbs in 0 56 88 112; do echo "## SEMDIFF_HEAD_L2_MB=$mbs"; SEMDIFF_HEAD_L2_MB=$mbs python tools/profile_ops.py --microbatch 256 2>&1 | awk '/===/{print} / 0 conv| 1 maxpool/{print}' ; done
