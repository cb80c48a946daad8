import os, sys
sys.path.insert(0, "/root/repo"); sys.path.insert(0, "/root/repo/tools")
import importlib.util
spec = importlib.util.spec_from_file_location("cm", "tools/conv_microbench.py")
src = open("tools/conv_microbench.py").read().split("for shape in")[0]
exec(src)
for shape in [(16, 56, 56, 64, 64, 3, 1, 1), (16, 56, 56, 576, 64, 1, 1, 0), (32, 28, 28, 128, 128, 3, 1, 1), (32, 28, 28, 1152, 128, 1, 1, 0),
              (64, 14, 14, 256, 256, 3, 1, 1), (64, 14, 14, 2304, 256, 1, 1, 0)]:
    ms, tf = run(*shape, _lib.CONV_TC_TMA, reps=20)
    print(f"{shape} {ms*1000:8.1f} us {tf:7.1f} TF/s", flush=True)
