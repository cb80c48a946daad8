#!/usr/bin/env python
"""Turn the long-format CSV of tools/gpu_ncu_light.sh into a per-launch markdown table + roofline_traffic.json."""
import csv, json, re, sys, collections
src, out_md, title = sys.argv[1], sys.argv[2], sys.argv[3]
rows = [r for r in csv.DictReader(l for l in open(src) if not l.startswith("=="))]
launches = collections.OrderedDict()
for r in rows:
    d = launches.setdefault(r["ID"], {"kernel": re.sub(r"\(.*", "", r["Kernel Name"]).replace("void ", "").replace("semdiff::", "")})
    v = float(r["Metric Value"].replace(",", ""))
    u = r["Metric Unit"]
    if "byte" in u:
        v *= {"byte": 1, "Kbyte": 1e3, "Mbyte": 1e6, "Gbyte": 1e9}[u]
    if u in ("ns", "nsecond"): v /= 1e3
    if u in ("ms", "msecond"): v *= 1e3
    d[r["Metric Name"]] = v
L = list(launches.values())
def g(d, k): return d.get(k, 0.0)
tot = sum(g(d, "gpu__time_duration.sum") for d in L)
with open(out_md, "w") as f:
    f.write(f"# {title}\n\nncu --metrics (light set) over every kernel of ONE forward of 256 pairs (512 images, 224x224, bf16), "
            "`tools/gpu_ncu_light.sh`.  Times under ncu are serialised / cold-cache: compare shares.\n\n")
    f.write("| # | kernel | grid | us | share % | tensor pipe % | DRAM % of ncu peak | L2 % | DRAM MB | regs |\n|---|---|---|---|---|---|---|---|---|---|\n")
    for i, d in enumerate(L):
        t = g(d, "gpu__time_duration.sum")
        f.write(f"| {i} | {d['kernel'][:60]} | {int(g(d,'launch__grid_size'))} | {t:.1f} | {100*t/tot:.1f} | "
                f"{g(d,'sm__pipe_tensor_cycles_active.avg.pct_of_peak_sustained_active'):.1f} | {g(d,'gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed'):.1f} | "
                f"{g(d,'lts__throughput.avg.pct_of_peak_sustained_elapsed'):.1f} | {(g(d,'dram__bytes_read.sum')+g(d,'dram__bytes_write.sum'))/1e6:.0f} | {int(g(d,'launch__registers_per_thread'))} |\n")
    conv = [d for d in L if "conv_tc" in d["kernel"] or "conv3x3_strip" in d["kernel"] or "conv_chain" in d["kernel"]]
    ct = sum(g(d, "gpu__time_duration.sum") for d in conv)
    tw = sum(g(d, "gpu__time_duration.sum") * g(d, "sm__pipe_tensor_cycles_active.avg.pct_of_peak_sustained_active") for d in conv) / ct
    cb = sum(g(d, "dram__bytes_read.sum") + g(d, "dram__bytes_write.sum") for d in conv)
    f.write(f"\nTotals: {len(L)} launches, {tot:.0f} us.  conv kernels (conv_tc + conv3x3_strip + conv_chain): {len(conv)} launches, {ct:.0f} us = {100*ct/tot:.1f} % of the step, "
            f"time-weighted tensor-pipe utilisation {tw:.1f} %, DRAM traffic {cb/1e9:.2f} GB.\n")
    dist = [d for d in L if "distance" in d["kernel"]]
    f.write(f"distance_kernel: {len(dist)} launches, {sum(g(d,'gpu__time_duration.sum') for d in dist):.0f} us, DRAM "
            f"{sum(g(d,'dram__bytes_read.sum')+g(d,'dram__bytes_write.sum') for d in dist)/1e9:.2f} GB (algorithmic 1.54 GB).\n")
if len(sys.argv) > 4:
    # the digest of the kernel sources that were profiled (written on the GPU box by tools/gpu_ncu_light.sh next to the CSV):
    # bench.py quotes the traffic figure only while csrc/ still hashes to it
    import os
    dig = os.path.splitext(src)[0] + ".digest"
    stamp = open(dig).read().strip() if os.path.isfile(dig) else None
    json.dump({"kernel_sources_sha256": stamp, "conv_tc_dram_bytes_per_launch": cb / len(conv), "conv_tc_launches": len(conv), "conv_tc_dram_bytes_per_step": cb,
               "conv_tc_share_of_step_under_ncu": ct / tot, "time_weighted_tensor_pipe_pct": tw, "source": out_md}, open(sys.argv[4], "w"), indent=1)
print(open(out_md).read()[-600:])
