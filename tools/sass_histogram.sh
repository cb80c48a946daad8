#!/bin/bash
# Blackwell evidence: per-kernel counts of the tcgen05 / TMEM / TMA SASS mnemonics in the built library
# (UTCHMMA = tcgen05.mma, LDTM = tcgen05.ld, UTMALDG / UTMASTG = TMA load / store, UTCBAR = tcgen05.commit).
# usage: tools/sass_histogram.sh > profiles/r2_sass_histogram.txt     (needs cuobjdump, no GPU)
so=measuring-semantic-differences-in-the-super-resolution-domain_b200/libsemdiff_b200.so
echo "# cuobjdump -sass $so ($(date -u +%F), $(git rev-parse --short HEAD 2>/dev/null)): mnemonic counts per kernel"
printf "%-110s %8s %6s %8s %8s %7s %8s\n" kernel UTCHMMA LDTM UTMALDG UTMASTG UTCBAR SYNCS
cuobjdump -sass "$so" | awk '
  /Function :/ { if (name != "") out(); name = $3; for (k in c) delete c[k]; next }
  { for (m in want) if (index($0, m " ") || index($0, m ".")) c[m]++ }
  function out() { cmd = "c++filt " name; cmd | getline d; close(cmd); sub(/\(.*/, "", d); sub(/^void /, "", d);
                   printf "%-110s %8d %6d %8d %8d %7d %8d\n", substr(d, 1, 110), c["UTCHMMA"], c["LDTM"], c["UTMALDG"], c["UTMASTG"], c["UTCBAR"], c["SYNCS"] }
  BEGIN { want["UTCHMMA"]; want["LDTM"]; want["UTMALDG"]; want["UTMASTG"]; want["UTCBAR"]; want["SYNCS"] }
  END { if (name != "") out() }' | sort
