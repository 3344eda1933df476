#!/bin/bash
# Counts the Blackwell-native SASS mnemonics per kernel of libmmt_b200.so (B200_PROFILING.md: UTC*MMA = tcgen05.mma,
# LDTM/STTM = tcgen05.ld/st, UTMALDG/UTMASTG = TMA) -> profiles/<tag>_sass_summary.txt
tag=${1:-r02}
out=profiles/${tag}_sass_summary.txt
cuobjdump -sass mmt_b200/libmmt_b200.so | awk '
  /Function :/ { fn=$3; next }
  { for (m in pat) if ($0 ~ pat[m]) c[fn, m]++ }
  BEGIN { pat["UTCHMMA"]="UTCHMMA"; pat["UTCQMMA"]="UTC[QI]MMA"; pat["LDTM"]="LDTM"; pat["STTM"]="STTM"; pat["UTMALDG"]="UTMALDG";
          pat["UTMASTG"]="UTMASTG"; pat["UTMAREDG"]="UTMAREDG"; pat["HMMA"]=" HMMA"; pat["SYNCS"]="SYNCS"; pat["ACQBULK"]="ACQBULK|UBLKCP" }
  END { for (k in c) { split(k, a, SUBSEP); print a[1], a[2], c[k] } }' | sort > /tmp/sass_counts.txt
{
  echo "# SASS mnemonic counts per kernel of mmt_b200/libmmt_b200.so ($(date -u +%Y-%m-%d), nvcc $(nvcc --version | grep release | sed 's/.*release //'))"
  echo "# kernel (demangled prefix) : mnemonic=count ..."
  python3 - <<'PY'
import collections, subprocess
c = collections.defaultdict(dict)
for line in open('/tmp/sass_counts.txt'):
  fn, m, n = line.split()
  c[fn][m] = int(n)
tot = collections.Counter()
for fn in sorted(c):
  d = c[fn]
  if not any(k in d for k in ("UTCHMMA", "UTCQMMA", "LDTM", "STTM", "UTMALDG", "UTMASTG")):
    continue
  name = subprocess.run(["c++filt", fn], capture_output=True, text=True).stdout.strip()
  name = name.replace("(anonymous namespace)::", "").replace("void ", "").split("(")[0]
  print("%-60s %s" % (name[:60], " ".join("%s=%d" % kv for kv in sorted(d.items()))))
  tot.update(d)
print("TOTAL over tensor-core / TMA kernels: " + " ".join("%s=%d" % kv for kv in sorted(tot.items())))
PY
} > $out
cat $out
