"""Per-kernel SASS opcode histogram of libgonova_hift.so: what proves tcgen05 / TMEM / TMA in the shipped binary
(UTCHMMA = tcgen05.mma, UTMALDG / UTMASTG = TMA loads / stores, LDTM / STTM = tcgen05.ld / st, UTCBAR = tcgen05.commit,
SYNCS = mbarrier ops; HMMA would be the legacy mma.sync path).  CPU only: cuobjdump reads the built library.
usage: python tools/sass_histogram.py > profiles/r02_sass_opcodes.txt"""
import collections
import re
import subprocess
import sys

lib = sys.argv[1] if len(sys.argv) > 1 else "gonova_tts_b200/lib/libgonova_hift.so"
out = subprocess.run(["cuobjdump", "-sass", lib], capture_output=True, text=True, check=True).stdout
names = subprocess.run(["c++filt"], input="\n".join(re.findall(r"Function : (\S+)", out)), capture_output=True, text=True).stdout.split("\n")
fn, hist, order = None, collections.defaultdict(collections.Counter), []
it = iter(names)
for line in out.splitlines():
    m = re.match(r"\s+Function : (\S+)", line)
    if m:
        fn = next(it)
        fn = re.sub(r"\(.*", "", fn).replace("void gnv::", "").replace("__nv_bfloat16", "bf16")
        order.append(fn)
        continue
    m = re.match(r"\s+/\*[0-9a-f]{4,}\*/\s+(?:@!?U?P\d+\s+)?([A-Z][A-Z0-9_]*)", line)
    if m and fn:
        hist[fn][m.group(1)] += 1
KEY = ["UTCHMMA", "UTCQMMA", "UTCBAR", "UTMALDG", "UTMASTG", "LDTM", "STTM", "SYNCS", "ELECT", "MUFU", "HMMA", "RED", "REDG", "LDGSTS"]
print("# SASS opcode counts per kernel (sm_100a cubin inside libgonova_hift.so); `total` = all instructions")
print("# UTCHMMA = tcgen05.mma, UTCBAR = tcgen05.commit, UTMALDG/UTMASTG = TMA load/store, LDTM/STTM = tcgen05.ld/st, SYNCS = mbarrier")
for fn in sorted(set(order)):
    h = hist[fn]
    print(f"{fn:70s} total {sum(h.values()):6d}  " + "  ".join(f"{k}={h[k]}" for k in KEY if h[k]))
