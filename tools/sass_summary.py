"""cuobjdump -sass lisec_b200/liblisec_b200.so | python tools/sass_summary.py > profiles/sass_rN_summary.txt
Per kernel, the count of the SASS mnemonics that prove the Blackwell-native paths (B200_PROFILING.md)."""
import collections
import re
import sys

KEYS = ["UTCHMMA", "UTCQMMA", "UTCMMA", "LDTM", "STTM", "UTMALDG", "UTMASTG", "UBLKCP", "UTCBAR", "UTCCP", "SYNCS", "LDGSTS",
        "REDG", "ATOMG", "DFMA", "DADD", "FFMA"]
txt = sys.stdin.read()
print("SASS evidence for liblisec_b200.so (cuobjdump -sass, sm_100a): per kernel, the count of the mnemonics that prove the\n"
      "Blackwell-native paths: UTC*MMA = tcgen05.mma, LDTM/STTM = tcgen05.ld/st, UTMALDG/UTMASTG = TMA tensor loads/stores,\n"
      "UBLKCP = cp.async.bulk (TMA bulk copy), UTCBAR = tcgen05.commit, SYNCS = mbarrier ops, LDGSTS = cp.async.\n")
for f in re.split(r"\n\s*Function : ", txt)[1:]:
    name = f.split("\n", 1)[0].strip()
    cnt = collections.Counter()
    for line in f.split("\n"):
        m = re.search(r"^\s*/\*[0-9a-f]+\*/\s+(?:@!?U?P\d+\s+)?([A-Z0-9_]+)", line)
        if m:
            for k in KEYS:
                if m.group(1).startswith(k):
                    cnt[k] += 1
    short = re.sub(r"_ZN5lisec\d*_GLOBAL__N__[0-9a-f_]+cu_[0-9a-f]+\d*", "", name)
    print("%-84s %s" % (short[:84], " ".join("%s=%d" % (k, cnt[k]) for k in KEYS if cnt[k])))
