"""SASS instruction counts of libsam2b200.so (cuobjdump -sass): total and per kernel.  usage: python scripts/sass_counts.py > profiles/rN_sass_counts.txt"""
import os, re, subprocess, collections
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
LIB = os.path.join(ROOT, "sam2_video_training_b200", "libsam2b200.so")
out = subprocess.run(["cuobjdump", "-sass", LIB], capture_output=True, text=True, check=True).stdout
ops = ["UTCHMMA", "LDTM", "STTM", "UTMALDG", "UTMASTG", "UTCBAR", "UTCATOMSWS", "SYNCS", "REDG"]
tot = collections.Counter(); per = collections.OrderedDict(); cur = None; hmma = 0
for line in out.splitlines():
    m = re.search(r"Function : (\S+)", line)
    if m:
        cur = m.group(1); per[cur] = collections.Counter(); continue
    for o in ops:
        if re.search(r"\b" + o + r"\b|\b" + o + r"\.", line):
            tot[o] += 1
            if cur: per[cur][o] += 1
    if re.search(r"\bHMMA\b|\bHMMA\.", line): hmma += 1
print("# SASS instruction counts of sam2_video_training_b200/libsam2b200.so (cuobjdump -sass, sm_100a), end of round 2 (scripts/sass_counts.py).")
print("# UTCHMMA = tcgen05.mma, LDTM / STTM = tcgen05.ld / st (tensor memory), UTMALDG / UTMASTG = TMA tensor load / store,")
print("# UTCBAR = tcgen05.commit -> mbarrier, UTCATOMSWS = TMEM alloc; HMMA (mma.sync fallback) must be 0.")
for o in ops: print(f"{o:12s} {tot[o]}")
print(f"HMMA (not UTCHMMA)  {hmma}")
print("\n# per kernel (UTCHMMA / LDTM / UTMALDG / UTMASTG), kernels that use the tensor cores:")
for fn, c in per.items():
    if c["UTCHMMA"] == 0: continue
    name = subprocess.run(["c++filt", fn], capture_output=True, text=True).stdout.strip()
    name = re.sub(r"CUtensorMap_st(, CUtensorMap_st)*", "CUtensorMap_st...", name)
    print(f"{c['UTCHMMA']:6d} {c['LDTM']:6d} {c['UTMALDG']:6d} {c['UTMASTG']:6d}  {name[:150]}")
