"""SASS evidence of one kernel of the shipped library: mnemonic histogram + excerpts around the first tensor MMA / tensor copy.
usage: python tools/sass_excerpt.py <mangled-name substring> <title>"""
import collections, re, subprocess, sys
pat, title = sys.argv[1], sys.argv[2]
out = subprocess.run(["cuobjdump", "-sass", "zest_nerf_b200/libzest_b200.so"], capture_output=True, text=True).stdout
blocks = out.split("Function : ")
blk = [b for b in blocks if pat in b.split("\n")[0]]
assert blk, "no kernel matches"
b = blk[0]
lines = [l for l in b.split("\n") if re.match(r"\s+/\*[0-9a-f]{4,}\*/\s", l)]
ins = [re.sub(r"/\*[0-9a-f]+\*/\s*$", "", re.sub(r"^\s+/\*[0-9a-f]+\*/\s+", "", l)).strip() for l in lines]
mn = collections.Counter()
for i in ins:
    t = i.split()
    m = t[1] if t and t[0].startswith("@") and len(t) > 1 else (t[0] if t else "")
    mn[m.split(".")[0]] += 1
print(f"# {title}")
print(f"# cuobjdump -sass of the shipped zest_nerf_b200/libzest_b200.so (sm_100a), kernel {b.split(chr(10))[0][:100]}: {len(ins)} SASS instructions")
print("# mnemonic histogram (UTC*MMA = tcgen05.mma, UTMALDG = cp.async.bulk.tensor (tensor map), UBLKCP = cp.async.bulk, LDTM = tcgen05.ld, UTCBAR = tcgen05.commit, SYNCS = mbarrier):")
for k in ("UTCHMMA", "UTCQMMA", "UTCBAR", "UTMALDG", "UTMAPF", "UBLKCP", "LDTM", "STTM", "SYNCS", "LDG", "STG", "LDS", "STS", "FENCE", "ELECT", "R2UR"):
    if mn.get(k):
        print(f"#   {k:10s} {mn[k]}")
print("# top 20 mnemonics: " + ", ".join(f"{k} {v}" for k, v in mn.most_common(20)))
for key in ("UTMALDG", "UTCHMMA"):
    idx = [i for i, s in enumerate(ins) if key in s]
    if idx:
        print(f"\n# ---- excerpt around the first {key} (instruction {idx[0]} of {len(ins)}; {len(idx)} in the kernel)")
        for l in lines[max(0, idx[0] - 8):idx[0] + 6]:
            print(re.sub(r"\s+/\* 0x[0-9a-f]+ \*/\s*$", "", l))
