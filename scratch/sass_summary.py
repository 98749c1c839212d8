"""profiles/sass_summary.txt: per kernel of libsilent_b200, the SASS mnemonics that show how it is built (TMA loads /
prefetches, bulk copies, mbarrier waits, packed fp32 math, byte permutes) -- from cuobjdump -sass of the built objects."""
import collections, glob, os, re, subprocess, sys
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
KEYS = ["UTMALDG", "UTMAPF", "UTMASTG", "UBLKCP", "UBLKPF", "SYNCS", "FFMA2", "FADD2", "FMUL2", "FFMA", "PRMT", "LDS.128", "STS.128",
        "LDG.E.128", "TLD", "REDUX", "CREDUX", "ATOMS", "BAR.SYNC"]
out = ["SASS mnemonic counts per kernel (static instruction counts, cuobjdump -sass of pysilent_b200/build/*.o, sm_100a)", ""]
for obj in sorted(glob.glob(os.path.join(ROOT, "pysilent_b200", "build", "*.o"))):
    txt = subprocess.run(["cuobjdump", "-sass", obj], capture_output=True, text=True).stdout
    cur, counts = None, collections.OrderedDict()
    for line in txt.splitlines():
        m = re.search(r"Function : (\S+)", line)
        if m:
            cur = subprocess.run(["c++filt", m.group(1)], capture_output=True, text=True).stdout.strip().split("(")[0]
            cur = cur.replace("void ", "").replace("silent::", "")
            counts[cur] = collections.Counter()
            continue
        if cur is None:
            continue
        m = re.match(r"\s+/\*[0-9a-f]+\*/\s+(?:@!?U?P\d+\s+)?([A-Z0-9_.]+)", line)
        if m:
            op = m.group(1)
            counts[cur]["total"] += 1
            for k in KEYS:
                if op == k or op.startswith(k + ".") or (k in ("LDS.128", "STS.128", "LDG.E.128") and op.startswith(k)):
                    counts[cur][k] += 1
    for name, c in counts.items():
        if c["total"] < 200 and not any(c[k] for k in KEYS[:6]):
            continue
        out.append("%s  [%s]" % (name, os.path.basename(obj)))
        out.append("    total %d; " % c["total"] + ", ".join("%s %d" % (k, c[k]) for k in KEYS if c[k]))
open(os.path.join(ROOT, "profiles", "sass_summary.txt"), "w").write("\n".join(out) + "\n")
print("\n".join(out[:60]))
