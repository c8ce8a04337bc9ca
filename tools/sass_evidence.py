"""Here (no GPU): counts the Blackwell-native SASS mnemonics per kernel of libbya.so (`cuobjdump -sass`) and prints one
example line of each — the listing tracked as profiles/r2_sass_tcgen05.txt.  usage: python tools/sass_evidence.py"""
import collections, os, re, subprocess, sys
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
so = os.path.join(ROOT, "bind-your-avatar-implementation_b200", "libbya.so")
out = subprocess.run(["cuobjdump", "-sass", so], capture_output=True, text=True).stdout
MN = ["UTCHMMA.2CTA", "UTCHMMA", "UTCBAR", "LDTM", "STTM", "UTMALDG", "UTMASTG", "UBLKCP", "HMMA", "MUFU.EX2", "SYNCS"]
per = collections.OrderedDict()
example = {}
name = None
for line in out.splitlines():
    m = re.search(r"Function : (\S+)", line)
    if m:
        name = subprocess.run(["c++filt", m.group(1)], capture_output=True, text=True).stdout.strip()
        name = re.sub(r"\(.*", "", name)
        per[name] = collections.Counter()
        continue
    if name is None:
        continue
    for k in MN:
        if re.search(r"\b" + re.escape(k) + r"\b", line):
            if k == "UTCHMMA" and "UTCHMMA.2CTA" in line:
                continue
            per[name][k] += 1
            example.setdefault(k, line.split("*/")[1].strip()[:110] if "*/" in line else line.strip()[:110])
print(f"SASS evidence of {os.path.relpath(so, ROOT)} (sm_100a), mnemonic counts per kernel\n")
print("| kernel | " + " | ".join(MN) + " |")
print("|---|" + "---|" * len(MN))
tot = collections.Counter()
for n, c in per.items():
    if sum(c.values()) == 0:
        continue
    tot.update(c)
    print(f"| `{n[:70]}` | " + " | ".join(str(c[k]) if c[k] else "" for k in MN) + " |")
print("| **total** | " + " | ".join(str(tot[k]) for k in MN) + " |")
print("\nexamples:")
for k in MN:
    if k in example:
        print(f"  {k:14s} {example[k]}")
