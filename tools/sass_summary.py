"""SASS evidence for profiles/: per-kernel counts of the Blackwell-specific instructions in libtheia_b200.so
(UTCHMMA = tcgen05.mma, UTCBAR = tcgen05.commit, LDTM = tcgen05.ld, UTMALDG = TMA tensor load, UBLKCP = cp.async.bulk,
SYNCS = mbarrier, HMMA = legacy mma.sync -- must be 0).   python tools/sass_summary.py > profiles/r02_sass_summary.txt"""
import collections, os, re, subprocess, sys
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
so = os.path.join(ROOT, "theia_b200", "libtheia_b200.so")
out = subprocess.run(["cuobjdump", "-sass", so], capture_output=True, text=True).stdout
pat = ["UTCHMMA", "UTCBAR", "LDTM", "UTMALDG", "UBLKCP", "SYNCS", "ELECT", "MUFU", "HMMA", "STG", "RED"]
cur, tab, arch = None, collections.OrderedDict(), set()
for line in out.splitlines():
    m = re.search(r"Function : (\S+)", line)
    if m:
        cur = subprocess.run(["c++filt", m.group(1)], capture_output=True, text=True).stdout.strip()
        cur = re.sub(r"\(.*", "", cur).replace("theia::", "").replace("void ", "")
        tab[cur] = collections.Counter()
        continue
    m = re.search(r"arch = (sm_\w+)", line)
    if m:
        arch.add(m.group(1))
    if cur:
        for p in pat:
            if re.search(r"\b" + p + r"[\.\s]", line) and "/*" in line:
                key = "HMMA" if p == "HMMA" and "UTCHMMA" not in line else p
                if p == "HMMA" and "UTCHMMA" in line:
                    continue
                tab[cur][key] += 1
print("libtheia_b200.so:", ", ".join(sorted(arch)))
print(f"{'kernel':58s} " + " ".join(f"{p:>7s}" for p in pat))
tot = collections.Counter()
for k, c in tab.items():
    if not any(c[p] for p in pat[:5]) and c["SYNCS"] == 0:
        continue
    print(f"{k[:58]:58s} " + " ".join(f"{c[p]:7d}" for p in pat))
    tot.update(c)
print(f"{'TOTAL (kernels with tcgen05 / TMA / mbarrier)':58s} " + " ".join(f"{tot[p]:7d}" for p in pat))
print("kernels in the library:", len(tab), "| legacy HMMA (mma.sync) instructions:", sum(c['HMMA'] for c in tab.values()))
