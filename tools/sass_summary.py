"""Per-kernel counts of the Blackwell-native SASS mnemonics in the shipped library (tcgen05.mma -> UTC*MMA, tcgen05.ld/st -> LDTM/STTM,
cp.async.bulk -> UBLKCP, TMA tensor copies -> UTMALDG/UTMASTG, legacy mma.sync -> HMMA): python tools/sass_summary.py [lib.so]"""
import collections, os, re, subprocess, sys
lib = sys.argv[1] if len(sys.argv) > 1 else os.path.join(os.path.dirname(os.path.dirname(os.path.abspath(__file__))), "magpie_tts_cpp_b200", "libmagpie_b200.so")
out = subprocess.run(["cuobjdump", "-sass", lib], capture_output=True, text=True).stdout
pats = ["UTCHMMA", "UTCQMMA", "LDTM", "STTM", "UBLKCP", "UTMALDG", "UTMASTG", "HMMA", "SYNCS", "UCGABAR"]
cur, counts = None, collections.OrderedDict()
for line in out.splitlines():
    m = re.search(r"Function : (\S+)", line)
    if m:
        name = subprocess.run(["c++filt", m.group(1)], capture_output=True, text=True).stdout.strip()
        name = name.replace("(anonymous namespace)::", "").replace("mgb::", "")
        name = re.sub(r"^void ", "", name).split("(")[0]
        cur = name
        counts.setdefault(cur, collections.Counter())
        continue
    if cur:
        for p in pats:
            if re.search(r"\b" + p + r"\b|\b" + p + r"\.", line):
                counts[cur][p] += 1
print("# SASS mnemonic counts per kernel of", os.path.basename(lib), "(cuobjdump -sass; sm_100a)")
print("%-70s " % "kernel" + " ".join("%8s" % p for p in pats))
tot = collections.Counter()
for k, c in counts.items():
    if sum(c.values()) == 0:
        continue
    print("%-70s " % k[:70] + " ".join("%8d" % c[p] for p in pats))
    tot.update(c)
print("%-70s " % "TOTAL" + " ".join("%8d" % tot[p] for p in pats))
print("# kernels in the library: %d; with tcgen05.mma (UTC*MMA): %d; with bulk copies (UBLKCP): %d; with legacy HMMA: %d" %
      (len(counts), sum(1 for c in counts.values() if c["UTCHMMA"] + c["UTCQMMA"]), sum(1 for c in counts.values() if c["UBLKCP"]), sum(1 for c in counts.values() if c["HMMA"])))
