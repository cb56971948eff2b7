"""Per-kernel counts of the SASS mnemonics that prove which hardware path a kernel uses
(B200_PROFILING.md: tcgen05.mma -> UTC*MMA (cta_group::2 -> UTCHMMA.2CTA), tcgen05.ld/st -> LDTM/STTM, TMA -> UTMALDG/UTMASTG,
mma.sync -> HMMA/*MMA.16816..., clusters -> UCGABAR / cluster barrier, PDL -> ACQBULK/...):
    python profiles/sass_mnemonics.py > profiles/r01_sass_mnemonics.txt
Runs offline (cuobjdump on the built .so), no GPU needed."""
import collections, os, re, subprocess, sys
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
SO = os.path.join(ROOT, "sign-language-nlp_b200", "libslnlp_b200.so")
PAT = ("UTCHMMA", "UTCHMMA.2CTA", "UTCQMMA", "UTCIMMA", "UTCMMA", "UTCBAR", "LDTM", "STTM", "UTMALDG", "UTMASTG", "UBLKCP", "HMMA", "IMMA",
       "SYNCS", "UCGABAR", "MUFU.TANH", "MUFU.EX2", "RED.E.ADD", "REDG", "ATOMG", "LDGSTS", "ACQBULK", "ELECT")
out = subprocess.run(["cuobjdump", "-sass", SO], capture_output=True, text=True).stdout
counts, name = collections.OrderedDict(), None
for line in out.splitlines():
    m = re.search(r"Function : (\S+)", line)
    if m:
        name = subprocess.run(["c++filt", m.group(1)], capture_output=True, text=True).stdout.strip()
        name = re.sub(r"\(.*", "", name).replace("void ", "").replace("slnlp::", "")
        counts[name] = collections.Counter()
        continue
    if name is None:
        continue
    m = re.search(r"^\s+/\*[0-9a-f]+\*/\s+(?:@!?U?P\d+\s+)?([A-Z0-9_.]+)", line)
    if m:
        op = m.group(1)
        for p in PAT:
            if op.startswith(p):
                counts[name][p] += 1
print(f"# SASS mnemonic counts per kernel of {os.path.basename(SO)} (cuobjdump -sass, sm_100a); kernels without any are omitted")
for k, c in counts.items():
    if c:
        print(f"{k[:70]:70s} " + "  ".join(f"{p}={n}" for p, n in sorted(c.items())))
