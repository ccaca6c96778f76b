"""Opcode histogram of the hot kernels' SASS (cuobjdump -sass of the built library) for profiles/.
usage: python scripts/sass_histogram.py giwaxsim_b200/libgiwaxs_b200.so out.md pattern [pattern ...]
A pattern selects functions whose mangled name contains it."""
import collections, re, subprocess, sys

so, out, pats = sys.argv[1], sys.argv[2], sys.argv[3:]
txt = subprocess.run(["cuobjdump", "-sass", so], capture_output=True, text=True).stdout
funcs, cur = collections.OrderedDict(), None
for line in txt.splitlines():
    m = re.search(r"Function : (\S+)", line)
    if m:
        cur = m.group(1)
        funcs[cur] = collections.Counter()
        continue
    m = re.match(r"\s+/\*[0-9a-f]{4,}\*/\s+(?:@!?U?P\d+\s+)?([A-Z][A-Z0-9_]*)((?:\.[A-Z0-9_]+)*)", line)
    if m and cur:
        funcs[cur][m.group(1)] += 1
        if m.group(1) in ("ATOMS", "RED", "ATOMG", "UTMALDG", "SYNCS", "LDS", "STS", "LDG", "STG", "BAR"):
            funcs[cur][m.group(1) + m.group(2)] += 0          # keep the qualified spelling visible below
with open(out, "w") as fh:
    fh.write("# SASS opcode histograms (cuobjdump -sass %s)\n\n" % so.split("/")[-1])
    fh.write("Static instruction counts per kernel (not executed counts). Blackwell-specific mnemonics: `UTMALDG` = "
             "cp.async.bulk.tensor (TMA load), `SYNCS` = mbarrier, `FADD2`/`FFMA2`/`FMUL2` = packed fp32x2.\n\n")
    for name, c in funcs.items():
        if not any(p in name for p in pats):
            continue
        total = sum(v for k, v in c.items() if "." not in k)
        fh.write("## `%s` (%d instructions)\n\n" % (name, total))
        top = [(k, v) for k, v in c.most_common() if "." not in k and v > 0]
        fh.write("| opcode | count |\n|---|---|\n")
        for k, v in top[:28]:
            fh.write("| %s | %d |\n" % (k, v))
        special = {k: c[k] for k in ("UTMALDG", "SYNCS", "FADD2", "FFMA2", "FMUL2", "ATOMS", "RED", "ATOMG", "I2F", "I2FP", "DFMA", "DMUL", "DADD")
                   if c.get(k)}
        fh.write("\nspecial: %s\n\n" % ", ".join("%s x %d" % kv for kv in special.items()))
print("wrote", out)
