"""
SASS opcode census of libverde_b200.so (no GPU needed): per kernel, the count of the instructions that identify how it
computes and moves data.  `python tools/sass_census.py > profiles/r02_sass_census.json`
"""
import collections, json, os, re, subprocess, sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
lib = os.path.join(ROOT, "verde_b200", "libverde_b200.so")
text = subprocess.run(["cuobjdump", "-sass", lib], capture_output=True, text=True, check=True).stdout
WATCH = ["DMMA", "DFMA", "DADD", "DMUL", "MUFU", "UBLKCP", "UBLKRED", "SYNCS", "RED", "REDG", "REDUX", "ATOM", "ATOMG", "ATOMS", "LDS", "STS", "LDG",
         "STG", "LDGSTS", "SHFL", "MATCH", "BAR", "UTCHMMA", "UTCIMMA", "UTMALDG", "LDTM", "HMMA", "IMMA", "NANOSLEEP", "ERRBAR", "MEMBAR"]
out = {"library": "verde_b200/libverde_b200.so", "arch": sorted(set(re.findall(r"arch = (sm_\w+)", text))), "kernels": {}}
name, counts, total = None, None, 0
for line in text.splitlines():
    m = re.match(r"\s*Function : (\S+)", line)
    if m:
        name = subprocess.run(["c++filt", m.group(1)], capture_output=True, text=True).stdout.strip()
        name = re.sub(r"\(.*", "", name.replace("(anonymous namespace)::", ""))
        counts = out["kernels"].setdefault(name, collections.Counter())
        continue
    m = re.match(r"\s*/\*[0-9a-f]{4}\*/\s+(?:@!?U?P\d+\s+)?([A-Z0-9_]+)(?:\.[A-Za-z0-9_.]+)?\s", line)
    if m and counts is not None:
        counts["instructions"] += 1
        op = m.group(1)
        if op in WATCH:
            counts[op] += 1
out["kernels"] = {k: dict(sorted(v.items())) for k, v in sorted(out["kernels"].items())}
tot = collections.Counter()
for v in out["kernels"].values():
    tot.update(v)
out["total"] = dict(sorted(tot.items()))
out["note"] = ("FP64 has no tcgen05 kind (SURVEY fact 10): the matrix products are DMMA.8x8x4 fed by TMA bulk copies (UBLKCP) and "
               "mbarriers (SYNCS), committed with RED.ADD.F64 / UBLKRED; UTC*MMA / LDTM / UTMALDG are therefore expected to be 0")
json.dump(out, sys.stdout, indent=1)
print()
