"""Per-kernel SASS opcode histogram of libsihmpc_b200.so (cuobjdump -sass).  usage: sass_histogram.py [lib] > profiles/rNN_sass_histogram.md"""
import collections, re, subprocess, sys
lib = sys.argv[1] if len(sys.argv) > 1 else "sihmpc_b200/libsihmpc_b200.so"
out = subprocess.run(["cuobjdump", "-sass", lib], capture_output=True, text=True).stdout
kern, hist, arch = None, collections.OrderedDict(), {}
for line in out.splitlines():
    m = re.match(r"\s*Function : (\S+)", line)
    if m:
        kern = m.group(1); hist[kern] = collections.Counter(); continue
    m = re.match(r"\s*arch = (\S+)", line)
    if m:
        last_arch = m.group(1); continue
    m = re.match(r"\s*/\*[0-9a-f]{4,}\*/\s+(?:@!?U?P\d+\s+)?([A-Z][A-Z0-9_]*)", line)
    if m and kern:
        hist[kern][m.group(1)] += 1
demangle = subprocess.run(["c++filt"], input="\n".join(hist), capture_output=True, text=True).stdout.splitlines()
keys = ["DMMA", "DFMA", "DADD", "DMUL", "MUFU", "LDG", "STG", "LDS", "STS", "SHFL", "BAR", "LDL", "STL", "ATOMS", "ATOMG", "RED",
        "UTMALDG", "UTCHMMA", "LDTM", "HMMA", "IMMA"]
print("# SASS opcode histogram per kernel of `%s` (sm_100a cubins, `cuobjdump -sass`)\n" % lib)
print("FP64 path: `DMMA` = FP64 tensor-core MMA (mma.sync.m8n8k4.f64), `DFMA/DADD/DMUL` = FP64 pipe.  There is no FP64 on "
      "tcgen05 (`UTC*MMA`, `LDTM` absent by design); `UTMALDG` (TMA) unused: constants are streamed with coalesced `LDG` "
      "from L2 or staged once per CTA in shared memory.  `LDL/STL` = local-memory spills.\n")
print("| kernel | total | " + " | ".join(keys) + " |")
print("|---|---|" + "---|" * len(keys))
for (k, h), d in zip(hist.items(), demangle):
    name = re.sub(r"\(.*", "", d).replace("void ", "").replace("sih::", "")
    print("| `%s` | %d | " % (name[:90], sum(h.values())) + " | ".join(str(h.get(x, 0)) for x in keys) + " |")
