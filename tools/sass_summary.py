"""profiles/<tag>_sass_summary.txt: per kernel of libpb200.so -- registers, stack / spill bytes, shared memory (from
`cuobjdump --dump-resource-usage`) and which of the Blackwell / Hopper data-movement instructions its SASS holds
(UBLKCP = cp.async.bulk, SYNCS = mbarrier, REDUX = redux.sync, MATCH = match.any, LDGSTS, ATOMG / RED)."""
import re, subprocess, sys, collections
lib = "phanotate_b200/libpb200.so"
tag = sys.argv[1] if len(sys.argv) > 1 else "r2"
res = subprocess.run(["cuobjdump", "--dump-resource-usage", lib], capture_output=True, text=True).stdout
sass = subprocess.run(["cuobjdump", "-sass", lib], capture_output=True, text=True).stdout
dem = {}
use = collections.OrderedDict()
cur = None
for line in res.splitlines():
    m = re.search(r"Function (\S+):", line)
    if m:
        cur = m.group(1)
        continue
    if cur and "REG:" in line:
        use[cur] = dict(re.findall(r"(\w+):(\d+)", line))
        cur = None
ops = collections.defaultdict(collections.Counter)
cur = None
for line in sass.splitlines():
    m = re.search(r"Function : (\S+)", line)
    if m:
        cur = m.group(1)
        continue
    if cur:
        m = re.search(r"^\s+/\*[0-9a-f]+\*/\s+(?:@!?U?P\d+\s+)?([A-Z0-9_.]+)", line)
        if m:
            op = m.group(1).split(".")[0]
            ops[cur][op] += 1
names = subprocess.run(["c++filt"], input="\n".join(use), capture_output=True, text=True).stdout.splitlines()
arch = re.findall(r"arch = (sm_\w+)", subprocess.run(["cuobjdump", "-lelf", lib], capture_output=True, text=True).stdout + res)
with open("profiles/%s_sass_summary.txt" % tag, "w") as fh:
    fh.write("libpb200.so: cubin targets %s; %d kernels\n" % (sorted(set(arch)) or "?", len(use)))
    fh.write("%-58s %5s %6s %7s %7s %8s  %s\n" % ("kernel", "regs", "stack", "shared", "const", "instr", "UBLKCP SYNCS REDUX MATCH LDGSTS ATOM/RED FP64(DFMA/DMUL/DADD)"))
    for mangled, name in zip(use, names):
        u, o = use[mangled], ops.get(mangled, {})
        short = re.sub(r"\(.*", "", name).replace("void ", "")
        fh.write("%-58s %5s %6s %7s %7s %8d  %6d %5d %5d %5d %6d %8d %8d\n" % (
            short[:58], u.get("REG", "?"), u.get("STACK", "?"), u.get("SHARED", "?"), u.get("CONSTANT", "?").split()[0] if u.get("CONSTANT") else "?",
            sum(o.values()), o.get("UBLKCP", 0), o.get("SYNCS", 0), o.get("REDUX", 0), o.get("MATCH", 0), o.get("LDGSTS", 0),
            o.get("ATOMG", 0) + o.get("RED", 0) + o.get("ATOM", 0), o.get("DFMA", 0) + o.get("DMUL", 0) + o.get("DADD", 0)))
print(open("profiles/%s_sass_summary.txt" % tag).read()[:3000])
