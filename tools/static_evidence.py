"""static evidence under profiles/: SASS instruction mix per kernel (cuobjdump -sass of the built library) and the ptxas -v
table (registers / spills / stack from csrc/*.ptxas.log).  usage: python tools/static_evidence.py [tag]   (default r2)"""
import collections
import glob
import os
import re
import subprocess
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
tag = sys.argv[1] if len(sys.argv) > 1 else "r2"
lib = os.path.join(ROOT, "decision-making-and-path-planning_b200", "libdmpp_b200.so")
sass = subprocess.run(["cuobjdump", "-sass", lib], capture_output=True, text=True).stdout
def pretty(name):
    d = subprocess.run(["c++filt", name], capture_output=True, text=True).stdout.strip() or name
    d = re.sub(r"\(anonymous namespace\)::", "", d)
    d = re.sub(r"^void ", "", d)
    return d.split("(")[0]


kern, mix = None, collections.OrderedDict()
for line in sass.splitlines():
    m = re.search(r"Function : (\S+)", line)
    if m:
        kern = pretty(m.group(1)); mix[kern] = collections.Counter(); continue
    m = re.match(r"\s+/\*[0-9a-f]+\*/\s+(?:@!?U?P\d+\s+)?([A-Z][A-Z0-9_.]*)", line)
    if m and kern:
        mix[kern][m.group(1).split(".")[0]] += 1; mix[kern]["_total"] += 1
cols = ["DFMA", "DADD", "DMUL", "DSETP", "FFMA", "FMUL", "FADD", "MUFU", "UBLKCP", "SYNCS", "LDS", "LDG", "LD", "STS", "STG", "ST", "ATOMS", "ATOMG", "ATOM", "RED",
        "BAR", "SHFL", "REDUX", "UCGABAR_ARV", "UCGABAR_WAIT", "MEMBAR", "HMMA", "UTCMMA", "DMMA"]
with open(os.path.join(ROOT, "profiles", "%s_sass_mix.txt" % tag), "w") as f:
    f.write("cuobjdump -sass decision-making-and-path-planning_b200/libdmpp_b200.so: instruction mix per kernel (static counts; tools/static_evidence.py)\n"
            "UBLKCP + SYNCS = the TMA bulk copy (cp.async.bulk) of the carried paths and its mbarrier; UCGABAR_* = the cluster barrier of the dense sweep;\n"
            "no tensor-core instruction (HMMA / DMMA / UTCMMA) anywhere: the path has no contraction with K > 4 and computes in FP64.\n\n")
    for k, c in mix.items():
        f.write("%-48s %6d SASS instr | %s\n" % (k[:48], c["_total"], " ".join("%s %d" % (n, c[n]) for n in cols if c[n])))
with open(os.path.join(ROOT, "profiles", "%s_ptxas_table.txt" % tag), "w") as f:
    f.write("ptxas -v: registers / spills / stack of every kernel (csrc/*.ptxas.log, nvcc -Xptxas -v, sm_100a; tools/static_evidence.py)\n")
    for log in sorted(glob.glob(os.path.join(ROOT, "decision-making-and-path-planning_b200", "csrc", "*.o.ptxas.log"))):
        fn = None; frame = ""
        for line in open(log):
            m = re.search(r"Compiling entry function '(\S+)'", line)
            if m:
                fn = pretty(m.group(1)); continue
            if "bytes stack frame" in line and fn:
                frame = line.strip(); continue
            m = re.search(r"Used (\d+) registers.*", line)
            if m and fn:
                f.write("%-56s %s | %s\n" % (fn[:56], m.group(0).replace("ptxas info    : ", ""), frame)); fn = None
print("written")
