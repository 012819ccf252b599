"""Per-kernel SASS opcode histogram of the built library -> profiles/rNN_sass_opcodes.txt.
usage: python tools/sass_opcodes.py [out.txt]   (needs cuobjdump and c++filt on PATH; no GPU)"""
import collections
import re
import subprocess
import sys

LIB = "vqvae_b200/lib/libvqvae_b200.so"
KEYS = ["UTCHMMA", "UTCQMMA", "UTCMMA", "UTCBAR", "UTCATOMSWS", "LDTM", "STTM", "UTMALDG", "UTMASTG", "UTMAPF", "SYNCS", "UBLKCP",
        "FFMA", "FMNMX", "FMNMX3", "LOP3", "ATOMG", "RED", "ATOMS", "NANOSLEEP", "ELECT", "BAR"]
HEADER = [
    "# SASS opcode histogram of %s (cuobjdump -sass, sm_100a), one line per kernel." % LIB,
    "# UTCHMMA = tcgen05.mma kind::f16/tf32, LDTM = tcgen05.ld, UTMALDG/UTMASTG = TMA tensor load/store, UTMAPF = TMA L2 prefetch,",
    "# UBLKCP = cp.async.bulk (1-D), SYNCS = mbarrier ops, UTCBAR = tcgen05.commit, UTCATOMSWS = tcgen05.alloc/dealloc.",
    "# regenerate: python tools/sass_opcodes.py", ""]


def main(out_path):
    sass = subprocess.run(["cuobjdump", "-sass", LIB], capture_output=True, text=True, check=True).stdout
    per, fn = collections.OrderedDict(), None
    for line in sass.splitlines():
        m = re.search(r"Function : (\S+)", line)
        if m:
            fn = m.group(1)
            per[fn] = collections.Counter()
            continue
        m = re.match(r"\s+/\*[0-9a-f]{4,}\*/\s+(?:@!?U?P\d+\s+)?([A-Z0-9_.]+)", line)
        if m and fn:
            per[fn][m.group(1).split(".")[0]] += 1
    names = subprocess.run(["c++filt"], input="\n".join(per), capture_output=True, text=True).stdout.split("\n")
    out = list(HEADER)
    for (fn, c), name in zip(per.items(), names):
        name = name.replace("(anonymous namespace)::", "").replace("void ", "")
        k = name.find("(")
        name = name[:k] if k > 0 else name
        out.append("%s: total=%d %s" % (name, sum(c.values()), " ".join("%s=%d" % (k, c[k]) for k in KEYS if c[k])))
    open(out_path, "w").write("\n".join(out) + "\n")


if __name__ == "__main__":
    main(sys.argv[1] if len(sys.argv) > 1 else "profiles/r02_sass_opcodes.txt")
