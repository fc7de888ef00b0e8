"""Per kernel of libvoxel_cuda.so: instruction count and the FP64 / TMA mnemonics that back the
claims of DESIGN.md (no DFMA: -fmad=false; UBLKCP = cp.async.bulk stores).
usage: python profiles/scripts/sass_summary.py [lib.so] > profiles/sass_summary.txt"""
import collections, re, subprocess, sys
lib = sys.argv[1] if len(sys.argv) > 1 else "voxel_game_b200/libvoxel_cuda.so"
sass = subprocess.run(["cuobjdump", "-sass", lib], capture_output=True, text=True, check=True).stdout
keys = ["DFMA", "DADD", "DMUL", "DSETP", "DMNMX", "FFMA", "UBLKCP", "UTMALDG", "UTMASTG", "LDS", "STS", "LDG", "STG", "BAR"]
print(f"# {lib}: cuobjdump -sass, instruction counts per kernel (static, not executed)")
print("kernel".ljust(28) + "total".rjust(7) + "".join(k.rjust(8) for k in keys))
tot = collections.Counter()
for f in re.split(r"\n\s*Function : ", sass)[1:]:
    name = f.split("\n")[0]
    short = name[:27]
    end = name.find("_kernel") + 7
    for n_chars in range(8, 40):  # Itanium mangling: <length><name>
        if end >= 7 and name[end - n_chars - 2:end - n_chars] == f"{n_chars:02d}":
            short = name[end - n_chars:end]
            break
    if "ILb1" in name: short += "<emit>"
    if "ILb0" in name: short += "<sizes>"
    ops = collections.Counter()
    for line in f.split("\n"):
        mm = re.match(r"\s+/\*[0-9a-f]{4}\*/\s+(@!?U?P\d\s+)?([A-Z0-9_]+)", line)
        if mm: ops[mm.group(2)] += 1
    tot.update(ops)
    print(short.ljust(28) + str(sum(ops.values())).rjust(7) + "".join(str(ops[k]).rjust(8) for k in keys))
print("ALL".ljust(28) + str(sum(tot.values())).rjust(7) + "".join(str(tot[k]).rjust(8) for k in keys))
