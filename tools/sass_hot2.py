#!/usr/bin/env python3
"""tools/sass_hot2.py <Functor> [f32|f64] [lib] -- static per-element instruction mix of the all-vector tile of
k_map_flat<Functor, U, 0>: instructions between the first global load and the last store of the first load/store run,
divided by the elements that run stores (16-byte packs); grouped by pipe (DP: FP64, XU: MUFU + conversions)."""
import collections, re, subprocess, sys
F = sys.argv[1]; ty = sys.argv[2] if len(sys.argv) > 2 else "f32"
lib = sys.argv[3] if len(sys.argv) > 3 else "numericals_b200/lib/libnuk.so"
elf = subprocess.run(["cuobjdump", "-elf", lib], capture_output=True, text=True).stdout
names = sorted(set(re.findall(r"_ZN3nuk10k_map_flatINS_[0-9]*%sELi[0-9]*ELi0EE[A-Za-z0-9_]*" % F, elf)))
if not names: sys.exit("no kernel for " + F)
sass = subprocess.run(["cuobjdump", "-sass", "-fun", names[0], lib], capture_output=True, text=True).stdout
ops = [re.sub(r"^\s*/\*[0-9a-f]+\*/\s*", "", l).split(";")[0].strip() for l in sass.splitlines() if re.match(r"\s+/\*[0-9a-f]{4,}\*/", l)]
ld = [i for i, o in enumerate(ops) if "LDG" in o]; st = [i for i, o in enumerate(ops) if "STG" in o]
nst = [i for i in st if not any(st[0] < j < i for j in ld)]
body = ops[ld[0]:nst[-1] + 1]
per = 16 // (8 if ty == "f64" else 4)
elems = len(nst) * per
def opn(o):
    t = o.split()
    return (t[1] if t[0].startswith("@") else t[0]).split(".")[0]
c = collections.Counter(opn(o) for o in body)
dp = sum(v for k, v in c.items() if k in ("DFMA", "DADD", "DMUL", "DSETP"))
xu = sum(v for k, v in c.items() if k in ("MUFU", "F2F", "I2F", "F2I", "FRND", "I2FP", "F2FP"))
lds = c.get("LDS", 0)
print("%-8s %3d instr/elem  DP %.1f  XU %.1f  LDS %.1f  other %.1f  (tile %d instr, %d elems, kernel %d) top: %s" % (
    F, round(len(body) / elems), dp / elems, xu / elems, lds / elems, (len(body) - dp - xu - lds) / elems, len(body), elems, len(ops),
    " ".join("%s:%.1f" % (k, v / elems) for k, v in c.most_common(10))))
