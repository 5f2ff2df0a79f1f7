#!/usr/bin/env python3
"""tools/trans_table.py SWEEP.jsonl [SWEEP_COOLDOWN.jsonl] ULP.jsonl -- the markdown table of DESIGN.md section 3:
% of the measured HBM peak per transcendental map (back-to-back sweep / with an idle second before each symbol)
and max error in ULP."""
import json, sys
fns = ["exp", "log", "sin", "cos", "tan", "asin", "acos", "atan", "sinh", "cosh", "tanh", "asinh", "acosh", "atanh", "pow", "atan2", "sqrt"]
def load(p):
    return {r["symbol"]: r for r in (json.loads(l) for l in open(p) if l.strip().startswith("{"))}
sw = load(sys.argv[1]); cool = load(sys.argv[2]) if len(sys.argv) > 3 else {}
ulp = {}
for l in open(sys.argv[-1]):
    r = json.loads(l); ulp[(r["fn"], "f32" if r["type"].startswith("f32") else "f64")] = r["max_ulp"]
print("| | " + " | ".join(fns) + " |")
print("|---|" + "---|" * len(fns))
for p, ty in (("s", "f32"), ("d", "f64")):
    cells = []
    for f in fns:
        a = sw.get("BMAS_" + p + f); c = cool.get("BMAS_" + p + f)
        t = "%d" % round(100 * a["frac_of_measured_peak"]) if a else "?"
        if c: t += " / %d" % round(100 * c["frac_of_measured_peak"])
        cells.append(t)
    print("| %s %% | " % ty + " | ".join(cells) + " |")
    print("| %s ULP | " % ty + " | ".join(("%.2f" % ulp[(f, ty)]) if (f, ty) in ulp else "0.5" for f in fns) + " |")
