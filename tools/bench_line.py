#!/usr/bin/env python
"""One-line digest of a bench.py JSON line (stdin or file)."""
import json
import sys
d = json.loads((open(sys.argv[1]) if len(sys.argv) > 1 else sys.stdin).read().strip().splitlines()[-1])
pk = d["roofline"]["per_kernel"]
e2e = d["e2e"]["value"] if d.get("e2e") else float("nan")
r = d["roofline"]
print("%.1f img/s | e2e %.1f | %.2f ms/step | top kernel %.0f TF/s frac %.3f | gemm family %.0f TF/s frac %.3f | step frac %.3f | traffic %s" % (
    d["value"], e2e, d["ms_per_step"], r["achieved"], r["frac"], r.get("gemm_family_tflops", 0), r.get("gemm_family_frac", 0),
    r["whole_step_frac"], r.get("traffic")))
if d.get("cls_only_last_block"):
    c = d["cls_only_last_block"]
    print("  cls_only_last_block %.1f img/s (%.2f ms/step, top-5 agreement %.4f)" % (c["value"], c["ms_per_step"], c["top5_agreement_with_full_schedule"]))
if d.get("e2e_from_images"):
    print("  e2e_from_images %.1f img/s" % d["e2e_from_images"]["value"])
print("  " + " ".join("%s=%.2f(%s)" % (k, v["ms_per_step"], ("%.0fTF" % v["tflops"]) if "tflops" in v else ("%.0fGB/s" % v.get("gbs", 0)))
                      for k, v in pk.items()))
print("  clocks", d["clocks"], "launches", d["gpu_launches"])
if d.get("cpu_baseline"):
    print("  cpu", d["cpu_baseline"]["value"], d["cpu_baseline"]["cores"], d["cpu_baseline"].get("top5_agreement_with_gpu"))
