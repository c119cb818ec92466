"""Print the main numbers of a bench.py JSON line.  usage: show_bench.py file.json"""
import json, sys
d = json.loads(open(sys.argv[1]).read().strip().splitlines()[-1])
r = d["roofline"]
print("N=%d value %.3f G reads/s  step %.4f ms  kernel %.4f ms  frac %.4f  launches %d  clocks %s" % (d["n_gpus"], d["value"] / 1e9, d["ms_per_step"], r["kernel_ms"], r["frac"], d.get("gpu_launches", -1), d.get("clocks")))
for k in ("e2e", "e2e_soa", "e2e_bam"):
    if k in d:
        print("%-8s %.1f M reads/s  %.3f ms  h2d %d  d2h %d" % (k, d[k]["value"] / 1e6, d[k]["ms_per_step"], d[k]["h2d_bytes_per_step"], d[k]["d2h_bytes_per_step"]))
if "cpu_baseline" in d:
    c = d["cpu_baseline"]; print("cpu      %.2f M reads/s on %d threads, parity %s %s" % (c["value"] / 1e6, c["cores"], c.get("parity_on_sample"), c.get("parity_detail")))
if "ont" in d:
    o = d["ont"]; print("ont      %d reads kernel %.4f ms frac %.4f step %.4f ms err %s" % (o["reads"], o["kernel_ms"], o["frac"], o["ms_per_step"], o["device_error_flags"]))
if "file_e2e" in d:
    f = d["file_e2e"]; h = f.get("host_zlib_level_6") or f.get("deflate_level_1")
    print("file     %.2f M reads/s (%.3f s, %d bytes) split %s | host zlib: %.2f M reads/s (%d bytes) split %s" % (f["value"] / 1e6, f["seconds"], f["trimmed_bam_bytes"], f["split_s"], h["value"] / 1e6, h["trimmed_bam_bytes"], h["split_s"]))
if "deep" in d:
    x = d["deep"]; print("deep     %d reads on %d ranks: %.3f ms/step = %.2f G reads/s; kernel %.3f allreduce %.3f ins %.3f call %.3f ms; parity %s" % (x["reads"], x["ranks"], x["ms_per_step"], x["reads_per_s"] / 1e9, x["kernel_ms"], x["allreduce_ms"], x["ins_exchange_ms"], x["call_ms"], x["deep_parity"]))
