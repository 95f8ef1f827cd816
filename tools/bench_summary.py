"""Runs bench.py with the given arguments and prints the key numbers of its JSON line (GPU box helper)."""
import json, subprocess, sys
r = subprocess.run([sys.executable, "bench.py"] + sys.argv[1:], capture_output=True, text=True)
ok = False
for l in r.stdout.splitlines():
    if l.startswith("{"):
        j = json.loads(l); ok = True
        print("value %.4e  step %.3f ms  kernel %.3f ms  frac %.4f  e2e %.4e  tail %.2f ms" % (
            j["value"], j["ms_per_step"], j["roofline"]["launch_ms"], j["roofline"]["frac"], j["e2e"]["value"], j["tail_ms"]))
        print("series", {k: "%.3f ms" % v["ms_per_step"] for k, v in j["series"].items()})
        for k in ("fit", "sweep", "allreduce_check"):
            if k in j:
                print(k, {a: b for a, b in j[k].items() if a not in ("what", "workload")})
if not ok:
    print("NO JSON LINE\n", r.stdout[-1500:], r.stderr[-3000:])
