"""One-line digest of a bench.py JSON line: value, e2e, per-iteration phases.  python tools/phase_line.py FILE"""
import json
import sys


def main(path):
    line = None
    for ln in open(path):
        ln = ln.strip()
        if ln.startswith("{"):
            line = json.loads(ln)
    if line is None:
        print("no JSON line in", path)
        return
    its = (line.get("details") or line["config"]).get("iterations_per_solve", 0) or 1
    ph = line.get("phases_ms_per_solve", {})
    e2e = line.get("e2e") or {}
    rf = line.get("roofline") or {}
    print("n_gpus=%s value=%.3f it/s e2e=%s ms/solve=%.1f its=%s | per it: %s | K1 %.2f TF/s frac %.3f potrf %.2f ms" % (
        line.get("n_gpus"), line["value"], ("%.3f" % e2e["value"]) if e2e.get("value") else None, line["ms_per_step"], its,
        " ".join("%s=%.2f" % (k.replace("_ms", ""), v / its) for k, v in ph.items()),
        rf.get("achieved") or 0.0, rf.get("frac") or 0.0, rf.get("potrf_ms_per_launch") or 0.0))
    if "c5" in line:
        print("   c5:", json.dumps(line["c5"]))


if __name__ == "__main__":
    main(sys.argv[1])
