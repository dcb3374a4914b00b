"""Regenerates example_fa_digests.json from the CPU oracle (run from anywhere)."""
import json
import os
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, ROOT)
import oracle  # noqa: E402

seq = "".join(l.strip() for l in open(os.path.join(ROOT, "tests/golden/example.fa")) if not l.startswith(">"))
out = {"_comment": "Derived with oracle/ (not outputs of the reference binary, which cannot be built here); "
                   "regenerate with tests/golden/make_example_digests.py. n matches the reference's published "
                   "349,910 / 349,900 (README.md:94-99, doc/api.md:16-25)."}
for k in (21, 31):
    t = oracle.OracleTable(k)
    n = t.consume(seq)
    d = t.digest()
    out[str(k)] = {"n_kmers": n, "distinct": d["n"], "sum": d["sum"], "min": t.min, "max": t.max,
                   "histo": t.histo(zero=False), "xor": d["xor"], "sum_hc": d["sum_hc"], "sha256": t.sha256_sorted()}
json.dump(out, open(os.path.join(ROOT, "tests/golden/example_fa_digests.json"), "w"), indent=1)
print(json.dumps(out)[:300])
