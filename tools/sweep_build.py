"""Build tuning variants of one workload's shape class (for tools/sweep_run.sh on the GPU box)."""
import sys, os, itertools
from concurrent.futures import ThreadPoolExecutor
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from neuromancer_b200 import build
case = sys.argv[1]
variants = sys.argv[2:]
out_dir = os.path.join(build.OUT_DIR, "variants")
def one(v):
    name = v.replace("=", "").replace(",", "_")
    try:
        build.build_variant(case, v, os.path.join(out_dir, f"lib_{case}_{name}.so"))
        return v, "ok"
    except Exception as e:
        return v, "FAILED " + str(e)[-300:]
with ThreadPoolExecutor(8) as ex:
    for v, st in ex.map(one, variants):
        print(v, st)
