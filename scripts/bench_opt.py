"""bench.py with library options set first: python scripts/bench_opt.py name=value [name=value ...] -- <bench.py arguments>"""
import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from vit_rpe_rope_b200 import _lib
split = sys.argv.index("--") if "--" in sys.argv else len(sys.argv)
lib = _lib.load()
for kv in sys.argv[1:split]:
    k, v = kv.split("=")
    _lib.check(lib.vrr_set_option(k.encode(), int(v)), k)
sys.argv = ["bench.py"] + sys.argv[split + 1:]
import bench
bench.main()
