"""python tools/run_with_options.py key=value [key=value ...] -- script.py [args]: set library tunables, then run a tool in this process."""
import os, runpy, sys
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
from verde_b200 import _cabi
split = sys.argv.index("--")
lib = _cabi.require_device()
for item in sys.argv[1:split]:
    key, _, val = item.partition("=")
    _cabi.check(lib.vb200_set_option(key.encode(), float(val)), "set_option " + item)
sys.argv = sys.argv[split + 1:]
runpy.run_path(sys.argv[0], run_name="__main__")
