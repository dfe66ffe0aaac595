"""A/B helper: run a script against another build of the library.
    python scripts/with_lib.py path/to/libb4cp_other.so bench.py --steps 20 ...
Points the loader (bert4clickpath_b200._lib.LIB_PATH) at the given file, then runs the script
as __main__.  Still the CUDA library or nothing - there is no other path to switch to."""
import os, runpy, sys
sys.path.insert(0, os.getcwd())
from bert4clickpath_b200 import _lib  # noqa: E402

_lib.LIB_PATH = os.path.abspath(sys.argv[1])
sys.argv = sys.argv[2:]
runpy.run_path(sys.argv[0], run_name="__main__")
