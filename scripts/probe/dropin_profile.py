"""Where does the time of DCCL.build_pyramid go inside the full default bench run?  Wraps its pieces with host timers."""
import atexit
import os
import sys
import time

ROOT = os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, ROOT)
import torch  # noqa: E402
from prior_flow_b200 import corr as pcorr, ops  # noqa: E402

acc = {}


def wrap(mod, name, sync=False):
    fn = getattr(mod, name)

    def w(*a, **k):
        t0 = time.perf_counter()
        r = fn(*a, **k)
        if sync:
            torch.cuda.synchronize()
        dt = (time.perf_counter() - t0) * 1e3
        lst = acc.setdefault(name, [])
        lst.append(dt)
        return r
    setattr(mod, name, w)


wrap(pcorr, "available_device_memory")
wrap(ops, "volume_pyramid_autograd")
wrap(torch.cuda, "mem_get_info")
wrap(torch.cuda, "memory_reserved")
wrap(torch.cuda, "memory_allocated")


@atexit.register
def report():
    for k, v in acc.items():
        tail = v[-12:]
        print(f"[probe] {k}: calls {len(v)}, last 12 (ms): {[round(x, 3) for x in tail]}", file=sys.stderr)


import bench  # noqa: E402
sys.argv = ["bench.py", "--steps", "5"] + sys.argv[1:]
bench.main()
