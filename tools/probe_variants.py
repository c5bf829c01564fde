import os, sys
sys.path.insert(0, ".")
from cocons_b200 import _lib
ms = _lib.ctypes.c_double()
for n, k in ((8192, 128), (16384, 512), (32768, 512)):
    _lib.check(_lib.lib().cocons_bench_syrk(0, n, k, 3, _lib.ctypes.byref(ms)))
    flops = (n / 128) * (n / 128 + 1) / 2 * 2 * 128 * 128 * k
    print("variant", os.environ.get("COCONS_GEMM_VARIANT", "0"), "syrk", n, k, "%.3f ms %.2f TF" % (ms.value, flops / ms.value / 1e9), flush=True)
