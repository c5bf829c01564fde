#!/bin/bash
# Sweep of the outer panel width of the factorisation (K of the trailing update): standalone SYRK rate and
# the whole evaluation at n = 50 000.  Usage (GPU box): bash tools/outer_sweep.sh > gpurun_out/outer_sweep.log
python - <<'PY'
import ctypes, sys
sys.path.insert(0, ".")
from cocons_b200 import _lib
L = _lib.lib()
for k in (256, 512, 768, 1024):
    ms = ctypes.c_double()
    _lib.check(L.cocons_bench_syrk(0, 32768, k, 5, ctypes.byref(ms)))
    fl = 32768 * 32769 / 2 * 2 * k
    print("SYRK n=32768 K=%d: %.2f ms  %.2f TFLOP/s" % (k, ms.value, fl / ms.value / 1e9), flush=True)
PY
for o in 4 6 8; do
  COCONS_CHOL_OUTER=$o python bench.py --steps 4 --warmup 3 --no-cpu-baseline 2>/dev/null | python -c "
import json,sys
d=json.loads(sys.stdin.read().strip().splitlines()[-1])
print('OUTER=$o value %.4f evals/s  step %.1f ms  phases %s  kernel %.2f TF  chol %.2f TF' % (d['value'], d['ms_per_step'], {k: round(v,1) for k,v in d['phases_ms'].items()}, d['roofline']['achieved'], d['roofline']['cholesky_phase']['achieved']))"
done
