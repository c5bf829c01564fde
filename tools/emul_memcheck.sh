#!/bin/bash
# Memory check of the SHIPPED kernels without a GPU: the host build of tests/host_emul compiled with AddressSanitizer
# and UBSan ("device" memory is the heap, "shared" memory are globals), the emulation tests run on top of it; the R glue
# and the miniature R runtime of tests/rmock are instrumented as well (marshalling of R objects).
# compute-sanitizer cannot be used on the GPU pool; this is the closest substitute for its memcheck tool.
cd "$(dirname "$0")/.."
export COCONS_EMUL_SANITIZE=1
export LD_PRELOAD="$(gcc -print-file-name=libasan.so) $(gcc -print-file-name=libubsan.so)"
export ASAN_OPTIONS=detect_leaks=0:abort_on_error=0:halt_on_error=1
exec python -m pytest tests/test_host_emul.py tests/test_product_on_host.py tests/test_rglue.py -m "not gpu" -x -q \
  --deselect tests/test_host_emul.py::test_race_check_of_the_shipped_kernels "$@"
