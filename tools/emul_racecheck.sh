#!/bin/bash
# Race check of the SHIPPED kernels without a GPU (the closest substitute for compute-sanitizer's racecheck, which the
# GPU pool does not allow): host build of every library source under ThreadSanitizer, every CUDA thread a TSan fiber,
# happens-before only through the kernels' own barriers / mbarriers / atomics.  See tests/host_emul/racecheck_main.cpp.
#   tools/emul_racecheck.sh            all kernel families once; silent = no race
#   tools/emul_racecheck.sh --racy     a deliberately racy kernel: must be reported
#   COCONS_EMUL_DROP_HANDBACK=1 tools/emul_racecheck.sh   mutation: ring-slot hand-back ordering removed: must be reported
cd "$(dirname "$0")/.."
work=$(mktemp -d)
exe=$(python -c "import sys; sys.path.insert(0, 'tests'); from host_emul import build; print(build.build_racecheck('$work'))" 2>/dev/null | tail -1)
TSAN_OPTIONS="halt_on_error=0" "$exe" "$@"
