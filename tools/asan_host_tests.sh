#!/bin/bash
# Host library (libmcmc.so) under AddressSanitizer + UndefinedBehaviorSanitizer: builds an instrumented
# copy into /tmp and runs the CPU tests of the host side against it.  No GPU involved.
set -e
ROOT="$(cd "$(dirname "$0")/.." && pwd)"
OUT=/tmp/ammsb_asan
mkdir -p "$OUT"
cd "$ROOT/mcmc-ammsb-gpu_b200/host"
SRCS="mcmc/types.cc mcmc/cuckoo.cc mcmc/data.cc mcmc/sample.cc mcmc/config.cc mcmc/random.cc mcmc/serialize.cc
      mcmc/phi.cc mcmc/beta.cc mcmc/perplexity.cc mcmc/learner.cc capi.cc"
g++ -O1 -g -fsanitize=address,undefined -fno-omit-frame-pointer -std=c++17 -fPIC -Wno-deprecated-declarations \
    -I . -I ../../include -shared -o "$OUT/libmcmc.so" $SRCS -L.. -lammsb -lz -lpthread -Wl,-rpath,"$ROOT/mcmc-ammsb-gpu_b200"
cat > "$OUT/run.py" <<PY
import sys
sys.path[:0] = ["$ROOT/tests", "$ROOT/mcmc-ammsb-gpu_b200", "$ROOT/oracle"]
import pymcmc
pymcmc.LIB_PATH = "$OUT/libmcmc.so"
import pytest
sys.exit(pytest.main(["-x", "-q", "-p", "no:cacheprovider", "$ROOT/tests/test_host.py",
                      "$ROOT/tests/test_checkpoint_wire.py", "$ROOT/tests/test_data_formats.py"]))
PY
LD_PRELOAD="$(gcc -print-file-name=libasan.so):$(gcc -print-file-name=libubsan.so)" \
  ASAN_OPTIONS=detect_leaks=0:halt_on_error=1 UBSAN_OPTIONS=print_stacktrace=1 \
  python "$OUT/run.py" 2>&1 | tee "$OUT/log.txt" | tail -3
if grep -q "runtime error\|AddressSanitizer" "$OUT/log.txt"; then echo "sanitizer findings: see $OUT/log.txt"; exit 1; fi
echo "sanitizers: clean"
