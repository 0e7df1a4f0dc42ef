#!/bin/bash
# The host mini-batch strategies under ThreadSanitizer (tools/tsan_sampler.cc).  No GPU involved.
set -e
ROOT="$(cd "$(dirname "$0")/.." && pwd)"
OUT=/tmp/ammsb_tsan
mkdir -p "$OUT"
cd "$ROOT/mcmc-ammsb-gpu_b200/host"
SRCS="mcmc/types.cc mcmc/cuckoo.cc mcmc/data.cc mcmc/sample.cc mcmc/config.cc mcmc/random.cc mcmc/serialize.cc
      mcmc/phi.cc mcmc/beta.cc mcmc/perplexity.cc mcmc/learner.cc"
g++ -O1 -g -fsanitize=thread -std=c++17 -fPIE -Wno-deprecated-declarations -I . -I ../../include \
    -o "$OUT/tsan_sampler" "$ROOT/tools/tsan_sampler.cc" $SRCS -L.. -lammsb -lz -lpthread -Wl,-rpath,"$ROOT/mcmc-ammsb-gpu_b200"
TSAN_OPTIONS=halt_on_error=1 "$OUT/tsan_sampler"
