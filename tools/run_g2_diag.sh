#!/bin/bash
# 2-GPU check of the column layout on real NVLink with K/G = 128 (the piece size of K = 1024 on 8 GPUs)
export G=2
source <(sed -n '/^run() {/,/^}/p' tools/run_g8_diag.sh)
run cols2_phi2 -- --store columns --K 256
run cols2_phi3 AMMSB_COLS_PHI3=1 AMMSB_BENCH_NO_PARITY=1 -- --store columns --K 256
