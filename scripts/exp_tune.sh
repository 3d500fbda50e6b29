#!/bin/bash
# parameter scan of the auxiliary-space preconditioner at n^3 (and C2): scripts/exp_tune.sh n out.log
n=$1; out=$2
run() { env "$@" python scripts/exp_iters.py $n >> $out 2>&1; }
run X=1
run MIMPY_B200_AUX_BW=0.4
run MIMPY_B200_AUX_BW=0.6
run MIMPY_B200_AUX_WJ=0.8
run MIMPY_B200_AUX_WJ=1.25
run MIMPY_B200_AUX_SCALE=1.3
run MIMPY_B200_AUX_SCALE=1.8
run MIMPY_B200_AUX_OMEGA=0.8
run MIMPY_B200_AUX_OMEGA=1.0
