#!/bin/bash
# timing-only experiment matrix on the exp build (results are garbage by design): which epilogue / transform costs matter
export RGFM_LIB_PATH=$PWD/ratio_guided_multimodal_fm_b200/csrc/librgfm_exp.so
run() { echo "=== $*"; env "$@" python tools/profile_step.py 2>&1 | grep -E "conv_tc |graph step|total"; }
run A=0
run RGFM_EXPERIMENT_NO_BIAS=1
run RGFM_TC_NO_STATS=1
run RGFM_EXPERIMENT_NO_BIAS=1 RGFM_TC_NO_STATS=1
run RGFM_EXPERIMENT_NO_BIAS=1 RGFM_TC_NO_STATS=1 RGFM_TC_STACK=2
run RGFM_EXPERIMENT_NO_BIAS=1 RGFM_TC_NO_STATS=1 RGFM_EXPERIMENT_NO_XFORM=1
run RGFM_EXPERIMENT_NO_BIAS=1 RGFM_TC_NO_STATS=1 RGFM_EXPERIMENT_NO_XFORM=1 RGFM_TC_STACK=2
run RGFM_TC_STACK=2
run RGFM_TC_STACK=0
