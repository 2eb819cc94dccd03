#!/bin/bash
# tools/ab_build.sh NAME [REV] [EXTRA nvcc flags]: builds liburgym_b200 of the working tree (or of git revision REV)
# into ur-gym_b200/variants/NAME.so, for A/B kernel timing on the GPU box:
#   URGYM_B200_LIB=ur-gym_b200/variants/NAME.so python tools/kernel_time.py
set -e
ROOT=$(cd "$(dirname "$0")/.." && pwd)
NAME=$1; REV=$2; EXTRA=$3
SRC=$ROOT/ur-gym_b200/csrc
if [ -n "$REV" ] && [ "$REV" != "-" ]; then
  TMP=$(mktemp -d); git -C $ROOT archive $REV ur-gym_b200/csrc include | tar x -C $TMP; SRC=$TMP/ur-gym_b200/csrc
fi
make -s -j8 -C $SRC BUILD=/tmp/ab_build_$NAME OUT=$ROOT/ur-gym_b200/variants/$NAME.so EXTRA="$EXTRA"
ls -la $ROOT/ur-gym_b200/variants/$NAME.so
