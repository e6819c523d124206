#!/bin/bash
# tools/with_variant.sh <name|base> <command...>: run a command with the variant library swapped in (GPU box scratch copy)
cd "$(dirname "$0")/.."
L=feature-3dgs_b200/libf3dgs_b200.so
[ -f $L.base ] || cp $L $L.base
n=$1; shift
if [ "$n" = base ]; then cp $L.base $L; else cp feature-3dgs_b200/variants/$n/libf3dgs_b200.so $L; fi
"$@"; rc=$?
cp $L.base $L
exit $rc
