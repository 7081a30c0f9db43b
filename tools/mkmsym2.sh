#!/bin/bash
# like mkmsym.sh for the 2B polynomial:  tools/mkmsym2.sh TAG [gen_msym.py arguments]  ->  build/mb/TAG/pip2b.cubin
set -e
here=$(cd "$(dirname "$0")" && pwd)
tag=$1; shift 1
out=$here/../build/mb/$tag
mkdir -p "$out"
(cd "$here" && python gen_msym.py --only pip2b "$@" --out "$out" > "$out/gen.log")
ptxas -arch=sm_100a ${PTXAS_OPT:--O3} -v -lineinfo -o "$out/pip2b.cubin" "$out/pip2b_eval.ptx" > "$out/pip2b.ptxas.log" 2>&1
echo "$(grep -h "registers\|spill" "$out"/pip2b.ptxas.log | tr '\n' ' ') $(grep -o "'fp64_instr': [0-9]*" $out/gen.log | head -1) <- $tag"
