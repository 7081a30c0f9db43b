#!/bin/bash
# Development tool: generate + assemble one variant of the molecule-symmetrised evaluator for tools/pip_microbench.py.
#   tools/mkmsym.sh TAG [gen_msym.py arguments]   ->  build/mb/TAG/pip3b.cubin
set -e
here=$(cd "$(dirname "$0")" && pwd)
tag=$1; shift 1
out=$here/../build/mb/$tag
mkdir -p "$out"
(cd "$here" && python gen_msym.py "$@" --out "$out" > "$out/gen.log")
for n in pip2b pip3b; do
    [ -f "$out/${n}_eval.ptx" ] || continue
    ptxas -arch=sm_100a ${PTXAS_OPT:--O3} -v -lineinfo -o "$out/$n.cubin" "$out/${n}_eval.ptx" > "$out/$n.ptxas.log" 2>&1
done
echo "$(grep -h "registers\|spill" "$out"/*.ptxas.log | tr '\n' ' ') <- $tag"
