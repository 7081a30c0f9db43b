#!/bin/bash
# Development tool: generate + assemble one variant of the PIP evaluators for tools/pip_microbench.py.
#   tools/mkvariant.sh TAG "ENV=.. ENV2=.." [gen_interp.py arguments]   ->  build/mb/TAG/pip{2b,3b}.cubin
set -e
here=$(cd "$(dirname "$0")" && pwd)
tag=$1; envs=$2; shift 2
out=$here/../build/mb/$tag
mkdir -p "$out"
(cd "$here" && env $envs python gen_interp.py "$@" --out "$out" > "$out/gen.log")
for n in pip2b pip3b; do
    [ -f "$out/${n}_eval.ptx" ] || continue
    ptxas -arch=sm_100a ${PTXAS_OPT:--O3} -v -lineinfo -o "$out/$n.cubin" "$out/${n}_eval.ptx" > "$out/$n.ptxas.log" 2>&1
done
echo "$(grep -h "registers\|spill" "$out"/*.ptxas.log | tr '\n' ' ') <- $tag"
