#!/bin/bash
# Development tool: the 2B variants of the molecule-symmetrised evaluator measured in round 2 (build/mb/*/pip2b.cubin).
set -e
here=$(cd "$(dirname "$0")" && pwd)
cd "$here/.."
rm -rf build/mb
mkdir -p build/mb/ship2
cp mbpol_openmm_plugin_b200/csrc/gen/pip2b.cubin mbpol_openmm_plugin_b200/csrc/gen/pip2b_eval.ptx mbpol_openmm_plugin_b200/csrc/gen/pip2b_meta.json build/mb/ship2/
C="--piece 16 --ahead 1 --anchors 1"
bash tools/mkmsym2.sh b2s      --mode "sets:2,3" --k0 12 --k1 12 --npre 8 --chains 2 $C
bash tools/mkmsym2.sh b2s_c1   --mode "sets:2,3" --k0 12 --k1 12 --npre 8 --chains 1 $C
bash tools/mkmsym2.sh b2s_k16  --mode "sets:2,3" --k0 16 --k1 16 --npre 8 --chains 2 $C
bash tools/mkmsym2.sh b2s_k8   --mode "sets:2,3" --k0 8 --k1 8 --npre 8 --chains 2 $C
bash tools/mkmsym2.sh b2s_n12  --mode "sets:2,3" --k0 12 --k1 12 --npre 12 --chains 2 $C
bash tools/mkmsym2.sh b2s_n4   --mode "sets:2,3" --k0 12 --k1 12 --npre 4 --chains 2 $C
bash tools/mkmsym2.sh b2s_b128 --mode "sets:2,3" --k0 12 --k1 12 --npre 8 --chains 2 --block 128 $C
bash tools/mkmsym2.sh b2single --mode single --k0 16 --npre 8 --chains 2 $C
bash tools/mkmsym2.sh b2s3     --mode "sets:2,3;0,1" --k0 12 --k1 12 --k2 4 --npre 8 --chains 2 $C
