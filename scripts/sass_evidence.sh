#!/bin/bash
# SASS evidence (runs anywhere nvcc's cuobjdump is installed, no GPU): per kernel, how many bulk-async / tensor-copy /
# programmatic-dependent-launch instructions the built library contains.
#   UBLKCP  = cp.async.bulk (1-D bulk copy: node program staging, streamed position tiles)
#   UTMALDG = cp.async.bulk.tensor (TMA tile load: marching-cubes corner stencil)
#   ACQBULK = griddepcontrol.wait, PREEXIT = griddepcontrol.launch_dependents (programmatic dependent launch)
cd "$(dirname "$0")/.."
cuobjdump -sass gsdf_b200/libgsdfb200.so | awk '
  /Function :/ { f=$3 }
  /UBLKCP|UTMALDG|ACQBULK|PREEXIT/ { m=$0; sub(/^ *\/\*[0-9a-f]+\*\/ */,"",m); sub(/ *\/\*.*$/,"",m); sub(/^@!?U?P[0-9] */,"",m); split(m,a," "); c[f" "a[1]]++ }
  END { for (k in c) print c[k], k }' | sort -k2 | while read n f op; do printf "%-12s x%-3s %s\n" "$op" "$n" "$(echo $f | c++filt | cut -c1-150)"; done

# the run-time compiled kernels are not in the library: compile the flange's specialisation here (NVRTC, no GPU) and look at its CUBINs
D=$(mktemp -d)
GSDF_JIT_DUMP=$D python - <<'PY' > /dev/null
import ctypes as C, sys
sys.path.insert(0, ".")
import numpy as np
from gsdf_b200 import gsdf, _lib
b = gsdf.Builder(); s = gsdf.scene(b, "npt-flange"); f = b.flatten(s); aux = np.ascontiguousarray(f["aux"], dtype=np.float32)
assert _lib.lib.gsdf_jit_compile(f["blob"], len(f["blob"]), aux.ctypes.data_as(C.POINTER(C.c_float)), aux.size) > 0, _lib.last_error()
PY
for k in k_jit_grid2 k_jit_grid4 k_jit_grid1 k_jit_centers; do
  [ -f $D/$k.cubin ] || continue
  n=$(cuobjdump -sass $D/$k.cubin | grep -cE "^\s+/\*[0-9a-f]{4,6}\*/")
  for op in UBLKCP ACQBULK PREEXIT; do printf "%-12s x%-3s %s (run-time compiled for npt-flange: %s SASS instructions)\n" $op $(cuobjdump -sass $D/$k.cubin | grep -c $op) $k $n; done
done
head -30 $D/k_jit_grid4.cu > profiles/r02_jit_source_flange_head.txt 2>/dev/null
rm -rf $D
