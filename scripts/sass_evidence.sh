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
