#!/bin/bash
# tools/sass_dump.sh KERNEL_SUBSTR [out]: SASS of one kernel of libtwilight_b200.so (instruction text only) + opcode histogram
K=${1:-talcoWavefrontKernelILi128ELi1ELi4E}; OUT=${2:-/tmp/sass_$K.txt}
cuobjdump -sass -fun $(cuobjdump -elf twilight_b200/libtwilight_b200.so 2>/dev/null | grep -o "_ZN3twl[A-Za-z0-9_]*${K}[A-Za-z0-9_]*" | sort -u | head -1) twilight_b200/libtwilight_b200.so \
  | grep -E '^\s+/\*[0-9a-f]{4}\*/' | sed -E 's#^\s+/\*([0-9a-f]{4})\*/\s+##; s#\s*/\*.*##' > $OUT
echo "$(wc -l < $OUT) instructions -> $OUT"
