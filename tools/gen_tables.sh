#!/bin/sh
# Regenerate concentus_b200/csrc/celt_tables_data.inc from the reference headers (dev container only).
set -e
R=${REF:-/root/reference/opus-fix}
HERE=$(cd "$(dirname "$0")" && pwd)
gcc -w -DUSE_ALLOCA -DOPUS_BUILD -DFIXED_POINT=1 -DDISABLE_FLOAT_API -I$R/include -I$R/celt "$HERE/extract_tables.c" -o /tmp/extract_tables -lm
/tmp/extract_tables > "$HERE/../concentus_b200/csrc/celt_tables_data.inc"
