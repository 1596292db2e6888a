#!/bin/bash
cd "$(dirname "$0")/.."
for lib in "$@"; do
  echo -n "$lib: "; CMOE_B200_LIB=$PWD/variants/libvar_$lib.so python profiles/cov_probe.py 5000 0 2>&1 | tail -1
done
