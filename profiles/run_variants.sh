#!/bin/bash
# Runs bench.py once per experiment variant built by profiles/build_variants.sh; one summary line each.
cd "$(dirname "$0")/.."
for lib in "$@"; do
  CMOE_B200_LIB=$PWD/variants/libvar_$lib.so python bench.py --no-cpu-baseline --no-extra 2>/dev/null | tail -1 | \
    python -c "import sys,json; d=json.loads(sys.stdin.read()); print('$lib', 'value=%.4g' % d['value'], 'ms=%.1f' % d['ms_per_step'], 'mc_share=%.3f' % d['roofline']['kernel_share_of_step'], 'frac=%.3f' % d['roofline']['frac'], 'chk=%.12f' % d['kg_checksum'], 'argmax', d['argmax_index'])"
done
