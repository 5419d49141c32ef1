#!/bin/bash
# run the SOR probe for every library variant (dev aid): variant_probe.sh <sizes> <kinds>
cd "$(dirname "$0")/.."
for so in 3dgsconverter_b200/lib/variants/libgsx_*.so; do
  name=$(basename $so .so)
  echo "== $name"
  GSX_LIB=$PWD/$so python scripts/sor_probe.py "$1" "$2" 2>&1 | python -c "
import sys, json
for l in sys.stdin:
    try: d=json.loads(l)
    except Exception: print(l.strip()[:200]); continue
    print(d['kind'], d['mode'], 'knn_ms', d['knn_ms'], 'filter_ms', d['filter_ms'], 'scanned', d['scanned_per_pt'])
"
done
