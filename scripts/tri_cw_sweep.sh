#!/bin/bash
# sweep of the compute-warp count of the one-CTA triangular solve (NSG_TRI_CW), device-timed ILU(0) apply on square_h0.0125
for cw in 6 8 12 16 20 24 31; do
  echo -n "cw $cw: "
  NSG_TRI_CW=$cw timeout 120 python scripts/precond_bench.py 0 - 2>/dev/null | python -c "
import json,sys; d=json.load(sys.stdin); print(d['ilu_apply_A_device_ms']['one CTA, shared-memory window'], d['ilu_apply_Mp_device_ms']['one CTA, shared-memory window'])"
done
