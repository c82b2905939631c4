#!/bin/bash
# A/B timing of K6 variants inside ONE gpurun call (box-to-box variance is ~5%): each line = env settings
for rep in 1 2; do
for cfg in "PSB_TC_BACKOFF=0" "PSB_TC_BACKOFF=100" "PSB_TC_BACKOFF=200" "PSB_TC_BACKOFF=400" "PSB_TC_BACKOFF=800"; do
  echo -n "$cfg  "; env $cfg python tools/ablate_k6.py 0 2>&1 | tail -1
done; done
