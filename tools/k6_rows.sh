#!/bin/bash
# K6 plan experiments inside ONE gpurun call
run() { env "$@" 2>&1 | tail -1; }
run PSB_TC_LAYOUT=1 python tools/k6_rows.py 360 3 40 off
run PSB_TC_LAYOUT=1 PSB_TC_FLUSH=8 python tools/k6_rows.py 360 3 40 off
run PSB_TC_LAYOUT=1 PSB_TC_FLUSH=6 python tools/k6_rows.py 360 3 40 off
run PSB_TC_LAYOUT=0 PSB_TC_MT=4 python tools/k6_rows.py 360 3 40 off
run PSB_TC_LAYOUT=1 python tools/k6_rows.py 360 3 40 256
run PSB_TC_LAYOUT=0 PSB_TC_MT=1 python tools/k6_rows.py 360 3 40 256
run PSB_TC_LAYOUT=0 PSB_TC_MT=2 python tools/k6_rows.py 360 3 40 256
run PSB_TC_LAYOUT=0 PSB_TC_MT=3 python tools/k6_rows.py 360 3 40 256
run PSB_TC_LAYOUT=0 PSB_TC_MT=4 python tools/k6_rows.py 360 3 40 256
run PSB_TC_LAYOUT=0 PSB_TC_MT=3 python tools/k6_rows.py 512 2 80 256,400
run PSB_TC_LAYOUT=0 PSB_TC_MT=2 python tools/k6_rows.py 512 2 80 256,400
run PSB_TC_LAYOUT=0 PSB_TC_MT=1 python tools/k6_rows.py 512 2 80 256,400
run PSB_TC_LAYOUT=0 python tools/k6_rows.py 400 2 80 256
