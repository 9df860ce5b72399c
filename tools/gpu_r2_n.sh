#!/bin/bash
for w in 8 16; do for m in wf mega; do timeout 300 python tools/overlap_probe.py $w $m 2>&1 | tail -n 1; done; done
