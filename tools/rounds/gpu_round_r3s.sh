#!/bin/bash
# Round 2, call AS: the coverage item loop of the rasteriser unrolled 2x / 3x.
set -x
tools/ab_checked.sh base cov2 cov3 base cov2
