#!/bin/bash
# Round 2, call K: edge records packed as 128-bit (base) vs the previous build.
set -x
mkdir -p gpurun_out
tools/ab_checked.sh prev base prev base
