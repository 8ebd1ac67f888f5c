#!/bin/bash
# Round 2, call AT: render kernel reads the per-env rows (state, attributes) with streaming loads (evict first from its 28 KB of L1).
set -x
tools/ab_checked.sh base stream base stream
